"""Install the UNMODIFIED reference (marchildon/gpgradpy v1.3.2, /root/reference) into baseline/_ref so that it travels
to the GPU box, where bench.py times it as the CPU arm (`--impl reference`, cpu_baseline kind "reference").

    python oracle/install_ref.py      # build container only; baseline/_ref/ is git-ignored, not gpurun-ignored

The reference imports `smt` (Surrogate Modeling Toolbox, unpinned, not in this image) at module scope
(gpgradpy/src/optz/GpHparaX0.py:12) although the likelihood path never calls it; oracle/ref_shim/smt -- a 15-line
stand-in for smt.sampling_methods.LHS backed by scipy.stats.qmc -- is copied beside the package.  No reference source
is copied into the repository's history.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force: bool = False) -> bool:
    """True when baseline/_ref holds an importable reference afterwards."""
    pkg = os.path.join(DST, "gpgradpy", "src", "GaussianProcess.py")
    if os.path.exists(pkg) and not force:
        return True
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
           "--find-links", "/opt/wheelhouse", "--target", DST, REF]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:      # the source tree is read-only: retry from a scratch copy
        tmp = "/tmp/gpgradpy_ref_src"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(REF, tmp)
        cmd[-1] = tmp
        r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        return False
    shim = os.path.join(DST, "smt")
    shutil.rmtree(shim, ignore_errors=True)
    shutil.copytree(os.path.join(HERE, "ref_shim", "smt"), shim)
    return os.path.exists(pkg)


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "ok" if ok else "unavailable")
    sys.exit(0 if ok else 1)
