"""CPU: libgegp.so loads and exports every symbol include/gegp.h declares; host-only entry points behave."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "gegp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gegp_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_contract():
    names = _declared_functions()
    for must in ("gegp_build_cov", "gegp_cross_cov", "gegp_potrf", "gegp_trsm_rows", "gegp_lml_eval",
                 "gegp_predict_setup", "gegp_predict", "gegp_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from gpgradpy_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/gegp.h but not exported"
    assert set(_lib.EXPORTS) == set(_declared_functions())


def test_host_only_entry_points():
    from gpgradpy_b200 import _lib
    lib = _lib.load()
    assert lib.gegp_abi_version() == _lib.ABI_VERSION
    assert lib.gegp_ld(5500) == 5504 and lib.gegp_ld(21000) == 21008 and lib.gegp_ld(16) == 16
    n, d = 500, 10
    N = n * (d + 1)
    w0 = lib.gegp_workspace_bytes(_lib.OP_LML, n, n, d, 1)
    w1 = lib.gegp_workspace_bytes(_lib.OP_LML_GRAD, n, n, d, 1)
    assert w0 >= (N + 2) * 5504 * 8 and w1 >= w0 + 2 * N * 5504 * 8
    assert lib.gegp_workspace_bytes(_lib.OP_LML, 0, 0, d, 1) == 0          # bad geometry
    assert lib.gegp_workspace_bytes(_lib.OP_PREDICT, n, n, d, 100) == 100 * 5504 * 8


def test_argument_checks_return_negative_codes_without_touching_the_gpu():
    from gpgradpy_b200 import _lib
    lib = _lib.load()
    assert lib.gegp_build_cov(0, 0, 3, 0, 0, 0, 0, 0.0, 0, 0, 0.0, 1.0, 0, 0, 0, 0, 0) == -1      # n <= 0
    assert lib.gegp_build_cov(4, 4, 3, 0, 0, 0, 0, 0.0, 0, 0, 0.0, 1.0, 0, 0, 0, 0, 0) == -4      # X is NULL
    # kernel family / kernel hyper-parameter: -50 (unknown family; rational-quadratic alpha must be positive)
    assert lib.gegp_build_cov(4, 4, 3, 16, 0, 16, 7, 0.0, 0, 0, 0.0, 1.0, 16, 16, 0, 0, 0) == -50
    assert lib.gegp_build_cov(4, 4, 3, 16, 0, 16, _lib.KERNEL_RATQUAD, 0.0, 0, 0, 0.0, 1.0, 16, 16, 0, 0, 0) == -50
    assert lib.gegp_lml_eval(1, 16, 0, _lib.KERNEL_RATQUAD, 0, 4, 4, 3, 16, 0, 16, 0, 1, 0.0, 0, 0.0, 0, 16, 0, 16, 1 << 20, 0) == -50
    assert lib.gegp_potrf(0, 0, 0, 0, 0, 0, 0) == -1
    assert lib.gegp_potrf(8, 0, 0, 8, 0, 0, 0) == -3
    assert lib.gegp_lml_eval(0, 0, 0, 0, 0, 4, 4, 3, 0, 0, 0, 0, 1, 0.0, 0, 0.0, 0, 0, 0, 0, 0, 0) == -1
    assert lib.gegp_trsm_rows(8, 0, 8, 0, 0, 8, 1, 0) == -2
    assert lib.gegp_potri(8, 0, 8, 0, 0, 8, 0, 8, 0) == -2
    assert lib.gegp_dgemm(0, 4, 4, 4, 1.0, 0, 4, 0, 4, 0.0, 0, 4, 0) == -6
    assert lib.gegp_dinv_doubles(129) == 2 * 128 * 128
    assert lib.gegp_set_option(99, 1) == -1 and lib.gegp_set_option(_lib.OPT_TMA_MIN_TILES, 0) == -2
    # early pieces of the explicit inverse: 0 off, 1 / 2 forced level, 3 by problem size (default); returns the old value
    assert lib.gegp_set_option(_lib.OPT_INV_EARLY, 4) == -2 and lib.gegp_set_option(_lib.OPT_INV_EARLY, -1) == -2
    old = lib.gegp_set_option(_lib.OPT_INV_EARLY, 0)
    assert old in (0, 1, 2, 3) and lib.gegp_set_option(_lib.OPT_INV_EARLY, old) == 0
    # condition-number / surrogate-derivative entry points
    import ctypes as C
    assert lib.gegp_symv(0, 0, 0, 0, 0, 0) == -1 and lib.gegp_symv(8, 0, 8, 0, 0, 0) == -2
    assert lib.gegp_row_abs_sum(8, 0, 8, 0, 0) == -2 and lib.gegp_row_sq_sum(8, 16, 7, 0, 0) == -3
    assert lib.gegp_lanczos_step(8, 300, 0, 8, 0, 0, 0, 0) == -2 and lib.gegp_lanczos_step(8, 0, 0, 8, 0, 0, 0, 0) == -3
    assert lib.gegp_lincomb(8, 0, 0, 8, 0, 0, 0) == -2
    assert lib.gegp_quad_grad(4, 4, 3, 16, 0, 16, 0, 0.0, 16, _lib.MODE_PRECON, 0.0, 0, 0, 16, 16, 1 << 20, 0) == -8   # base only
    assert lib.gegp_weighted_grad(4, 4, 3, 16, 0, 16, 0, 0.0, 0, 16, _lib.MODE_BASE, 0.0, 0, 0, 16, 16, 1 << 20, 0) == -7  # W NULL
    assert lib.gegp_quad_grad_work_bytes(4, 4, 3) > 0 and lib.gegp_quad_grad_work_bytes(0, 0, 3) == 0
    assert lib.gegp_predict_grad(4, 4, 3, 0, 0, 0, 0, 0.0, 0, 16, 0, 0, 1, 0.0, 1.0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0) == -4
    assert lib.gegp_predict_hess(4, 4, 3, 16, 0, 16, 0, 0.0, 16, 16, 16, 16, 0, 1, 0.0, 1.0, 16, 16, 16, 0, 16, 16, 16, 0, 0, 0, 0) == -11
    out8 = (C.c_int64 * 8)()
    assert lib.gegp_lml_layout(500, 500, 10, 1, 1, out8) == 0
    header, ld, per, offA, offP, offD, offU, offK = (int(v) for v in out8)
    assert header == 256 and ld == 5504 and offA == 0 and offU > offD > offP > offA and offK == offU + 5500 * ld
    assert 8 * per + header == lib.gegp_workspace_bytes(_lib.OP_LML_GRAD, 500, 500, 10, 1)
    assert lib.gegp_lml_layout(500, 500, 10, 0, 1, out8) == 0 and out8[6] == -1 and out8[7] == -1


def test_no_cpu_fallback_when_library_is_missing(monkeypatch):
    from gpgradpy_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgegp.so")
    with pytest.raises(RuntimeError):
        _lib.load()


def test_integration_doc_stub_matches_the_abi():
    """The reference-side ctypes stub shown in INTEGRATION.md binds gegp_lml_eval with the prototype of this ABI version,
    names every exported entry point, and uses the output slots of include/gegp.h."""
    from gpgradpy_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in _declared_functions():
        assert name in doc, f"{name} is not mentioned in INTEGRATION.md"
    m = re.search(r"_lib\.gegp_lml_eval\.argtypes = \(\[(.*?)\]\)", doc, flags=re.S)
    assert m, "stub prototype not found"
    stub = re.findall(r"C\.c_(\w+)", m.group(1))
    lib = _lib.load() if os.path.exists(_lib.LIB_PATH) else None
    assert lib is not None
    kinds = {ctypes.c_int: "int", ctypes.c_double: "double", ctypes.c_size_t: "size_t", ctypes.c_void_p: "void_p"}
    want = [kinds.get(t, "void_p") for t in lib.gegp_lml_eval.argtypes]   # every pointer type binds as void_p in the stub
    assert stub == want
    assert f"OUT_GRAD = 0, 1, 2, 4, {_lib.OUT_GRAD}" in doc
    hdr = open(os.path.join(ROOT, "include", "gegp.h")).read()
    assert re.search(rf"#define GEGP_OUT_GRAD\s+{_lib.OUT_GRAD}\b", hdr)
