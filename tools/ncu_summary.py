"""Summarise ncu output for profiles/: launch lists (gpu__time_duration CSV) and --set full captures (.ncu-rep).

    python tools/ncu_summary.py launches <csv> [title]        -> per-kernel table on stdout
    python tools/ncu_summary.py full <name>=<file.ncu-rep> ... -> key metrics per capture (text on stdout, JSON beside)
    python tools/ncu_summary.py dominant <file.ncu-rep> <out.json> <alg_bytes> <alg_flops> <source note>
                                                              -> the per-launch record bench.py puts into roofline.traffic
"""
import collections, csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path, title=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1000.0 if r["Metric Unit"].startswith("ns") or r["Metric Unit"] == "nsecond" else v
        k = r["Kernel Name"]
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1000:10.3f} ms {100 * us / tot:5.1f}% n={n:5d} avg {us / n:9.1f} us  {k[:110]}")
    print(f"total ms {tot / 1000:.3f} launches {sum(v[0] for v in agg.values())}")


def full(pairs):
    out = {}
    for pr in pairs:
        name, path = pr.split("=", 1)
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"## {name}\nkernel: {d.get('Kernel Name', ('?',))[0][:120]}")
        out[name] = {"kernel": d.get("Kernel Name", ("?",))[0]}
        for k in KEYS:
            if k in d:
                print(f"  {k:95s} {d[k][0]:>16s} {d[k][1]}")
                out[name][k] = {"value": d[k][0], "unit": d[k][1]}
        print()
    return out


def dominant(path, out_json, alg_bytes, alg_flops, note):
    d = full([f"dominant={path}"])["dominant"]
    val = lambda k: float(d[k]["value"].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(val(k) * scale[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    dur = val("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(d["gpu__time_duration.sum"]["unit"], 1.0)
    rec = {"source": note, "kernel": d["kernel"], "grid_size": d["launch__grid_size"]["value"], "duration_us": dur,
           "dram_bytes_per_launch": dram, "dram_read_bytes": val("dram__bytes_read.sum") * scale[d["dram__bytes_read.sum"]["unit"]],
           "dram_write_bytes": val("dram__bytes_write.sum") * scale[d["dram__bytes_write.sum"]["unit"]],
           "algorithmic_bytes_per_launch": float(alg_bytes), "algorithmic_flops_per_launch": float(alg_flops),
           "dmma_pipe_active_pct": d["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]["value"],
           "l2_hit_pct": d["lts__t_sector_hit_rate.pct"]["value"]}
    json.dump(rec, open(out_json, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], " ".join(sys.argv[3:]))
    elif sys.argv[1] == "dominant":
        dominant(sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5]), " ".join(sys.argv[6:]))
    else:
        res = full(sys.argv[2:])
        json.dump(res, open("ncu_full_summary.json", "w"), indent=1)
