"""Extract the tracked numbers from an `ncu -i X.ncu-rep --page raw --csv` export.

usage: python tools/ncu_extract.py RAW.csv LABEL [--evals-refine N] [--source TEXT]
  appends {LABEL: [per-kernel metric dicts]} to profiles/r01_ncu_set_full_extract.json and,
  with --evals-refine, rewrites profiles/r01_refine_traffic.json (read by bench.py)."""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("label")
    ap.add_argument("--evals-refine", type=int, default=0)
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i] + " " + units[i]).strip()
        stall = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i])
                 for i, h in enumerate(hdr)
                 if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h
                 and r[i] not in ("", "n/a")}
        tot = sum(stall.values()) or 1.0
        d["warp_state_samples_pct"] = {k: round(100 * v / tot, 1)
                                       for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:10]}
        out.append(d)
    p = os.path.join(ROOT, "profiles", "r01_ncu_set_full_extract.json")
    allp = json.load(open(p)) if os.path.exists(p) else {}
    allp[a.label] = out
    json.dump(allp, open(p, "w"), indent=1)
    if a.evals_refine:
        ref = [(r, d) for r, d in zip(rows[2:], out) if "refine" in d["Kernel Name"]]
        r, d = ref[-1]
        g = lambda k: float(r[hdr.index(k)])
        to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(g(k) * to_bytes[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        dur = g("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[hdr.index("gpu__time_duration.sum")]]
        t = {"kernel": d["Kernel Name"].replace("void ", "").split("(")[0],
             "workload": "bench.py N=1 (1048576 seeds, 16 views 1280x960, mu=7)",
             "source": a.source or a.raw,
             "dram_bytes_per_launch": dram,
             "warp_inst_per_launch": g("smsp__inst_executed.sum"),
             "evals_per_launch": a.evals_refine,
             "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "duration_ms_under_ncu": dur}
        json.dump(t, open(os.path.join(ROOT, "profiles", "r01_refine_traffic.json"), "w"), indent=1)
        print(json.dumps(t, indent=1))


if __name__ == "__main__":
    main()
