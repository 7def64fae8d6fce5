"""Extract the tracked numbers from an `ncu -i X.ncu-rep --page raw --csv` export.

usage: python tools/ncu_extract.py RAW.csv LABEL [--evals-refine N] [--source TEXT]
                                    [--hbm-evals-score N --hbm-evals-refine M]
  appends {LABEL: [per-kernel metric dicts]} to profiles/r02_ncu_set_full_extract.json and,
  with --evals-refine, rewrites profiles/r02_refine_traffic.json; with --hbm-evals-score,
  profiles/r02_hbm_traffic.json (both read by bench.py).  Each carries src_sha256 = the hash of
  the CUDA sources in the tree (densepoints_b200.build.source_hash) -- run this on the same tree
  the capture was made from; bench.py reports the numbers only while the hash still matches."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densepoints_b200.build import source_hash  # noqa: E402
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
    ap.add_argument("--hbm-evals-score", type=int, default=0)
    ap.add_argument("--hbm-evals-refine", type=int, default=0)
    ap.add_argument("--hbm-refine-raw", default="", help="raw csv of a light metrics pass over the refine launch")
    ap.add_argument("--src-sha256", default="", help="hash printed by the profiled run (default: this tree)")
    a = ap.parse_args()
    sha = a.src_sha256 or source_hash()
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
    p = os.path.join(ROOT, "profiles", "r02_ncu_set_full_extract.json")
    allp = json.load(open(p)) if os.path.exists(p) else {}
    allp[a.label] = out
    json.dump(allp, open(p, "w"), indent=1)
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    to_ms = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}

    def pick(word, first=False):
        ref = [(r, d) for r, d in zip(rows[2:], out) if word in d["Kernel Name"]]
        r, d = ref[0] if first else ref[-1]
        g = lambda k: float(r[hdr.index(k)])
        dram = sum(g(k) * to_bytes[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        dur = g("gpu__time_duration.sum") * to_ms[units[hdr.index("gpu__time_duration.sum")]]
        return d, g, dram, dur

    if a.hbm_evals_score:
        ds, gs, dram_s, dur_s = pick("score", first=True)   # score launch, then the filter launch
        t = {"workload": "bench.py roofline_hbm leg (64 views 1920x1080, 1.5 M patches, mu=7)",
             "source": a.source or a.raw, "src_sha256": sha,
             "score_kernel": ds["Kernel Name"].replace("void ", "").split("(")[0],
             "score_dram_bytes_per_launch": dram_s, "score_evals_per_launch": a.hbm_evals_score,
             "score_duration_ms_under_ncu": dur_s,
             "score_warp_inst_per_launch": gs("smsp__inst_executed.sum")}
        if a.hbm_evals_refine:
            if a.hbm_refine_raw:  # the 0.6 s refine launch: DRAM bytes and duration only (one pass)
                rows2 = list(csv.reader(open(a.hbm_refine_raw)))
                h2, u2 = rows2[0], rows2[1]
                r2 = [r for r in rows2[2:] if "refine" in r[h2.index("Kernel Name")]][-1]
                g2 = lambda k: float(r2[h2.index(k)])
                dram_r = sum(g2(k) * to_bytes[u2[h2.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                dur_r = g2("gpu__time_duration.sum") * to_ms[u2[h2.index("gpu__time_duration.sum")]]
                dr = {"Kernel Name": r2[h2.index("Kernel Name")]}
            else:
                dr, gr, dram_r, dur_r = pick("refine")
            t.update({"refine_kernel": dr["Kernel Name"].replace("void ", "").split("(")[0],
                      "refine_dram_bytes_per_launch": dram_r,
                      "refine_evals_per_launch": a.hbm_evals_refine,
                      "refine_duration_ms_under_ncu": dur_r})
        json.dump(t, open(os.path.join(ROOT, "profiles", "r02_hbm_traffic.json"), "w"), indent=1)
        print(json.dumps(t, indent=1))
    if a.evals_refine:
        d, g, dram, dur = pick("refine")
        t = {"kernel": d["Kernel Name"].replace("void ", "").split("(")[0],
             "workload": "bench.py N=1 (1048576 seeds, 16 views 1280x960, mu=7)",
             "source": a.source or a.raw, "src_sha256": sha,
             "dram_bytes_per_launch": dram,
             "warp_inst_per_launch": g("smsp__inst_executed.sum"),
             "evals_per_launch": a.evals_refine,
             "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "duration_ms_under_ncu": dur}
        json.dump(t, open(os.path.join(ROOT, "profiles", "r02_refine_traffic.json"), "w"), indent=1)
        print(json.dumps(t, indent=1))


if __name__ == "__main__":
    main()
