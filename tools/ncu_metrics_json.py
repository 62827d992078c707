#!/usr/bin/env python3
"""Write profiles/ncu_solve_metrics.json and profiles/ncu_traffic.json (read by bench.py for the `roofline` object) from
the raw page of an `ncu --set full` capture of ONE launch of the solve kernel:

    ncu -i gpurun_out/solve_full.ncu-rep --page raw --csv > gpurun_out/solve_full_raw.csv
    python tools/ncu_metrics_json.py gpurun_out/solve_full_raw.csv 4096 "capture description"
"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, nqp = sys.argv[1], int(sys.argv[2])
desc = sys.argv[3] if len(sys.argv) > 3 else raw
rows = list(csv.reader(open(raw)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
names, units, vals = rows[h], rows[h + 1], rows[h + 2]


def get(suffix, scale=True):
    for k, u, v in zip(names, units, vals):
        if k.endswith(suffix) and v != "":
            x = float(v.replace(",", ""))
            if scale:
                x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u.split("/")[0], 1.0)
            return x
    return None


stalls = {}
for k, v in zip(names, vals):
    if "warps_issue_stalled_" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v != "":
        r = k.split("warps_issue_stalled_")[1].replace("_per_issue_active.ratio", "")
        if r != "selected" and float(v) >= 0.25:
            stalls[r] = round(float(v), 2)
stalls = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
inst = get("smsp__inst_executed.sum", False) or get("sm__inst_executed.sum", False)
dur = get("gpu__time_duration.sum")
kern = vals[names.index("Kernel Name")]
metrics = {
    "capture": desc, "kernel": kern,
    "warp_instructions_per_qp": round(inst / nqp) if inst else None,
    "issue_slots_busy_pct": round(get("sm__inst_issued.avg.pct_of_peak_sustained_active", False), 1),
    "fp64_pipe_active_pct": round(get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False), 1),
    "warps_active_pct_of_peak": round(get("sm__warps_active.avg.pct_of_peak_sustained_active", False), 1),
    "l2_sector_hit_rate_pct": round(get("lts__t_sector_hit_rate.pct", False), 1),
    "dram_read_mb": round(rd / 1e6, 1), "dram_write_mb": round(wr / 1e6, 1), "dram_bytes_per_qp": round((rd + wr) / nqp),
    "registers_per_thread": int(get("launch__registers_per_thread", False)),
    "dynamic_smem_kb_per_cta": round(get("launch__shared_mem_per_block_dynamic") / 1e3, 1),
    "stall_cycles_per_issue": stalls, "duration_ms": round(dur, 2), "qps_per_launch": nqp,
}
traffic = {"source": desc, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "dram_bytes_per_qp": round((rd + wr) / nqp),
           "qps_per_launch": nqp}
json.dump(metrics, open(os.path.join(ROOT, "profiles", "ncu_solve_metrics.json"), "w"), indent=1)
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(metrics, indent=1))
