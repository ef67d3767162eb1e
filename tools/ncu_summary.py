#!/usr/bin/env python3
"""Summarise one kernel launch of an ncu report (--set full) as JSON: duration, DRAM bytes, issue-slot use, warps per scheduler, FP64 /
FP32 pipe utilisation, stall reasons per issued instruction, registers, grid.  Usage: ncu_summary.py report.ncu-rep out.json [note]"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.split("\n")))
h, u, v = rows[0], rows[1], rows[2]
ix = {k: i for i, k in enumerate(h)}
def num(k, scale=1.0):
    if k not in ix: return None
    try: x = float(v[ix[k]].replace(",", ""))
    except ValueError: return None
    unit = u[ix[k]]
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(unit, 1.0)
    return x * mult * scale
res = {
    "report": rep.split("/")[-1], "note": note, "kernel": v[ix["Kernel Name"]][:120] if "Kernel Name" in ix else None,
    "duration_s": num("gpu__time_duration.sum"),
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "grid": num("launch__grid_size"), "block": num("launch__block_size"), "registers_per_thread": num("launch__registers_per_thread"),
    "issue_active_per_cycle": num("smsp__issue_active.avg.per_cycle_active"),
    "warps_active_per_scheduler": num("smsp__warps_active.avg.per_cycle_active"),
    "fp64_pipe_pct_of_peak": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct_of_peak": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct_of_peak": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": num("smsp__inst_executed.sum"),
    "threads_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "stall_cycles_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v[i]) for i, k in enumerate(h)
                               if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[i] not in ("", "n/a")},
}
if res["dram_bytes_read"] is not None and res["dram_bytes_write"] is not None:
    res["dram_bytes"] = res["dram_bytes_read"] + res["dram_bytes_write"]
    if res["duration_s"]: res["dram_GBps"] = res["dram_bytes"] / res["duration_s"] / 1e9
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:600])
