#!/usr/bin/env python3
"""Fold the artefacts of one tools/round_gpu_batch.sh run (gpurun_out/<tag>_*) into profiles/<name>_* : bench lines, the ncu launch list as
kernel shares, JSON summaries of the ncu --set full captures (tools/ncu_summary.py) and the racecheck log.
Usage: collect_profiles.py <tag> <name>     e.g.  collect_profiles.py r2_f2 r02_v3"""
import collections, csv, json, os, shutil, subprocess, sys
tag, name = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
here = os.path.dirname(os.path.abspath(__file__))

def last_json_line(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return line
    raise SystemExit(f"no JSON line in {path}")

for src, dst in (("bench.json", "bench_full.json"), ("bench_ref.json", "bench_reference_arm.json")):
    open(f"{P}/{name}_{dst}", "w").write(last_json_line(f"{G}/{tag}_{src}") + "\n")
if os.path.exists(f"{G}/{tag}_bench_configs.json"):
    shutil.copy(f"{G}/{tag}_bench_configs.json", f"{P}/{name}_bench_configs.json")

# launch list -> shares per kernel (serialised, cold-cache times: only the SHARE is meaningful)
tot = collections.defaultdict(lambda: [0.0, 0])
with open(f"{G}/{tag}_launches.csv") as f:
    rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
for r in rows[1:]:
    ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
    k = r[ik].split("(")[0]
    tot[k][0] += ns; tot[k][1] += 1
allns = sum(v[0] for v in tot.values())
with open(f"{P}/{name}_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 700: python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline "
            "(serialised, cold-cache launch times: the kernel's SHARE is what must agree with the bench line)\n")
    for k, (ns, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
        f.write(f"{ns / 1e6:8.2f} ms {100 * ns / allns:5.1f}% n={n:4d} {k[:150]}\n")

for rep, out, note in ((f"{tag}_tile_attack.ncu-rep", f"{name}_chain_tile_kernel_attack.json", "chunk launch 0 (attack) of tools/prof_run.py 0.5 1 0.5: 8128 renders, tremolo 0.5"),
                       (f"{tag}_tile_sustain.ncu-rep", f"{name}_chain_tile_kernel_sustain.json", "chunk launch 2 (sustain) of tools/prof_run.py 0.5 1 0.5: 8128 renders, tremolo 0.5, 8192 base samples"),
                       (f"{tag}_osc.ncu-rep", f"{name}_tremolo_group_tile_kernel.json", "launch 2 of tools/trem_time.py: one Twin-T oscillator, 16 384 preamp-rate steps")):
    if os.path.exists(f"{G}/{rep}"):
        subprocess.run([sys.executable, f"{here}/ncu_summary.py", f"{G}/{rep}", f"{P}/{out}", note], check=True, stdout=subprocess.DEVNULL)
if os.path.exists(f"{P}/{name}_chain_tile_kernel_sustain.json"):
    d = json.load(open(f"{P}/{name}_chain_tile_kernel_sustain.json"))
    # rows read once and written once in place (8128 x 8192 x 8 B x 2) + 16 384 per-sample matrix records of 190 doubles read once
    d["algorithmic_bytes"] = 8128 * 8192 * 8 * 2 + 16384 * 190 * 8
    json.dump(d, open(f"{P}/chain_kernel_traffic.json", "w"), indent=1)  # bench.py reads roofline.traffic from here
if os.path.exists(f"{G}/{tag}_racecheck_chain_tile.log"):
    shutil.copy(f"{G}/{tag}_racecheck_chain_tile.log", f"{P}/{name}_sanitizer_racecheck_chain_tile.log")
for src, dst in (("racecheck_chain_split.log", "sanitizer_racecheck_chain_split.log"), ("pa_time.json", "pa_time.json")):
    if os.path.exists(f"{G}/{tag}_{src}"):
        shutil.copy(f"{G}/{tag}_{src}", f"{P}/{name}_{dst}")
print("collected", sorted(x for x in os.listdir(P) if x.startswith(name)))
