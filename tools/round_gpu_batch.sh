#!/bin/bash
# One GPU-box call with everything a round's evidence needs (outputs under gpurun_out/, tag = $1): GPU tests, both bench arms, the ncu launch
# list of the bench command, two ncu --set full captures of the dominant kernel (attack and sustain chunks), the other BASELINE configs,
# racecheck of the lane-tiled kernel on a tiny workload.
TAG=${1:-rXX}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${TAG}_pytest.log
python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > $O/${TAG}_bench_under_ncu.json 2>&1
ncu --set full --import-source on --clock-control none -k regex:chain_tile --launch-skip 0 -c 1 -o $O/${TAG}_tile_attack -f python tools/prof_run.py 0.5 1 0.5 > $O/${TAG}_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:chain_tile --launch-skip 2 -c 1 -o $O/${TAG}_tile_sustain -f python tools/prof_run.py 0.5 1 0.5 >> $O/${TAG}_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tremolo_group_tile --launch-skip 2 -c 1 -o $O/${TAG}_osc -f python tools/trem_time.py >> $O/${TAG}_ncu.log 2>&1
timeout 300 python tools/pa_time.py 8128 0.5 > $O/${TAG}_pa_time.json 2> $O/${TAG}_pa_time.err
timeout 900 python tools/bench_configs.py > $O/${TAG}_bench_configs.json 2> $O/${TAG}_bench_configs.err
timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=chain_tile --log-file $O/${TAG}_racecheck_chain_tile.log python tools/racecheck_run.py > $O/${TAG}_racecheck_chain_tile.out 2>&1
tail -3 $O/${TAG}_pytest.log
python3 - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms_per_step"])
for k,v in d.get("variants",{}).items(): print(k[:70], v["value"], v["ms_per_step"])
r=json.loads(open("$O/${TAG}_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"])
PY
OWG_CHAIN_KERNEL=split timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=chain_split --log-file $O/${TAG}_racecheck_chain_split.log python tools/racecheck_run.py > $O/${TAG}_racecheck_chain_split.out 2>&1
grep -c "Race reported" $O/${TAG}_racecheck_chain_tile.log; tail -1 $O/${TAG}_racecheck_chain_tile.log; tail -1 $O/${TAG}_racecheck_chain_split.log
