#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -rA > gpurun_out/gputest_r2.log 2>&1
tail -4 gpurun_out/gputest_r2.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2.csv gpurun_out/launches_r2.md "launch list, round 2 final (third session)" | head -20
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r2_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("ref_gpu"), d.get("vs_ref_gpu"), d["clocks"])
print(d["roofline"]["op_breakdown_ms_per_step"])
P
