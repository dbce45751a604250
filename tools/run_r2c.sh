#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -rA > gpurun_out/gputest_r2.log 2>&1
tail -3 gpurun_out/gputest_r2.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err
for c in c1 c2 c3 c4; do python bench.py --config $c --steps 30 --warmup 5 >> gpurun_out/bench_r2_configs.jsonl 2>> gpurun_out/bench_final.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2.csv gpurun_out/launches_r2.md "launch list, round 2 final (third session)" | head -14
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r2_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["ref_gpu"]["value"], d["vs_ref_gpu"]["ratio"], d["clocks"], d["cpu_baseline"]["value"], d["gpu_launches_per_step"], d["peak_hbm_gb"])
print(d["roofline"]["op_breakdown_ms_per_step"])
r=d["roofline"]; print(r["frac"], r["valid_pair_frac"], r["share_of_step"], r["whole_step"])
for l in open("gpurun_out/bench_r2_configs.jsonl"):
    c=json.loads(l); print(c["config"]["workload"][:12], round(c["ms_per_step"],3), round(c["graphed"]["ms_per_step"],3))
P
