#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "speculative or graphed" 2>&1 | tail -25
rm -f gpurun_out/bench_cfg_graph.jsonl
for c in c1 c2 c3 c4; do python bench.py --config $c --steps 30 --warmup 5 --no-ref-gpu --no-cpu-baseline >> gpurun_out/bench_cfg_graph.jsonl 2>> gpurun_out/bench_graph.err; done
tail -5 gpurun_out/bench_graph.err
python - <<'P'
import json
for l in open("gpurun_out/bench_cfg_graph.jsonl"):
    try:
        d=json.loads(l); print(d["config"]["workload"][:12], d["ms_per_step"], d["graphed"])
    except Exception as e: print("ERR",e)
P
