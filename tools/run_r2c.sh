#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "speculative or isotropic" 2>&1 | tail -3
rm -f gpurun_out/bench_cfg_spec.jsonl gpurun_out/bench_cfg_nospec.jsonl
python bench.py --steps 10 --warmup 3 --views 8 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_spec_v8.json 2>> gpurun_out/bench_spec.err
VOGE_NO_SPECULATION=1 python bench.py --steps 10 --warmup 3 --views 8 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_nospec_v8.json 2>> gpurun_out/bench_nospec.err
for c in c1 c2 c3 c4; do python bench.py --config $c --steps 30 --warmup 5 --no-ref-gpu --no-cpu-baseline >> gpurun_out/bench_cfg_spec.jsonl 2>> gpurun_out/bench_spec.err; VOGE_NO_SPECULATION=1 python bench.py --config $c --steps 30 --warmup 5 --no-ref-gpu --no-cpu-baseline >> gpurun_out/bench_cfg_nospec.jsonl 2>> gpurun_out/bench_nospec.err; done
python - <<'P'
import json
for f in ("bench_spec_v8","bench_nospec_v8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
for f in ("bench_cfg_spec","bench_cfg_nospec"):
    for l in open("gpurun_out/%s.jsonl"%f):
        try:
            d=json.loads(l); print(f, d["config"]["workload"][:12], d["ms_per_step"])
        except Exception as e: print(f,"ERR",e)
P
V=8 python tools/step_phases.py 2>&1 | grep -v "^k:" | tail -12
