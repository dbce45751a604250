#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_bc.json 2> gpurun_out/bench_bc.err
python bench.py --steps 10 --warmup 3 --views 8 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_bc_v8.json 2>> gpurun_out/bench_bc.err
python bench.py --steps 10 --warmup 3 --views 2 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_bc_v2.json 2>> gpurun_out/bench_bc.err
python - <<'P'
import json
for f in ("bench_bc","bench_bc_v8","bench_bc_v2"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["loss"])
    o=d["roofline"]["op_breakdown_ms_per_step"]; print({k:round(v,4) for k,v in o.items() if 'bin' in k or 'trace' in k})
P
