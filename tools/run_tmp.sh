#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err
VOGE_NO_SEGMENT_ALIGN=1 python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_noal.json 2>> gpurun_out/bench_al.err
python - <<'P'
import json
for f in ("bench_al","bench_noal"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["loss"], d["peak_hbm_gb"])
    o=d["roofline"]["op_breakdown_ms_per_step"]; print({k:round(v,3) for k,v in o.items() if k in ('voge_trace_hits','voge_select_topk','voge_bin_count','voge_bin_fill','bin_views')})
P
