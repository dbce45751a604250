#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/scale_r2_n1.json 2> gpurun_out/scale.err
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/scale_r2_n$n.json 2>> gpurun_out/scale.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --config c3 --steps 30 --warmup 5 > gpurun_out/bench_c3_n8.json 2>> gpurun_out/scale.err
python - <<'P'
import json
for f in ("scale_r2_n1","scale_r2_n2","scale_r2_n4","scale_r2_n8","bench_c3_n8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), d.get("e2e") and round(d["e2e"]["value"],1), d.get("graphed"))
    except Exception as e: print(f,"ERR",e)
P
