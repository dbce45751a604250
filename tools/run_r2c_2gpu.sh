#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k two_rank 2>&1 | tail -8 > gpurun_out/gputest_r2_2gpu.log
cat gpurun_out/gputest_r2_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c3 --steps 30 --warmup 5 > gpurun_out/bench_c3_n2.json 2>> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'P'
import json
for f in ("bench_n2","bench_c3_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e") and d["e2e"]["value"], d.get("graphed"))
    except Exception as e: print(f,"ERR",e)
P
