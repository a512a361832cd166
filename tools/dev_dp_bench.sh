#!/bin/bash
# data-parallel bench lines on N GPUs of one box: weak scaling, strong scaling (global batch 64), and the DP equivalence test (N = 2)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -2; fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null > gpurun_out/r2h_bench_n$N.json; cut -c1-420 gpurun_out/r2h_bench_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --scaling strong 2>/dev/null > gpurun_out/r2h_bench_n${N}_strong.json; cut -c1-300 gpurun_out/r2h_bench_n${N}_strong.json
python - <<PY
import json
for f in ("gpurun_out/r2h_bench_n$N.json", "gpurun_out/r2h_bench_n${N}_strong.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["value"], d["ms_per_step"], d.get("exposed_comm_ms_per_step"), d["e2e"]["value"])
PY
