#!/bin/bash
# bring-up: exposed communication on N GPUs: bucketed vs one collective
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('exposed_comm_ms_per_step'), d['e2e']['value'])"; }
echo bucketed; B2T_DP_BUCKETS=bucketed run 29541
echo single; B2T_DP_BUCKETS=single run 29542
