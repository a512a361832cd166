#!/bin/bash
# bring-up: data-parallel step time on 2 GPUs under NCCL algorithm / protocol choices
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('exposed_comm_ms_per_step'), d['e2e']['value'])"; }
echo warm; run 29520
echo default; run 29521
echo ring; NCCL_ALGO=Ring run 29522
echo nvls; NCCL_ALGO=NVLS run 29523
echo simple; NCCL_PROTO=Simple run 29524
echo ll128; NCCL_PROTO=LL128 run 29525
echo ctas8; NCCL_MAX_CTAS=8 run 29526
echo ctas16; NCCL_MAX_CTAS=16 run 29527
