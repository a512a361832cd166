#!/bin/bash
cd /root/repo
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('exposed_comm_ms_per_step'), d['e2e']['value'])"; }
echo warm; run 29520
echo default; run 29521
echo serial; B2T_BENCH_COMM_SERIAL=1 run 29522
echo nch32; NCCL_MIN_NCHANNELS=32 run 29523
echo nch32+sms32; NCCL_MIN_NCHANNELS=32 B2T_COMM_SMS=32 run 29524
echo nvls_off; NCCL_NVLS_ENABLE=0 run 29525
echo debug; NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras 2>&1 | grep -i "channel\|NVLS\|algo\|Connected\|nranks" | head -12
