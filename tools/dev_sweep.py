"""Time the bench training step under different recurrence geometries (env knobs of engine.cu); one subprocess per config.

    python tools/dev_sweep.py "B2T_REC_BG_BWD=16 B2T_REC_LANES_BWD=1" "B2T_REC_CHUNKS=6" ...
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch, b2t_pkg, bench
E = b2t_pkg.submodule("engine")
from torch_cpu_port import PortModel
torch.manual_seed(0)
cfg = E.make_config(**bench.CFG)
eng = E.Engine(cfg, E.flat_from_state_dict(cfg, PortModel(**bench.CFG).state_dict()).cuda(), max_batch=64, max_T=400, max_label_len=64, training=True)
bs = [{k: v.cuda() for k, v in b.items()} for b in bench.synth_batches(1, 4)]
in_len = torch.full((64,), 97, dtype=torch.int32)
def step(i):
    hb = bs[i %% 4]
    eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
    l = eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
    eng.backward()
    eng.optimizer_step([1e-3] * 3, [0, 0, 1e-3], 0.9, 0.999, 0.1, 10.0)
    return l
for i in range(5): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record(); h0 = time.perf_counter()
for i in range(20): l = step(i)
h1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
torch.cuda.synchronize()
h2 = time.perf_counter(); step(0); h3 = time.perf_counter(); torch.cuda.synchronize(); h4 = time.perf_counter()
print("MS_PER_STEP %%.3f host_enqueue_ms(queue backed up) %%.3f single step: host_enqueue %%.3f ms, until done %%.3f ms loss %%.4f" %% (e0.elapsed_time(e1) / 20, (h1 - h0) * 50, (h3 - h2) * 1e3, (h4 - h2) * 1e3, l.mean().item()))
''' % (ROOT, ROOT)
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        if "=" in kv:
            k, v = kv.split("=", 1); env[k] = v
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
        line = [l for l in out.stdout.splitlines() if l.startswith("MS_PER_STEP")]
        print(f"{spec:70s} -> {line[0] if line else 'FAILED ' + out.stderr[-300:]}", flush=True)
    except subprocess.TimeoutExpired:
        print(f"{spec:70s} -> TIMEOUT", flush=True)
