"""Bring-up: the K=768 input-projection GEMM (6208 x 2304 x 768) on a capped grid, with and without completion counters."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for ctas in (0, 28, 7):
        for done in (0, 1):
            env = dict(os.environ, B2T_GEMM_MAXCTAS=str(ctas), B2T_GEMM_DONE=str(done))
            subprocess.run([sys.executable, __file__, "run"], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch, b2t_pkg
E = b2t_pkg.submodule("engine")
M, N, K = 6208, 2304, 768
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16); bias = torch.randn(N, device="cuda")
for _ in range(3): E.gemm_bf16(a, w, bias=bias)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): E.gemm_bf16(a, w, bias=bias)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ctas = int(os.environ["B2T_GEMM_MAXCTAS"]) or 148
tiles = 49 * 18
print(f"max_ctas={ctas:3d} done={os.environ['B2T_GEMM_DONE']}  {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s  per-tile {ms*1e3/ (tiles/ctas):6.2f} us")
