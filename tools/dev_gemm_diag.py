"""GPU bring-up diagnostics for the tcgen05 GEMM (not a test; prints error structure)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b2t_pkg
E = b2t_pkg.submodule("engine")

def run(M, N, K, a_mn, b_mn, pattern="rand"):
    g = torch.Generator(device="cuda").manual_seed(1)
    if pattern == "rand":
        A = torch.randn(M, K, device="cuda", generator=g)
        B = torch.randn(N, K, device="cuda", generator=g)
    else:  # small-integer matrices: exact in bf16, errors are structural not rounding
        A = torch.randint(-3, 4, (M, K), device="cuda", generator=g).float()
        B = torch.randint(-3, 4, (N, K), device="cuda", generator=g).float()
    A = A.to(torch.bfloat16); B = B.to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    try:
        out = E.gemm_bf16(Ain, Bin, a_mn=a_mn, b_mn=b_mn)
        torch.cuda.synchronize()
    except Exception as ex:
        print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: EXC {ex}")
        return False
    err = (out - ref).abs()
    bad = (err > 1e-2 * max(1.0, ref.abs().max().item() / 50))
    ok = not bool(bad.any())
    print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} {pattern}: max_err={err.max().item():.4g} bad_frac={bad.float().mean().item():.4f} {'OK' if ok else 'FAIL'}")
    if not ok:
        rows_bad = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols_bad = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows_bad, "cols", cols_bad)
        print("   out[0:4,0:8]", out[0:4, 0:8].tolist())
        print("   ref[0:4,0:8]", ref[0:4, 0:8].tolist())
    return ok

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    allok = True
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
        for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 512), (200, 48, 768), (1552, 192, 448)]:
            allok &= run(M, N, K, a_mn, b_mn, "int")
    allok &= run(6208, 2304, 7168, False, False, "rand")
    print("ALL OK" if allok else "SOME FAILED")
