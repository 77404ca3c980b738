"""Prefill GEMM epilogue cost: plain bf16 epilogue (mode 0) vs residual-add epilogue (mode 5: fp32 read-modify-write)
on the O-projection and down-projection shapes of one 8192-token chunk, alternating in one process."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rr_b200 import _lib
lib = _lib.lib
def bench(T, N, K, iters=6):
    A = torch.randn(T, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    o16 = torch.empty(T, N, device="cuda", dtype=torch.bfloat16)
    x32 = torch.zeros(T, N, device="cuda", dtype=torch.float32)
    f0 = lambda: lib.rr_gemm_bf16(A.data_ptr(), T, K, B.data_ptr(), N, K, K, o16.data_ptr(), N, 0, 1, 0, 256, None)
    f5 = lambda: lib.rr_gemm_bf16(A.data_ptr(), T, K, B.data_ptr(), N, K, K, x32.data_ptr(), N, 0, 1, 5, 256, None)
    res = {0: [], 5: []}
    for f in (f0, f5):
        for _ in range(2): assert f() == 0
    torch.cuda.synchronize()
    for _ in range(iters):
        for m, f in ((0, f0), (5, f5)):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
            res[m].append(e0.elapsed_time(e1) / 2 * 1e3)
    return {m: sorted(v)[len(v) // 2] for m, v in res.items()}
for name, T, N, K in [("o", 8192, 4096, 4096), ("down", 8192, 4096, 14336), ("qkv-like", 8192, 6144, 4096)]:
    r = bench(T, N, K)
    fl = 2.0 * T * N * K
    print(f"{name:9s} bf16 epilogue {r[0]:7.1f} us ({fl / r[0] / 1e6:5.0f} TF)   residual epilogue {r[5]:7.1f} us ({fl / r[5] / 1e6:5.0f} TF)   +{r[5] - r[0]:.1f} us")
