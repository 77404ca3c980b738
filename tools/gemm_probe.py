"""Time rr_gemm_bf16 on the Llama-3-8B shapes (CUDA events, L2 flushed between iterations by
cycling over weight copies larger than L2). Prints achieved GB/s (decode) and TFLOP/s (prefill)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rr_b200 import _lib

lib = _lib.lib


def time_gemm(rowsA, rowsB, K, mode, bn, splits, iters=20, ncopies=4):
    As = [torch.randn(rowsA, K, device="cuda").bfloat16() for _ in range(ncopies)]
    B = torch.randn(rowsB, K, device="cuda").bfloat16()
    if mode == 0:
        out = torch.empty(rowsA, rowsB, device="cuda", dtype=torch.bfloat16); ldo, ldr = rowsB, 0
    else:
        planes = splits
        out = torch.zeros(planes, rowsB, rowsA, device="cuda", dtype=torch.float32); ldo, ldr = rowsA, rowsB
    def run(i):
        A = As[i % ncopies]
        rc = lib.rr_gemm_bf16(A.data_ptr(), rowsA, K, B.data_ptr(), rowsB, K, K, out.data_ptr(),
                              ldo, ldr, splits, mode, bn, None)
        assert rc == 0, rc
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): run(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    print("decode orientation (weights streamed): rowsA=N_out rowsB=batch")
    for name, N, K, splits_list in [("qkv", 6144, 4096, [3]), ("o", 4096, 4096, [4]),
                                    ("gate_up", 28672, 4096, [1]), ("down", 4096, 14336, [4]),
                                    ("lm_head", 128256, 4096, [1])]:
        for s in splits_list:
            for bn in (64,):
                ms = time_gemm(N, 64, K, 1, bn, s, ncopies=max(2, int(300e6 / (N * K * 2)) + 1))
                gb = N * K * 2 / ms / 1e6
                print(f"  {name:8s} N={N:6d} K={K:5d} bn={bn} splits={s:2d}: {ms*1e3:8.1f} us  {gb:7.0f} GB/s")
    if "--decode-only" in sys.argv: sys.exit(0)
    print("prefill orientation: rowsA=tokens rowsB=N_out")
    for name, T, N, K in [("qkv", 8192, 6144, 4096), ("gate_up", 8192, 28672, 4096), ("down", 8192, 4096, 14336),
                          ("qkv2k", 2048, 6144, 4096)]:
        for bn in (128, 256):
            ms = time_gemm(T, N, K, 0, bn, 1, iters=5, ncopies=1)
            tf = 2.0 * T * N * K / ms / 1e9
            print(f"  {name:8s} T={T} N={N} K={K} bn={bn}: {ms:8.3f} ms  {tf:7.0f} TFLOP/s")
