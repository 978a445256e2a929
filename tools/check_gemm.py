"""Standalone validation of the tcgen05 GEMM kernels against torch (run under `timeout` on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskplanner_b200 import _cabi
from maskplanner_b200._cabi import check, ptr, stream_ptr
lib = _cabi.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)

def gemm(A, B, out_fp32):
    M, K = A.shape; N = B.shape[0]
    C = torch.empty(M, N, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=dev)
    check(lib.mpb_gemm_bf16_tn(ptr(A), ptr(B), ptr(C), M, N, K, int(out_fp32), stream_ptr()), "gemm")
    return C

def wgrad(dZ, A):
    M, N = dZ.shape; K = A.shape[1]
    dW = torch.zeros(N, K, dtype=torch.float32, device=dev)
    check(lib.mpb_gemm_bf16_wgrad(ptr(dZ), ptr(A), ptr(dW), M, N, K, stream_ptr()), "wgrad")
    return dW

ok = True
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fwd"):
    for (M, N, K) in [(128, 64, 64), (256, 64, 128), (1000, 128, 192), (4096, 256, 128), (8192, 512, 320), (8192, 1024, 512), (300, 160, 64), (70000, 64, 64), (5, 32, 64)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
        want = A.float() @ B.float().t()
        for f32 in (True, False):
            got = gemm(A, B, f32).float()
            torch.cuda.synchronize()
            err = float((got - want).abs().max() / want.abs().max())
            tol = 1e-5 if f32 else 1e-2
            flag = err < tol
            ok &= flag
            print("fwd M=%d N=%d K=%d f32=%d relerr=%.3e %s" % (M, N, K, f32, err, "ok" if flag else "FAIL"), flush=True)
if which in ("all", "wgrad"):
    for (M, N, K) in [(64, 128, 64), (128, 64, 64), (1024, 64, 64), (5000, 128, 192), (100000, 256, 128), (8192, 1024, 512), (8192, 256, 320), (777, 64, 64)]:
        dZ = torch.randn(M, N, device=dev).bfloat16(); A = torch.randn(M, K, device=dev).bfloat16()
        want = dZ.float().t() @ A.float()
        got = wgrad(dZ, A)
        torch.cuda.synchronize()
        err = float((got - want).abs().max() / want.abs().max())
        flag = err < 1e-4
        ok &= flag
        print("wgrad M=%d N=%d K=%d relerr=%.3e %s" % (M, N, K, err, "ok" if flag else "FAIL"), flush=True)
if which in ("all", "perf") and ok:
    def bench(fn, n=10):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    for (M, N, K) in [(1 << 20, 64, 64), (1 << 20, 128, 64), (1 << 19, 128, 192), (1 << 19, 128, 128), (1 << 19, 256, 128), (8192, 1024, 512)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); dZ = torch.randn(M, N, device=dev).bfloat16()
        t = bench(lambda: gemm(A, B, False))
        byts = (M * K + M * N) * 2
        print("perf fwd  M=%d N=%d K=%d  %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (M, N, K, t, 2 * M * N * K / t / 1e9, byts / t / 1e6), flush=True)
        t = bench(lambda: wgrad(dZ, A))
        print("perf wgrad M=%d N=%d K=%d  %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (M, N, K, t, 2 * M * N * K / t / 1e9, byts / t / 1e6), flush=True)
        t = bench(lambda: torch.matmul(A, B.t()))
        print("perf cublas bf16 fwd            %.3f ms" % t, flush=True)
print("ALL OK" if ok else "SOME FAILED")
