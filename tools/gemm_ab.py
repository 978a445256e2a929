"""A/B timing of the shared-MLP GEMM kernels at the training step's shapes: every (shape, xform, epi) variant timed alone
with CUDA events, L2 flushed between iterations, median of 9.  Developer tool (GPU box).  argv[1]: 'tn', 'wg' or 'all'."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maskplanner_b200 import _cabi
from maskplanner_b200._cabi import check, ptr, stream_ptr

lib = _cabi.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def med(fn, n=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def tn(M, N, K, xform, epi, dt=0):
    tdt = torch.bfloat16 if dt == 0 else torch.float32
    esz = 2 if dt == 0 else 4
    A = torch.randn(M, K, device=dev).to(tdt)
    B = (torch.randn(N, K, device=dev) / K ** 0.5).to(tdt)
    C = torch.empty(M, N, dtype=tdt, device=dev)
    sc = torch.rand(K, device=dev) + 0.5 if xform else None
    sh = torch.randn(K, device=dev) * 0.3 if xform else None
    Z = torch.randn(M, N, device=dev).to(tdt) if epi == 2 else None
    zs = torch.rand(N, device=dev) + 0.5 if epi == 2 else None
    zh = torch.randn(N, device=dev) * 0.3 if epi == 2 else None
    nparts = lib.mpb_sa_gemm_stat_partials(dt, M, N, K, int(xform), epi) if epi else 0
    part = torch.empty(max(nparts, 1), 2, N, device=dev) if epi else None

    def fn():
        check(lib.mpb_sa_gemm_tn(dt, ptr(A), ptr(B), ptr(B if dt == 2 else None), ptr(C), M, N, K, ptr(sc), ptr(sh), epi, ptr(part), nparts, ptr(Z),
                                 ptr(zs), ptr(zh), stream_ptr()), "gemm_tn")
    ms = med(fn)
    nbytes = esz * (M * K + N * K + M * N + (M * N if epi == 2 else 0))
    print("tn    M=%-8d N=%-4d K=%-4d xform=%d epi=%d  %7.1f us %6.0f GB/s %6.1f TFLOP/s" % (M, N, K, xform, epi, ms * 1e3, nbytes / ms / 1e6,
                                                                                           2.0 * M * N * K / ms / 1e9), flush=True)


def wg(M, N, K, xform, dt=0):
    tdt = torch.bfloat16 if dt == 0 else torch.float32
    esz = 2 if dt == 0 else 4
    dZ = torch.randn(M, N, device=dev).to(tdt)
    A = torch.randn(M, K, device=dev).to(tdt)
    sc = torch.rand(K, device=dev) + 0.5 if xform else None
    sh = torch.randn(K, device=dev) * 0.3 if xform else None
    ws = torch.empty(lib.mpb_sa_gemm_wgrad_workspace(dt, M, N, K, int(xform)) // 4, device=dev)
    dW = torch.empty(N, K, device=dev)

    def fn():
        check(lib.mpb_sa_gemm_wgrad(dt, ptr(dZ), ptr(A), M, N, K, ptr(sc), ptr(sh), ptr(ws), N, K, 0, 0, ptr(dW), stream_ptr()), "wgrad")
    ms = med(fn)
    nbytes = esz * M * (N + K) + 4 * N * K
    print("wgrad M=%-8d N=%-4d K=%-4d xform=%d        %7.1f us %6.0f GB/s %6.1f TFLOP/s" % (M, N, K, xform, ms * 1e3, nbytes / ms / 1e6,
                                                                                           2.0 * M * N * K / ms / 1e9), flush=True)


def tn_pool(M, N, K, Kg, epi):
    Zl = torch.randn(M, K, device=dev).bfloat16()
    G = M // Kg
    arg = torch.randint(0, Kg, (G, K), device=dev, dtype=torch.int32)
    pgo = torch.randn(G, K, device=dev)
    negw_e = torch.randn(2, K, device=dev)
    B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    C = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    Z = torch.randn(M, N, device=dev).bfloat16() if epi == 2 else None
    zs = torch.rand(N, device=dev) + 0.5 if epi == 2 else None
    zh = torch.randn(N, device=dev) * 0.3 if epi == 2 else None
    nparts = lib.mpb_sa_gemm_stat_partials(0, M, N, K, 2, epi) if epi else 0
    part = torch.empty(max(nparts, 1), 2, N, device=dev) if epi else None

    def fn():
        check(lib.mpb_sa_gemm_tn_pool(0, ptr(Zl), ptr(B), ptr(C), M, N, K, Kg, ptr(arg), ptr(pgo), ptr(negw_e), epi, ptr(part), nparts, ptr(Z),
                                      ptr(zs), ptr(zh), stream_ptr()), "gemm_tn_pool")
    ms = med(fn)
    nbytes = 2 * (M * K + N * K + M * N + (M * N if epi == 2 else 0))
    print("tn_pl M=%-8d N=%-4d K=%-4d Kg=%-3d   epi=%d  %7.1f us %6.0f GB/s" % (M, N, K, Kg, epi, ms * 1e3, nbytes / ms / 1e6), flush=True)


def wg_pool(M, N, K, Kg, xform):
    Zl = torch.randn(M, N, device=dev).bfloat16()
    G = M // Kg
    arg = torch.randint(0, Kg, (G, N), device=dev, dtype=torch.int32)
    pgo = torch.randn(G, N, device=dev)
    negw_e = torch.randn(2, N, device=dev)
    A = torch.randn(M, K, device=dev).bfloat16()
    sc = torch.rand(K, device=dev) + 0.5 if xform else None
    sh = torch.randn(K, device=dev) * 0.3 if xform else None
    ws = torch.empty(lib.mpb_sa_gemm_wgrad_workspace(0, M, N, K, int(xform)) // 4, device=dev)
    dW = torch.empty(N, K, device=dev)

    def fn():
        check(lib.mpb_sa_gemm_wgrad_pool(0, ptr(Zl), ptr(A), M, N, K, ptr(sc), ptr(sh), Kg, ptr(arg), ptr(pgo), ptr(negw_e), ptr(ws), N, K, 0, 0,
                                         ptr(dW), stream_ptr()), "wgrad_pool")
    ms = med(fn)
    nbytes = 2 * M * (N + K) + 4 * N * K
    print("wg_pl M=%-8d N=%-4d K=%-4d Kg=%-3d xform=%d  %7.1f us %6.0f GB/s" % (M, N, K, Kg, xform, ms * 1e3, nbytes / ms / 1e6), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "pooldbg":
    for dbg in (0, 1, 2, 3, 4, 6, 7):
        os.environ["MPB_POOL_DBG"] = str(dbg)
        print("MPB_POOL_DBG=%d" % dbg)
        tn_pool(1048576, 64, 128, 32, 0)
        tn_pool(524288, 128, 256, 64, 0)
if which == "pool":
    for (M, N, K, Kg) in [(1048576, 64, 128, 32), (524288, 128, 256, 64)]:
        for epi in (0, 2):
            tn(M, N, K, 0, epi)
            tn_pool(M, N, K, Kg, epi)
    tn(1048576, 64, 128, 1, 0)
    tn(524288, 128, 256, 1, 0)
    for (M, N, K, Kg) in [(1048576, 128, 64, 32), (524288, 256, 128, 64)]:
        for xform in (0, 1):
            wg(M, N, K, xform)
            wg_pool(M, N, K, Kg, xform)
if which in ("tn", "all"):
    for (M, N, K) in [(1048576, 64, 64), (1048576, 128, 64), (524288, 128, 128), (524288, 256, 128), (524288, 128, 192)]:
        for xform in (0, 1):
            for epi in (0, 1):
                tn(M, N, K, xform, epi)
    for (M, N, K) in [(1048576, 64, 128), (1048576, 64, 64), (524288, 128, 256), (524288, 128, 128), (524288, 192, 128)]:
        for epi in (0, 2):
            tn(M, N, K, 0, epi)
    for (M, N, K) in [(8192, 256, 320), (8192, 512, 256), (8192, 1024, 512)]:
        for xform in (0, 1):
            for epi in (0, 1):
                tn(M, N, K, xform, epi)
    for (M, N, K) in [(8192, 512, 1024), (8192, 256, 512), (8192, 320, 256)]:
        for epi in (0, 2):
            tn(M, N, K, 0, epi)
if which in ("wg", "all"):
    for (M, N, K) in [(1048576, 128, 64), (1048576, 64, 64), (524288, 256, 128), (524288, 128, 128), (524288, 128, 192),
                      (8192, 1024, 512), (8192, 512, 256), (8192, 256, 320)]:
        for xform in (0, 1):
            wg(M, N, K, xform)
