"""Standalone validation of the tcgen05 GEMM kernels (mpb_sa_gemm_tn / mpb_sa_gemm_wgrad) against float64 torch, every
dtype x operand-transform x epilogue combination; prints one line per case (run under `timeout` on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maskplanner_b200 import _cabi
from maskplanner_b200._cabi import check, ptr, stream_ptr

lib = _cabi.load()
dev = torch.device("cuda", 0)
NAMES = {0: "bf16", 1: "tf32", 2: "tf32x3"}
TOL = {0: 1.5e-2, 1: 2e-3, 2: 2e-5}


def split_hi_lo(w):
    """hi = cvt.rna.tf32(w) (round to nearest, ties away), lo = w - hi."""
    bits = w.contiguous().view(torch.int32)
    hi = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, (w - hi)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))


def run_tn(dt, M, N, K, xform, epi, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + M + N + K)
    tdt = torch.bfloat16 if dt == 0 else torch.float32
    A = torch.randn(M, K, device=dev, generator=g).to(tdt)
    B = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(tdt)
    B_hi, B_lo = (B, None) if dt != 2 else split_hi_lo(B)
    C = torch.full((M, N), float("nan"), dtype=tdt, device=dev)
    sc = (torch.rand(K, device=dev, generator=g) + 0.5) if xform else None
    sh = (torch.randn(K, device=dev, generator=g) * 0.3) if xform else None
    Z = torch.randn(M, N, device=dev, generator=g).to(tdt) if epi == 2 else None
    zs = (torch.rand(N, device=dev, generator=g) + 0.5) if epi == 2 else None
    zh = (torch.randn(N, device=dev, generator=g) * 0.3) if epi == 2 else None
    nparts = lib.mpb_sa_gemm_stat_partials(dt, M, N, K, int(xform), epi) if epi else 0
    if epi and not nparts:
        return "skip (cannot fuse)"
    part = torch.full((max(nparts, 1), 2, N), float("nan"), device=dev) if epi else None
    check(lib.mpb_sa_gemm_tn(dt, ptr(A), ptr(B_hi), ptr(B_lo), ptr(C), M, N, K, ptr(sc), ptr(sh), epi, ptr(part), nparts, ptr(Z),
                             ptr(zs), ptr(zh), stream_ptr()), "gemm_tn")
    torch.cuda.synchronize()
    Af = A.double()
    if xform:
        Af = torch.relu(Af * sc.double() + sh.double())
        if dt == 0:
            Af = Af.float().bfloat16().double()
    want = Af @ B.double().t()
    err = rel(C, want)
    msg = "C rel %.2e" % err
    ok = err < TOL[dt]
    if epi == 1:
        Cs = C.double()
        e0, e1 = rel(part[:, 0].double().sum(0), Cs.sum(0)), rel(part[:, 1].double().sum(0), (Cs * Cs).sum(0))
        msg += " sum %.1e sumsq %.1e" % (e0, e1)
        ok = ok and e0 < 1e-4 and e1 < 1e-4
    if epi == 2:
        Cs, Zd = C.double(), Z.double()
        dy = torch.where(Zd.float() * zs + zh > 0, Cs, torch.zeros_like(Cs))
        e0, e1 = rel(part[:, 0].double().sum(0), dy.sum(0)), rel(part[:, 1].double().sum(0), (dy * Zd).sum(0))
        msg += " sum_dy %.1e sum_dyz %.1e" % (e0, e1)
        ok = ok and e0 < 1e-3 and e1 < 1e-3
    return ("ok   " if ok else "FAIL ") + msg


def run_wg(dt, M, N, K, xform, cout=None, cin=None, xyz_last=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + M + N + K)
    tdt = torch.bfloat16 if dt == 0 else torch.float32
    dZ = torch.randn(M, N, device=dev, generator=g).to(tdt)
    A = torch.randn(M, K, device=dev, generator=g).to(tdt)
    sc = (torch.rand(K, device=dev, generator=g) + 0.5) if xform else None
    sh = (torch.randn(K, device=dev, generator=g) * 0.3) if xform else None
    cout, cin = cout or N, cin or K
    ws_bytes = lib.mpb_sa_gemm_wgrad_workspace(dt, M, N, K, int(xform))
    if ws_bytes < 0:
        return "skip (unsupported)"
    ws = torch.empty(ws_bytes // 4, device=dev)
    dW = torch.full((cout, cin), float("nan"), device=dev)
    check(lib.mpb_sa_gemm_wgrad(dt, ptr(dZ), ptr(A), M, N, K, ptr(sc), ptr(sh), ptr(ws), cout, cin, int(xyz_last), 0, ptr(dW), stream_ptr()),
          "wgrad")
    torch.cuda.synchronize()
    Af = A.double()
    if xform:
        Af = torch.relu(Af * sc.double() + sh.double())
        if dt == 0:
            Af = Af.float().bfloat16().double()
    want = (dZ.double().t() @ Af)[:cout]
    if xyz_last and cin > 3:
        want = torch.cat([want[:, cin - 3:cin], want[:, :cin - 3]], dim=1)
    else:
        want = want[:, :cin]
    err = rel(dW, want)
    # determinism: a second run must be bit-identical
    dW2 = torch.empty_like(dW)
    check(lib.mpb_sa_gemm_wgrad(dt, ptr(dZ), ptr(A), M, N, K, ptr(sc), ptr(sh), ptr(ws), cout, cin, int(xyz_last), 0, ptr(dW2), stream_ptr()),
          "wgrad")
    torch.cuda.synchronize()
    same = bool(torch.equal(dW, dW2))
    ok = err < {0: 2e-3, 1: 2e-3, 2: 2e-5}[dt] and same
    return ("ok   " if ok else "FAIL ") + "dW rel %.2e bitwise-repeatable %s" % (err, same)


def _pool_inputs(M, C, Kg, g):
    """Random stored pre-activation Zl [M,C] (bf16), arg-max rows, pgo and (-w, e), plus the float64 dZ they define."""
    Zl = torch.randn(M, C, device=dev, generator=g).bfloat16()
    G = M // Kg
    arg = torch.randint(0, Kg, (G, C), device=dev, generator=g, dtype=torch.int32)
    pgo = torch.randn(G, C, device=dev, generator=g) * (torch.rand(G, C, device=dev, generator=g) > 0.2)
    negw_e = torch.stack([torch.randn(C, device=dev, generator=g) * 0.5, torch.randn(C, device=dev, generator=g) * 0.3]).contiguous()
    k = (torch.arange(M, device=dev) % Kg).view(G, Kg, 1)
    sparse = torch.where(k == arg.view(G, 1, C).long(), pgo.view(G, 1, C).double(), torch.zeros((), device=dev, dtype=torch.float64)).view(M, C)
    dz = sparse + negw_e[0].double() * Zl.double() + negw_e[1].double()
    return Zl, arg, pgo, negw_e, dz.float().bfloat16().double()      # the kernel rounds the rebuilt tile to bf16


def run_tn_pool(M, N, K, Kg, epi, seed=0):
    """mpb_sa_gemm_tn_pool: C = dZ @ B^T with dZ rebuilt from (Zl, argmax, pgo, negw_e) in shared memory."""
    g = torch.Generator(device="cuda").manual_seed(seed + M + N + K)
    Zl, arg, pgo, negw_e, dz = _pool_inputs(M, K, Kg, g)
    B = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
    C = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    Z = torch.randn(M, N, device=dev, generator=g).bfloat16() if epi == 2 else None
    zs = (torch.rand(N, device=dev, generator=g) + 0.5) if epi == 2 else None
    zh = (torch.randn(N, device=dev, generator=g) * 0.3) if epi == 2 else None
    nparts = lib.mpb_sa_gemm_stat_partials(0, M, N, K, 2, epi) if epi else 0
    if epi and not nparts:
        return "skip (cannot fuse)"
    part = torch.full((max(nparts, 1), 2, N), float("nan"), device=dev) if epi else None
    check(lib.mpb_sa_gemm_tn_pool(0, ptr(Zl), ptr(B), ptr(C), M, N, K, Kg, ptr(arg), ptr(pgo), ptr(negw_e), epi, ptr(part), nparts, ptr(Z),
                                  ptr(zs), ptr(zh), stream_ptr()), "gemm_tn_pool")
    torch.cuda.synchronize()
    want = dz @ B.double().t()
    err = rel(C, want)
    msg = "C rel %.2e" % err
    ok = err < TOL[0]
    if epi == 2:
        Cs, Zd = C.double(), Z.double()
        dy = torch.where(Zd.float() * zs + zh > 0, Cs, torch.zeros_like(Cs))
        e0, e1 = rel(part[:, 0].double().sum(0), dy.sum(0)), rel(part[:, 1].double().sum(0), (dy * Zd).sum(0))
        msg += " sum_dy %.1e sum_dyz %.1e" % (e0, e1)
        ok = ok and e0 < 1e-3 and e1 < 1e-3
    return ("ok   " if ok else "FAIL ") + msg


def run_wg_pool(M, N, K, Kg, xform, seed=0):
    """mpb_sa_gemm_wgrad_pool: dW = dZ^T @ f(A) with dZ rebuilt from (Zl, argmax, pgo, negw_e)."""
    g = torch.Generator(device="cuda").manual_seed(seed + M + N + K)
    Zl, arg, pgo, negw_e, dz = _pool_inputs(M, N, Kg, g)
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    sc = (torch.rand(K, device=dev, generator=g) + 0.5) if xform else None
    sh = (torch.randn(K, device=dev, generator=g) * 0.3) if xform else None
    ws = torch.empty(lib.mpb_sa_gemm_wgrad_workspace(0, M, N, K, int(xform)) // 4, device=dev)
    outs = []
    for _ in range(2):
        dW = torch.full((N, K), float("nan"), device=dev)
        check(lib.mpb_sa_gemm_wgrad_pool(0, ptr(Zl), ptr(A), M, N, K, ptr(sc), ptr(sh), Kg, ptr(arg), ptr(pgo), ptr(negw_e), ptr(ws), N, K, 0, 0,
                                         ptr(dW), stream_ptr()), "wgrad_pool")
        torch.cuda.synchronize()
        outs.append(dW)
    Af = A.double()
    if xform:
        Af = torch.relu(Af * sc.double() + sh.double()).float().bfloat16().double()
    err = rel(outs[0], dz.t() @ Af)
    same = bool(torch.equal(outs[0], outs[1]))
    return ("ok   " if err < 2e-3 and same else "FAIL ") + "dW rel %.2e bitwise-repeatable %s" % (err, same)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    dts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
    fails = 0
    if which in ("all", "tn"):
        for dt in dts:
            for (M, N, K) in [(128, 64, 64), (1000, 128, 192), (4096, 256, 128), (8192, 512, 320), (8192, 1024, 512), (300, 160, 64),
                              (70000, 64, 64), (5, 32, 64), (33000, 128, 128)]:
                for xform in (False, True):
                    for epi in (0, 1, 2):
                        if xform and epi == 2:
                            continue
                        try:
                            r = run_tn(dt, M, N, K, xform, epi)
                        except Exception as e:
                            r = "FAIL exception %s" % e
                        fails += r.startswith("FAIL")
                        print("tn    %-6s M=%-6d N=%-4d K=%-4d xform=%d epi=%d  %s" % (NAMES[dt], M, N, K, xform, epi, r), flush=True)
    if which in ("all", "wg"):
        for dt in dts:
            for (M, N, K) in [(64, 128, 64), (1024, 64, 64), (5000, 128, 192), (100000, 256, 128), (8192, 1024, 512), (8192, 256, 320), (777, 64, 64)]:
                for xform in (False, True):
                    try:
                        r = run_wg(dt, M, N, K, xform)
                    except Exception as e:
                        r = "FAIL exception %s" % e
                    fails += r.startswith("FAIL")
                    print("wgrad %-6s M=%-6d N=%-4d K=%-4d xform=%d        %s" % (NAMES[dt], M, N, K, xform, r), flush=True)
            r = run_wg(dt, 4096, 128, 192, True, cout=100, cin=131, xyz_last=True)
            fails += r.startswith("FAIL")
            print("wgrad %-6s crop 100x131 xyz_last                  %s" % (NAMES[dt], r), flush=True)
    if which in ("all", "pool"):
        for (M, N, K, Kg) in [(4096, 64, 128, 32), (8192, 128, 256, 64), (70016, 64, 128, 32), (1024, 320, 256, 128), (1024, 64, 128, 256), (128, 64, 64, 16)]:
            for epi in (0, 2):
                r = run_tn_pool(M, N, K, Kg, epi)
                fails += r.startswith("FAIL")
                print("tn_pool    M=%-6d N=%-4d K=%-4d Kg=%-3d epi=%d  %s" % (M, N, K, Kg, epi, r), flush=True)
        for (M, N, K, Kg) in [(4096, 128, 64, 32), (8192, 256, 128, 64), (70016, 128, 64, 32), (1024, 64, 192, 128), (1024, 128, 64, 256), (64, 64, 64, 16)]:
            for xform in (False, True):
                r = run_wg_pool(M, N, K, Kg, xform)
                fails += r.startswith("FAIL")
                print("wgrad_pool M=%-6d N=%-4d K=%-4d Kg=%-3d xform=%d %s" % (M, N, K, Kg, xform, r), flush=True)
    print("FAILURES: %d" % fails)
    sys.exit(1 if fails else 0)
