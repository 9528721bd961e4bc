"""GEMM micro-benchmark (GPU box): time yamb_pointwise_gemm on bench shapes with features toggled."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yet_another_mobilenet_series_b200 import native as nat  # noqa: E402


def run(tag, M, N, K, stats=False, xform=0, a_mn=0, b_mn=0, epi=0, iters=10, dgrad=False,
        wgrad=False):
    lib = nat.lib()
    dev = "cuda"
    bf = torch.bfloat16
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(bf)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(bf)
    D = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == 2 else bf)
    g = nat.Gemm()
    g.M, g.N, g.K = M, N, K
    g.a_mn_major, g.b_mn_major = a_mn, b_mn
    g.A, g.lda = A.data_ptr(), A.stride(0)
    g.B, g.ldb = B.data_ptr(), B.stride(0)
    g.D, g.ldd = D.data_ptr(), N
    g.epi = epi
    keep = []
    if xform:
        Cdim = M if a_mn else K
        sc, sh = torch.ones(Cdim, device=dev), torch.zeros(Cdim, device=dev)
        keep += [sc, sh]
        g.a_xform, g.a_act = 1, 1
        g.a_scale, g.a_shift = sc.data_ptr(), sh.data_ptr()
    if wgrad:
        # project-wgrad shape: A = ca*dy + cb*h3 + cc (two sources, M channels, MN-major),
        # B = act(s*h2 + t) (N channels, MN-major), split-K atomic epilogue
        A2 = torch.randn(K, M, device=dev).to(bf)
        co = [torch.ones(M, device=dev), torch.zeros(M, device=dev), torch.ones(M, device=dev) * 0.1]
        cb2 = [torch.ones(N, device=dev), torch.zeros(N, device=dev)]
        g.a_xform = 2
        g.a_scale, g.a_shift, g.a_scale2 = co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr()
        g.A2, g.lda2 = A2.data_ptr(), M
        g.b_xform, g.b_act = 1, 2
        g.b_scale, g.b_shift = cb2[0].data_ptr(), cb2[1].data_ptr()
        keep += [A2] + co + cb2
    if dgrad:
        # project-dgrad shape: A = ca*dy + cb*h3 + cc (two sources, K channels), B = W3 MN-major,
        # epilogue dz = acc * act'(s*h2+t) + BatchNorm-backward statistics
        nct = lib.yamb_max_ctas()
        A2 = torch.randn(M, K, device=dev).to(bf)
        co = [torch.ones(K, device=dev), torch.zeros(K, device=dev), torch.ones(K, device=dev) * 0.1]
        g.a_xform = 2
        g.a_scale, g.a_shift, g.a_scale2 = co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr()
        g.A2, g.lda2 = A2.data_ptr(), K
        H = torch.randn(M, N, device=dev).to(bf)
        hs, ht = torch.ones(N, device=dev), torch.zeros(N, device=dev)
        mean, invstd, gamma = torch.zeros(N, device=dev), torch.ones(N, device=dev), torch.ones(N, device=dev)
        bw = nat.BnBwd()
        parts, cnt = torch.zeros(nct * 2 * N, device=dev), torch.zeros(1, device=dev, dtype=torch.int32)
        dg, db = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        cs = [torch.zeros(N, device=dev) for _ in range(3)]
        bw.partials, bw.counter = parts.data_ptr(), cnt.data_ptr()
        bw.gamma, bw.mean, bw.invstd = gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr()
        bw.dgamma, bw.dbeta = dg.data_ptr(), db.data_ptr()
        bw.ca, bw.cb, bw.cc = [o.data_ptr() for o in cs]
        bw.count, bw.use_batch_stats = M, 1
        g.epi = 1
        g.H, g.ldh = H.data_ptr(), N
        g.h_scale, g.h_shift, g.h_act = hs.data_ptr(), ht.data_ptr(), 2
        g.bn_bwd = C.pointer(bw)
        keep += [A2, H, hs, ht, mean, invstd, gamma, bw, parts, cnt, dg, db] + co + cs
    if stats:
        nct = lib.yamb_max_ctas()
        f = nat.BnFwd()
        bufs = [torch.zeros(nct * 2 * N, device=dev), torch.zeros(1, device=dev, dtype=torch.int32)]
        outs = [torch.zeros(N, device=dev) for _ in range(4)]
        keep += bufs + outs + [f]
        f.partials, f.counter = bufs[0].data_ptr(), bufs[1].data_ptr()
        f.eps, f.momentum = 1e-3, 0.01
        f.scale, f.shift, f.mean, f.invstd = [o.data_ptr() for o in outs]
        f.count = M
        g.bn_fwd = C.pointer(f)
    st = nat.stream_handle()
    for _ in range(3):
        nat.check(lib.yamb_pointwise_gemm(C.byref(g), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        nat.check(lib.yamb_pointwise_gemm(C.byref(g), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = 2 * (M * K + N * K) + (4 if epi == 2 else 2) * M * N * (0 if epi == 2 else 1)
    if dgrad:
        nbytes += 2 * M * K + 2 * M * N
    if wgrad:
        nbytes += 2 * M * K
    print("%-34s M=%7d N=%4d K=%4d  %.3f ms  %7.1f GB/s" % (tag, M, N, K, ms, nbytes / ms / 1e6))
    sys.stdout.flush()


def main():
    M = 256 * 56 * 56
    run("expand b3 plain", M, 144, 24)
    run("expand b3 +stats", M, 144, 24, stats=True)
    run("project b3 plain", M, 24, 144)
    run("project b3 +stats", M, 24, 144, stats=True)
    run("project b3 +xform", M, 24, 144, xform=1)
    run("project b3 +xform+stats", M, 24, 144, stats=True, xform=1)
    run("dgrad-like b3 (b_mn) plain", M, 144, 24, b_mn=1)
    run("wgrad b3 expand (144x24)", 144, 24, M, a_mn=1, b_mn=1, epi=2)
    run("wgrad b3 +xform", 144, 24, M, a_mn=1, b_mn=1, epi=2, xform=1)
    M2 = 256 * 112 * 112
    run("expand b2 plain", M2, 96, 16)
    run("expand b2 +stats", M2, 96, 16, stats=True)
    M3 = 256 * 14 * 14
    run("expand b12 plain", M3, 576, 96)
    run("expand b12 +stats", M3, 576, 96, stats=True)
    run("project b12 +xform+stats", M3, 96, 576, stats=True, xform=1)
    run("square 8192x1024x1024 plain", 8192, 1024, 1024)
    print("torch copy reference:")
    a = torch.empty(M * 144, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("  copy %d MB: %.3f ms %.1f GB/s" % (a.numel() * 2 / 1e6, ms, 2 * a.numel() * 2 / ms / 1e6))


if __name__ == "__main__":
    main()
