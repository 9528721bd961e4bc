"""Depthwise micro-benchmark (GPU box): yamb_depthwise_fwd / _bwd at the bench shapes."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yet_another_mobilenet_series_b200 import native as nat  # noqa: E402


def run(tag, N, H, Cc, k, s, iters=10):
    lib = nat.lib()
    dev = "cuda"
    bf = torch.bfloat16
    Ho = (H - 1) // s + 1
    x = torch.randn(N, H, H, Cc, device=dev).to(bf)
    y = torch.zeros(N, Ho, Ho, Cc, device=dev, dtype=bf)
    w = torch.randn(Cc, 1, k, k, device=dev)
    sc, sh = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    nct = lib.yamb_max_ctas()
    part = torch.zeros(nct * 2 * Cc, device=dev)
    cnt = torch.zeros(1, device=dev, dtype=torch.int32)
    outs = [torch.zeros(Cc, device=dev) for _ in range(4)]
    f = nat.BnFwd()
    f.partials, f.counter, f.eps, f.momentum = part.data_ptr(), cnt.data_ptr(), 1e-3, 0.01
    f.scale, f.shift, f.mean, f.invstd = [o.data_ptr() for o in outs]
    f.count = N * Ho * Ho
    d = nat.DwFwd()
    d.N, d.H, d.W, d.C, d.ldc, d.k, d.stride = N, H, H, Cc, Cc, k, s
    d.x, d.in_scale, d.in_shift, d.in_act = x.data_ptr(), sc.data_ptr(), sh.data_ptr(), 1
    d.w, d.y, d.bn = w.data_ptr(), y.data_ptr(), C.pointer(f)
    st = nat.stream_handle()

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms = timeit(lambda: nat.check(lib.yamb_depthwise_fwd(C.byref(d), st)))
    nb = 2 * Cc * N * (H * H + Ho * Ho)
    print("%-18s fwd N=%d H=%3d C=%4d k=%d s=%d  %.3f ms %7.1f GB/s" % (tag, N, H, Cc, k, s, ms,
                                                                       nb / ms / 1e6))
    # backward
    dz = torch.randn(N, Ho, Ho, Cc, device=dev).to(bf)
    h = torch.randn(N, Ho, Ho, Cc, device=dev).to(bf)
    dx = torch.zeros(N, H, H, Cc, device=dev, dtype=bf)
    dw = torch.zeros_like(w)
    co = [torch.ones(Cc, device=dev) for _ in range(3)]
    b = nat.BnBwd()
    b.partials, b.counter = part.data_ptr(), cnt.data_ptr()
    b.mean, b.invstd = outs[2].data_ptr(), outs[3].data_ptr()
    oc = [torch.zeros(Cc, device=dev) for _ in range(5)]
    b.dgamma, b.dbeta, b.ca, b.cb, b.cc = [o.data_ptr() for o in oc]
    b.count, b.use_batch_stats = N * H * H, 1
    e = nat.DwBwd()
    e.N, e.H, e.W, e.C, e.ldc, e.k, e.stride = N, H, H, Cc, Cc, k, s
    e.dz, e.h = dz.data_ptr(), h.data_ptr()
    e.ca, e.cb, e.cc = [o.data_ptr() for o in co]
    e.w, e.dw, e.x = w.data_ptr(), dw.data_ptr(), x.data_ptr()
    e.in_scale, e.in_shift, e.in_act = sc.data_ptr(), sh.data_ptr(), 1
    e.dx, e.bn = dx.data_ptr(), C.pointer(b)
    ms = timeit(lambda: nat.check(lib.yamb_depthwise_bwd(C.byref(e), st)))
    nb = 2 * Cc * N * (2 * H * H + 2 * Ho * Ho)
    print("%-18s bwd N=%d H=%3d C=%4d k=%d s=%d  %.3f ms %7.1f GB/s" % (tag, N, H, Cc, k, s, ms,
                                                                       nb / ms / 1e6))
    sys.stdout.flush()


def main():
    N = 256
    run("b2 (112->56)", N, 112, 96, 3, 2)
    run("b3 (56)", N, 56, 144, 3, 1)
    run("b4 (56->28)", N, 56, 144, 3, 2)
    run("b6 (28)", N, 28, 192, 3, 1)
    run("b9 (14)", N, 14, 384, 3, 1)
    run("b13 (14)", N, 14, 576, 3, 1)
    run("b16 (7)", N, 7, 960, 3, 1)
    run("k5 (28)", N, 28, 120, 5, 1)
    run("k7 (14)", N, 14, 240, 7, 2)


if __name__ == "__main__":
    main()
