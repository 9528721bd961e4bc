"""GPU probe for yamb_pointwise_gemm: each case runs in its own process (a trapped kernel kills the
CUDA context) and is compared with a plain torch fp32 computation of the same op.

Usage:  python tests/gpu_probe_gemm.py            # run all cases, one subprocess each
        python tests/gpu_probe_gemm.py <case>     # run one case in-process
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _act(z, act):
    import torch
    if act == 1:
        return torch.relu(z)
    if act == 2:
        return torch.clamp(z, 0, 6)
    if act == 3:
        return z * torch.sigmoid(z)
    if act == 4:
        return z * torch.clamp(z + 3, 0, 6) / 6
    return z


def _act_grad(z, act):
    import torch
    if act == 1:
        return (z > 0).float()
    if act == 2:
        return ((z > 0) & (z < 6)).float()
    if act == 3:
        s = torch.sigmoid(z)
        return s * (1 + z * (1 - s))
    if act == 4:
        return torch.where(z <= -3, torch.zeros_like(z),
                           torch.where(z >= 3, torch.ones_like(z), (2 * z + 3) / 6))
    return torch.ones_like(z)


CASES = {
    # name: dict(M,N,K, mode...)
    "plain_small": dict(M=256, N=64, K=64),
    "plain_k16": dict(M=384, N=96, K=16),
    "plain_k24_n24": dict(M=200, N=24, K=24),
    "plain_k144_n32": dict(M=1000, N=32, K=144),
    "plain_multi_n": dict(M=640, N=320, K=960),
    "plain_n960": dict(M=300, N=960, K=160),
    "plain_big": dict(M=128 * 148 * 3 + 40, N=96, K=16),
    "stats": dict(M=1000, N=144, K=24, stats=True),
    "stats_multi_n": dict(M=5000, N=576, K=96, stats=True),
    "xform1": dict(M=1000, N=32, K=192, a_xform=1, act=1, stats=True),
    "xform1_swish": dict(M=777, N=160, K=576, a_xform=1, act=3),
    "dgrad": dict(M=1000, N=144, K=24, b_mn=1),
    "dgrad_big": dict(M=900, N=960, K=160, b_mn=1),
    "residual": dict(M=1000, N=24, K=144, b_mn=1, residual=True),
    "wgrad": dict(M=96, N=16, K=5000, a_mn=1, b_mn=1, epi=2),
    "wgrad_big": dict(M=320, N=960, K=3000, a_mn=1, b_mn=1, epi=2),
    "wgrad_xform": dict(M=24, N=144, K=4000, a_mn=1, b_mn=1, epi=2, b_xform=1, act=2),
    "dz": dict(M=1000, N=144, K=24, b_mn=1, epi=1, act=1),
    "dz_swish_multi": dict(M=700, N=576, K=96, b_mn=1, epi=1, act=3),
    "xform2": dict(M=1000, N=24, K=144, b_mn=1, a_xform=2, residual=True),
    "wgrad_xform2": dict(M=144, N=24, K=4000, a_mn=1, b_mn=1, epi=2, a_xform=2),
}


def run_case(name):
    import torch
    from yet_another_mobilenet_series_b200 import native as nat
    cfg = dict(CASES[name])
    torch.manual_seed(0)
    dev = "cuda"
    M, N, K = cfg["M"], cfg["N"], cfg["K"]
    a_mn, b_mn = cfg.get("a_mn", 0), cfg.get("b_mn", 0)
    epi = cfg.get("epi", 0)
    act = cfg.get("act", 0)
    lib = nat.lib()
    bf = torch.bfloat16
    A = torch.randn(M, K, device=dev).to(bf)      # logical [M,K]
    B = (torch.randn(N, K, device=dev) / (K ** 0.5)).to(bf)  # logical [N,K]
    A_mem = A.t().contiguous() if a_mn else A.contiguous()
    B_mem = B.t().contiguous() if b_mn else B.contiguous()
    g = nat.Gemm()
    g.M, g.N, g.K = M, N, K
    g.a_mn_major, g.b_mn_major = a_mn, b_mn
    g.A, g.lda = A_mem.data_ptr(), A_mem.stride(0)
    g.B, g.ldb = B_mem.data_ptr(), B_mem.stride(0)
    g.epi = epi
    Af, Bf = A.float(), B.float()
    keep = []
    nct = lib.yamb_max_ctas()
    # operand transforms
    for which in ("a", "b"):
        xf = cfg.get(which + "_xform", 0)
        if not xf:
            continue
        Cdim = (M if a_mn else K) if which == "a" else (N if b_mn else K)
        sc = (torch.rand(Cdim, device=dev) + 0.5).float()
        sh = (torch.randn(Cdim, device=dev) * 0.3).float()
        keep += [sc, sh]
        setattr(g, which + "_xform", xf)
        setattr(g, which + "_act", act if xf == 1 else 0)
        setattr(g, which + "_scale", sc.data_ptr())
        setattr(g, which + "_shift", sh.data_ptr())
        X = Af if which == "a" else Bf
        mn = a_mn if which == "a" else b_mn
        # channel dim: K if K-major else the M/N dim
        bshape = (1, -1) if not mn else (-1, 1)
        if xf == 1:
            Xn = _act(X * sc.view(bshape) + sh.view(bshape), act)
        else:
            X2 = torch.randn_like(X).to(bf)
            s2 = (torch.randn(Cdim, device=dev) * 0.5).float()
            X2_mem = X2.t().contiguous() if mn else X2.contiguous()
            keep += [X2_mem, s2]
            setattr(g, which + "_scale2", s2.data_ptr())
            if which == "a":
                g.A2, g.lda2 = X2_mem.data_ptr(), X2_mem.stride(0)
            else:
                g.B2, g.ldb2 = X2_mem.data_ptr(), X2_mem.stride(0)
            Xn = X * sc.view(bshape) + X2.float() * s2.view(bshape) + sh.view(bshape)
        Xn = Xn.to(bf).float()
        if which == "a":
            Af = Xn
        else:
            Bf = Xn
    ref = Af @ Bf.t()
    results = {}
    if epi == 2:
        D = torch.zeros(M, N, device=dev, dtype=torch.float32)
        g.D, g.ldd = D.data_ptr(), N
    else:
        D = torch.full((M, N), float("nan"), device=dev, dtype=bf)
        g.D, g.ldd = D.data_ptr(), N
    if cfg.get("residual"):
        R = torch.randn(M, N, device=dev).to(bf)
        g.residual, g.ldr = R.data_ptr(), N
        ref = ref + R.float()
    fwd = None
    if cfg.get("stats"):
        fwd = nat.BnFwd()
        partials = torch.zeros(nct * 2 * N, device=dev)
        counter = torch.zeros(1, device=dev, dtype=torch.int32)
        gamma = torch.rand(N, device=dev) + 0.5
        beta = torch.randn(N, device=dev)
        rm = torch.zeros(N, device=dev)
        rv = torch.ones(N, device=dev)
        nbt = torch.zeros(1, device=dev, dtype=torch.int64)
        outs = [torch.zeros(N, device=dev) for _ in range(4)]
        keep += [partials, counter, gamma, beta, rm, rv, nbt] + outs
        fwd.partials, fwd.counter = partials.data_ptr(), counter.data_ptr()
        fwd.gamma, fwd.beta = gamma.data_ptr(), beta.data_ptr()
        fwd.eps, fwd.momentum = 1e-3, 0.01
        fwd.running_mean, fwd.running_var = rm.data_ptr(), rv.data_ptr()
        fwd.num_batches_tracked = nbt.data_ptr()
        fwd.scale, fwd.shift, fwd.mean, fwd.invstd = [o.data_ptr() for o in outs]
        fwd.count = M
        g.bn_fwd = C.pointer(fwd)
    bwd = None
    if epi == 1:
        H = torch.randn(M, N, device=dev).to(bf)
        hs = torch.rand(N, device=dev) + 0.5
        ht = torch.randn(N, device=dev) * 0.3
        mean = torch.randn(N, device=dev) * 0.1
        invstd = torch.rand(N, device=dev) + 0.5
        gamma = torch.rand(N, device=dev) + 0.5
        bwd = nat.BnBwd()
        partials = torch.zeros(nct * 2 * N, device=dev)
        counter = torch.zeros(1, device=dev, dtype=torch.int32)
        dg, db = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        co = [torch.zeros(N, device=dev) for _ in range(3)]
        keep += [H, hs, ht, mean, invstd, gamma, partials, counter, dg, db] + co
        bwd.partials, bwd.counter = partials.data_ptr(), counter.data_ptr()
        bwd.gamma, bwd.mean, bwd.invstd = gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr()
        bwd.dgamma, bwd.dbeta = dg.data_ptr(), db.data_ptr()
        bwd.ca, bwd.cb, bwd.cc = [o.data_ptr() for o in co]
        bwd.count = M
        bwd.use_batch_stats = 1
        g.H, g.ldh = H.data_ptr(), N
        g.h_scale, g.h_shift, g.h_act = hs.data_ptr(), ht.data_ptr(), act
        g.bn_bwd = C.pointer(bwd)
        z = H.float() * hs + ht
        ref = ref * _act_grad(z, act)
    torch.cuda.synchronize()
    nat.check(lib.yamb_pointwise_gemm(C.byref(g), nat.stream_handle()))
    torch.cuda.synchronize()
    Df = D.float()
    ok = True

    def report(tag, got, want, tol):
        nonlocal ok
        err = (got - want).norm() / (want.norm() + 1e-20)
        mx = (got - want).abs().max()
        bad = not bool(torch.isfinite(got).all()) or float(err) > tol
        ok = ok and not bad
        print("  %-12s rel_l2=%.3e max_abs=%.3e %s" % (tag, float(err), float(mx),
                                                       "FAIL" if bad else "ok"))

    report("D", Df, ref if epi == 2 else ref.to(bf).float(), 4e-3 if epi != 2 else 2e-3)
    if fwd is not None:
        Db = ref.to(bf).float()
        mean_ref = Db.mean(0)
        var_ref = Db.var(0, unbiased=False)
        invstd_ref = 1 / torch.sqrt(var_ref + 1e-3)
        report("mean", outs[2], mean_ref, 2e-3)
        report("invstd", outs[3], invstd_ref, 2e-3)
        report("scale", outs[0], gamma * invstd_ref, 2e-3)
        report("shift", outs[1], beta - mean_ref * gamma * invstd_ref, 5e-3)
        report("run_mean", rm, 0.01 * mean_ref, 2e-3)
        report("run_var", rv, 0.99 + 0.01 * Db.var(0, unbiased=True), 2e-3)
        print("  nbt", int(nbt), "counter", int(counter))
        ok = ok and int(nbt) == 1 and int(counter) == 0
    if bwd is not None:
        dzb = Df  # statistics are defined on the stored (bf16) dz
        xhat = (H.float() - mean) * invstd
        s = dzb.sum(0)
        q = (dzb * xhat).sum(0)
        report("dbeta", db, s, 3e-3)
        report("dgamma", dg, q, 3e-3)
        sc = gamma * invstd
        m1, m2 = s / M, q / M
        report("ca", co[0], sc, 1e-4)
        report("cb", co[1], -sc * invstd * m2, 5e-3)
        report("cc", co[2], sc * (mean * invstd * m2 - m1), 5e-3)
    print("CASE %s %s" % (name, "PASS" if ok else "FAIL"))
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1:
        sys.exit(run_case(sys.argv[1]))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    summary = []
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name],
                               capture_output=True, text=True, timeout=180)
            out = r.stdout + r.stderr
            status = "PASS" if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or "") + (e.stderr or "") if isinstance(e.stdout, str) else "timeout"
            status = "TIMEOUT"
        summary.append("%-18s %s" % (name, status))
        print("==== %s: %s\n%s" % (name, status, out[-3000:]))
        sys.stdout.flush()
    print("\n".join(summary))
    with open(os.path.join(out_dir, "gemm_probe_summary.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")


if __name__ == "__main__":
    main()
