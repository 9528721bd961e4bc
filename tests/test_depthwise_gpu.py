"""GPU parity of the depthwise kernels against torch fp32 (F.conv2d / torch.nn.grad) of the same
op with the same bf16 rounding points.  Tolerances: outputs are bf16 (rel 2^-8 per element ->
rel-L2 4e-3), statistics / weight gradients are fp32 sums (1e-3)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_probe_gemm import _act, _act_grad  # noqa: E402


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def _nhwc(t):  # [N,C,H,W] float -> [N,H,W,C] bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


CASES = [
    # N, H, W, Ctot, c0, C, k, stride, act, prologue
    (2, 14, 14, 96, 0, 96, 3, 1, 1, True),
    (3, 15, 13, 144, 0, 144, 3, 2, 2, True),
    (2, 12, 12, 64, 16, 32, 5, 1, 3, True),
    (2, 12, 12, 64, 32, 32, 7, 2, 3, True),
    (2, 9, 9, 32, 0, 32, 3, 1, 0, False),
    (1, 7, 7, 960, 0, 960, 3, 1, 1, True),
    (2, 16, 16, 48, 8, 40, 5, 2, 4, True),
    # narrow branches of searched networks: 8- / 16-channel tiles (csrc/depthwise_narrow.cu) are
    # selected for C <= 8 with >= 28 rows and C <= 16 with >= 14 rows
    (2, 56, 56, 40, 24, 8, 3, 1, 3, True),       # AtomNAS block 3: 24 | 8 | 8 channels at 56 x 56
    (2, 56, 56, 40, 32, 8, 7, 1, 3, True),
    (2, 57, 59, 24, 0, 8, 5, 2, 1, True),        # odd sizes, stride 2
    (3, 30, 29, 56, 16, 16, 3, 2, 2, True),      # 16-channel tiles
    (2, 28, 28, 56, 40, 16, 5, 1, 3, True),
    (2, 33, 31, 16, 0, 16, 7, 2, 3, True),
    (2, 64, 64, 8, 0, 8, 3, 1, 0, False),        # no prologue
]


@pytest.mark.parametrize("N,H,W,Ct,c0,Cs,k,s,act,pro", CASES)
def test_depthwise_fwd_bwd(built_lib, N, H, W, Ct, c0, Cs, k, s, act, pro):
    from yet_another_mobilenet_series_b200 import native as nat
    lib = built_lib
    dev = "cuda"
    torch.manual_seed(1)
    pad = (k - 1) // 2
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    x = torch.randn(N, Ct, H, W, device=dev)
    xb = _nhwc(x)
    xf = xb.float().permute(0, 3, 1, 2)[:, c0:c0 + Cs]
    w = torch.randn(Cs, 1, k, k, device=dev) * 0.3
    sc = torch.rand(Cs, device=dev) + 0.5
    sh = torch.randn(Cs, device=dev) * 0.3
    nct = lib.yamb_max_ctas()
    partials = torch.zeros(nct * 2 * Cs, device=dev)
    counter = torch.zeros(1, device=dev, dtype=torch.int32)
    # ---------------- forward ----------------
    yb = torch.zeros(N, Ho, Wo, Ct, device=dev, dtype=torch.bfloat16)
    gamma = torch.rand(Cs, device=dev) + 0.5
    beta = torch.randn(Cs, device=dev)
    rm, rv = torch.zeros(Cs, device=dev), torch.ones(Cs, device=dev)
    nbt = torch.zeros(1, device=dev, dtype=torch.int64)
    o_scale, o_shift, o_mean, o_invstd = [torch.zeros(Cs, device=dev) for _ in range(4)]
    bn = nat.BnFwd()
    bn.partials, bn.counter = partials.data_ptr(), counter.data_ptr()
    bn.gamma, bn.beta, bn.eps, bn.momentum = gamma.data_ptr(), beta.data_ptr(), 1e-3, -1.0
    bn.running_mean, bn.running_var, bn.num_batches_tracked = rm.data_ptr(), rv.data_ptr(), \
        nbt.data_ptr()
    bn.scale, bn.shift, bn.mean, bn.invstd = o_scale.data_ptr(), o_shift.data_ptr(), \
        o_mean.data_ptr(), o_invstd.data_ptr()
    bn.count = N * Ho * Wo
    d = nat.DwFwd()
    d.N, d.H, d.W, d.C, d.ldc, d.k, d.stride = N, H, W, Cs, Ct, k, s
    d.x = xb.data_ptr() + c0 * 2
    if pro:
        d.in_scale, d.in_shift, d.in_act = sc.data_ptr(), sh.data_ptr(), act
    d.w, d.y = w.data_ptr(), yb.data_ptr() + c0 * 2
    d.bn = C.pointer(bn)
    nat.check(lib.yamb_depthwise_fwd(C.byref(d), nat.stream_handle()))
    torch.cuda.synchronize()
    # the kernel stages act(bn(x)) in shared memory as bf16
    a1 = _act(xf * sc[None, :, None, None] + sh[None, :, None, None], act).to(
        torch.bfloat16).float() if pro else xf
    y_acc = F.conv2d(a1, w, None, s, pad, 1, Cs)
    y_ref = y_acc.to(torch.bfloat16).float()
    y_got = yb.float().permute(0, 3, 1, 2)[:, c0:c0 + Cs]
    assert _rel(y_got, y_ref) < 4e-3
    if c0 > 0:  # channels outside the slice untouched
        assert float(yb[..., :c0].abs().max()) == 0.0
    # statistics come from the fp32 accumulators (before the bf16 rounding of the stored tensor)
    mean_ref = y_acc.mean((0, 2, 3))
    var_ref = y_acc.var((0, 2, 3), unbiased=False)
    assert _rel(o_mean, mean_ref) < 1e-3
    assert _rel(o_invstd, torch.rsqrt(var_ref + 1e-3)) < 1e-3
    cnt = N * Ho * Wo
    # momentum=None (cumulative): first update replaces the running statistics
    assert _rel(rm, mean_ref) < 1e-3
    assert _rel(rv, var_ref * cnt / (cnt - 1)) < 1e-3
    assert int(nbt) == 1 and int(counter) == 0
    # ---------------- backward ----------------
    dz = torch.randn(N, Ct, Ho, Wo, device=dev)
    h = torch.randn(N, Ct, Ho, Wo, device=dev)
    dzb, hb = _nhwc(dz), _nhwc(h)
    ca = torch.rand(Cs, device=dev) + 0.5
    cb = torch.randn(Cs, device=dev) * 0.2
    cc = torch.randn(Cs, device=dev) * 0.1
    mean1 = torch.randn(Cs, device=dev) * 0.1
    invstd1 = torch.rand(Cs, device=dev) + 0.5
    g1 = torch.rand(Cs, device=dev) + 0.5
    dw = torch.zeros(Cs, 1, k, k, device=dev)
    dxb = torch.zeros(N, H, W, Ct, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(Cs, device=dev), torch.zeros(Cs, device=dev)
    oca, ocb, occ = [torch.zeros(Cs, device=dev) for _ in range(3)]
    bb = nat.BnBwd()
    bb.partials, bb.counter = partials.data_ptr(), counter.data_ptr()
    bb.gamma, bb.mean, bb.invstd = g1.data_ptr(), mean1.data_ptr(), invstd1.data_ptr()
    bb.dgamma, bb.dbeta = dg.data_ptr(), db.data_ptr()
    bb.ca, bb.cb, bb.cc = oca.data_ptr(), ocb.data_ptr(), occ.data_ptr()
    bb.count = N * H * W
    bb.use_batch_stats = 1
    e = nat.DwBwd()
    e.N, e.H, e.W, e.C, e.ldc, e.k, e.stride = N, H, W, Cs, Ct, k, s
    e.dz, e.h = dzb.data_ptr() + c0 * 2, hb.data_ptr() + c0 * 2
    e.ca, e.cb, e.cc = ca.data_ptr(), cb.data_ptr(), cc.data_ptr()
    e.w, e.dw = w.data_ptr(), dw.data_ptr()
    e.x = xb.data_ptr() + c0 * 2
    res = None
    if pro:
        e.in_scale, e.in_shift, e.in_act = sc.data_ptr(), sh.data_ptr(), act
        e.bn = C.pointer(bb)
    else:
        res = _nhwc(torch.randn(N, Ct, H, W, device=dev))
        e.residual = res.data_ptr() + c0 * 2
    e.dx = dxb.data_ptr() + c0 * 2
    nat.check(lib.yamb_depthwise_bwd(C.byref(e), nat.stream_handle()))
    torch.cuda.synchronize()
    sl = slice(c0, c0 + Cs)
    dzf = dzb.float().permute(0, 3, 1, 2)[:, sl]
    hf = hb.float().permute(0, 3, 1, 2)[:, sl]
    v = lambda t: t[None, :, None, None]
    dh = (v(ca) * dzf + v(cb) * hf + v(cc)).to(torch.bfloat16).float()  # staged as bf16
    da = torch.nn.grad.conv2d_input(a1.shape, w, dh, s, pad, 1, Cs)
    dw_ref = torch.nn.grad.conv2d_weight(a1, w.shape, dh, s, pad, 1, Cs)
    if pro:
        z = xf * v(sc) + v(sh)
        dx_ref = da * _act_grad(z, act)
    else:
        dx_ref = da + res.float().permute(0, 3, 1, 2)[:, sl]
    dx_acc = dx_ref                      # fp32 values: what the kernel's statistics see
    dx_ref = dx_ref.to(torch.bfloat16).float()
    dx_got = dxb.float().permute(0, 3, 1, 2)[:, sl]
    assert _rel(dx_got, dx_ref) < 4e-3
    assert _rel(dw, dw_ref) < 1e-3
    if pro:
        xhat = (xf - v(mean1)) * v(invstd1)
        # BatchNorm-backward sums are taken from the fp32 gradients (before the bf16 rounding of
        # the stored dx), like the forward statistics
        s_ref = dx_acc.sum((0, 2, 3))
        q_ref = (dx_acc * xhat).sum((0, 2, 3))
        assert _rel(db, s_ref) < 2e-3
        assert _rel(dg, q_ref) < 2e-3
        M = N * H * W
        scl = g1 * invstd1
        assert _rel(oca, scl) < 1e-5
        assert _rel(ocb, -scl * invstd1 * q_ref / M) < 2e-3
        assert _rel(occ, scl * (mean1 * invstd1 * q_ref / M - s_ref / M)) < 2e-3
