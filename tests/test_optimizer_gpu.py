"""GPU parity of the fused flat-arena RMSprop kernel against the oracle restatement of
utils/rmsprop.py and against the golden sequences produced by the live reference optimizer.
Tolerance: parameters and square_avg <= 2e-6 rel-L2 over 12 steps (fp32; the kernel may contract
a*b+c into FMA), the remaining state <= 1e-5."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.detach().float().cpu().numpy().astype(np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("tag", ["mnas", "plain", "wd"])
def test_rmsprop_vs_golden(built_lib, golden_dir, tag):
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    rec = torch.load(os.path.join(golden_dir, "optim.pt"), weights_only=False)[tag]
    p = torch.nn.Parameter(rec["p0"].clone().cuda())
    opt = RMSprop([p], **rec["kw"])
    for i in range(rec["grads"].shape[0]):
        opt.zero_grad()
        p.grad.copy_(rec["grads"][i])
        opt.step()
        assert _rel(p, rec["ps"][i].numpy()) < 2e-6, i   # SURVEY 8c gate iii: 1e-6 rel over >= 10 steps (rel-L2; fp32 FMA contraction)
    st = opt.state[p]
    assert _rel(st["square_avg"], rec["square_avg"].numpy()) < 2e-6
    if rec["kw"].get("momentum", 0) > 0:
        assert _rel(st["momentum_buffer"], rec["momentum_buffer"].numpy()) < 1e-5
    assert st["step"] == rec["grads"].shape[0]


def test_rmsprop_multi_tensor_l2_ema_scale(built_lib):
    """Several tensors (odd sizes -> padded arena), folded L2 mask, EMA warm-up rule, 1/world
    gradient scale, bf16 mirror — against oracle/optim.py."""
    from oracle import optim as oo
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    torch.manual_seed(0)
    shapes = {"features.0.0.weight": (8, 3, 3, 3), "features.0.1.weight": (8,),
              "features.0.1.bias": (8,), "features.1.ops.0.0.0.weight": (8, 1, 3, 3),
              "classifier.1.weight": (5, 13), "classifier.1.bias": (5,)}
    params = {k: torch.nn.Parameter(torch.randn(s).cuda()) for k, s in shapes.items()}
    kw = dict(lr=0.016, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True)
    opt = RMSprop(list(params.values()), **kw)
    opt.fold_l2(1e-2, list(params.items()))
    decay = 0.9999 ** (256 / 4096.0)
    opt.attach_ema(decay)
    opt.grad_scale = 0.25
    mask = oo.l2_decay_mask([(k, s) for k, s in shapes.items()])
    ref = {k: dict(p=v.detach().cpu().numpy().ravel().copy(), sq=np.zeros(v.numel(), np.float32),
                   mom=np.zeros(v.numel(), np.float32)) for k, v in params.items()}
    for k in ref:
        ref[k]["ema"] = ref[k]["p"].copy()
    for t in range(1, 12):
        opt.zero_grad()
        for k, p in params.items():
            g = torch.randn(p.shape) * 2
            p.grad.copy_(g)
            r = ref[k]
            gg = g.numpy().ravel().astype(np.float32) * np.float32(0.25)
            if mask[k]:
                gg = gg + oo.l2_grad(r["p"], 1e-2)
            r["p"], r["sq"], r["mom"], _ = oo.rmsprop_step(r["p"], gg, r["sq"], r["mom"], **kw)
            r["ema"] = oo.ema_update(r["ema"], r["p"], decay, t)
        opt.step(num_updates=t)
    for k, p in params.items():
        assert _rel(p, ref[k]["p"].reshape(p.shape)) < 1e-5, k
        assert _rel(opt.ema_shadow(p), ref[k]["ema"].reshape(p.shape)) < 1e-5, k
        assert _rel(p._yamb_bf16, p.detach().to(torch.bfloat16).float().cpu().numpy()) == 0.0
        assert p.grad.data_ptr() >= opt.arenas()["g"].data_ptr()


def test_state_dict_roundtrip(built_lib):
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(7, 5).cuda()), torch.nn.Parameter(torch.randn(11).cuda())]
    opt = RMSprop(ps, lr=0.01, alpha=0.9, momentum=0.9)
    for _ in range(3):
        opt.zero_grad()
        for p in ps:
            p.grad.copy_(torch.randn_like(p))
        opt.step()
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "square_avg", "momentum_buffer"}
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = RMSprop(ps2, lr=0.01, alpha=0.9, momentum=0.9)
    opt2.load_state_dict(sd)
    g = [torch.randn_like(p) for p in ps]
    for o, pl in ((opt, ps), (opt2, ps2)):
        o.zero_grad()
        for p, gg in zip(pl, g):
            p.grad.copy_(gg)
        o.step()
    for a, b in zip(ps, ps2):
        assert torch.equal(a.detach(), b.detach())


def test_negative_hyper_raises():
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    p = [torch.nn.Parameter(torch.zeros(3))]
    for kw in (dict(lr=-1), dict(eps=-1), dict(momentum=-1), dict(weight_decay=-1),
               dict(alpha=-1)):
        with pytest.raises(ValueError):
            RMSprop(p, **kw)


def test_stale_bf16_mirror_is_refreshed_after_inplace_weight_write(built_lib):
    """the block kernels read the optimizer's bf16 weight mirror.  A load_state_dict /
    broadcast / re-init AFTER the arenas were built writes the fp32 masters in place; the next
    forward must see the new weights (the mirror is re-cast when `_version` moved)."""
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    torch.manual_seed(3)
    bn = {"momentum": 0.01, "eps": 1e-3}

    def make():
        b = mb.InvertedResidualChannels(16, 24, 1, [96], [3], True, mb.get_active_fn("nn.ReLU"), bn)
        b.apply(mb.init_weights_mnas)
        return b.cuda().eval()

    blk, other = make(), make()
    opt = RMSprop(blk.parameters(), lr=0.01, momentum=0.9)
    opt.arenas()                                   # mirrors exist from here on
    x = torch.randn(4, 16, 12, 12, device="cuda")
    with torch.no_grad():
        y0 = blk(x).float()
        blk.load_state_dict(other.state_dict())    # in-place copy_ into the arena views
        y1 = blk(x).float()
        want = other(x).float()
    assert float((y1 - want).abs().max()) == 0.0   # same kernels, same weights: bit-identical
    assert float((y1 - y0).abs().max()) > 1e-2     # and it really changed
    opt.sync_mirror()                              # the BatchNorm / depthwise masters moved too
    assert opt.sync_mirror() == 0                  # nothing left stale


def test_step_uses_grads_that_left_the_arena_and_skips_gradless_params(built_lib):
    """`model.zero_grad(set_to_none=True)` detaches the `.grad` views; autograd then
    allocates fresh gradient tensors.  step() must use them (copy into the arena, re-attach), and
    a parameter without a gradient is skipped like the reference does (utils/rmsprop.py:77-78)."""
    from yet_another_mobilenet_series_b200.fused_rmsprop import RMSprop
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(37, device="cuda"))
    b = torch.nn.Parameter(torch.randn(5, 3, device="cuda"))
    opt = RMSprop([a, b], lr=0.1, alpha=0.9, momentum=0.9, eps=1e-3, eps_inside_sqrt=True)
    opt.arenas()
    a0, b0 = a.detach().clone(), b.detach().clone()
    a.grad = None
    b.grad = None                                   # what nn.Module.zero_grad() does
    (a * 2.0).sum().backward()                      # fresh .grad tensor for a, none for b
    assert a.grad.data_ptr() != opt.arenas()["gptr"][0]
    opt.step()
    torch.cuda.synchronize()
    g = 2.0
    sq = 0.1 * g * g
    want = a0 - 0.1 * (g / (sq + 1e-3) ** 0.5)
    assert torch.allclose(a, want, rtol=1e-6, atol=1e-7)
    assert torch.equal(b, b0)                       # skipped: no state change at all
    assert float(opt.state[b]["square_avg"].abs().max()) == 0.0
    assert opt.state[a]["step"] == 1 and opt.state[b]["step"] == 0
    assert a.grad.data_ptr() == opt.arenas()["gptr"][0]   # re-attached to the arena
