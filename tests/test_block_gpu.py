"""GPU parity of the fused inverted-residual block (C-ABI kernels sequenced by engine.py).

Two references, both on the same seeded inputs:
  * the committed golden fixtures produced by the LIVE reference modules in fp32
    (tests/golden/blocks.pt, oracle/make_golden.py).  The CUDA path computes in bf16 with fp32
    accumulation, so the bound is the reference's own bf16 error budget (SURVEY.md §8c measured
    3.6e-3..5.6e-3 rel-L2 forward, 3e-2 dgrad / 4e-2 wgrad for ReLU blocks for the reference under
    autocast-bf16; ReLU-mask flips of near-zero pre-activations dominate): forward 1e-2,
    gradients 1e-1 rel-L2.  (oracle quant=True vs the same fixtures measures 3.8e-3..5.8e-3
    forward, 1.4e-2..7.0e-2 dx and up to 2.6e-1 on a BN-gamma gradient of the 288-sample
    multi-branch case: the bound is the bf16 budget of these tiny batches, not kernel error; the
    kernels are held to the quant oracle below.)
  * the oracle in `quant=True` mode, which rounds to bf16 at exactly the points where the CUDA
    path materialises bf16 tensors: forward 3e-3, gradients 1.5e-2 rel-L2 (north-star "1e-3 rel"
    is met per element up to bf16 output rounding, 2^-9 = 2e-3).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _build(rec, dev):
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb
    blk = getattr(mb, rec["cls"])(*rec["args"], active_fn=mb.get_active_fn(rec["act"]),
                                  batch_norm_kwargs=rec["bn"], **rec["extra"])
    blk.load_state_dict(rec["state"])
    return blk.to(dev)


def _cases():
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    rec = torch.load(os.path.join(d, "blocks.pt"), weights_only=False)
    # non-local blocks (reference models/mobilenet_base.py:131-178), live-reference fixtures
    rec.update(torch.load(os.path.join(d, "blocks_nl.pt"), weights_only=False))
    return rec


GOLD = _cases()
SUPPORTED = list(GOLD)


@pytest.mark.parametrize("name", SUPPORTED)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_block_vs_golden_and_oracle(built_lib, name, mode):
    from oracle import ir_block as ob
    rec = GOLD[name]
    dev = torch.device("cuda")
    blk = _build(rec, dev)
    blk.train(mode == "train")
    x = rec["x"].to(dev).requires_grad_(True)
    y = blk(x)
    assert y.dtype == torch.bfloat16 and y.shape == rec[mode]["y"].shape
    dy = rec[mode]["dy"].to(dev)
    y.backward(dy.to(y.dtype))
    torch.cuda.synchronize()
    gold = rec[mode]
    # ---- the oracle with the same bf16 rounding points ----
    blk_cpu = _build(rec, "cpu")
    cfg, P = ob.extract(blk_cpu)
    yo, S = ob.forward(rec["x"], cfg, P, training=(mode == "train"), quant=True)
    dxo, G = ob.backward(dy.cpu(), cfg, P, S, training=(mode == "train"), quant=True)
    # ---- vs the live reference (fp32): no further away than the bf16 rounding explains ----
    assert _rel(y, gold["y"]) < 1e-2
    assert _rel(x.grad, gold["dx"]) < 1.2 * _rel(dxo, gold["dx"]) + 1e-2
    worst = 0.0
    for k, p in blk.named_parameters():
        assert p.grad is not None, k
        worst = max(worst, _rel(p.grad, gold["grads"][k]))
    assert worst < 0.35, worst  # small-sample bf16 budget (see module docstring); tight check below
    if mode == "train":
        for k, v in blk.state_dict().items():
            ref = gold["state_after"][k]
            if "running_" in k:
                # non-local branch: f sums up to H'W' products per element, O(100) in magnitude;
                # the bf16 rounding of f and of the depthwise output (2^-9 relative per element)
                # is then percent-level on the variance the fp32 fixture holds
                rt = 3e-2 if k.startswith("nl_op") else 5e-3
                assert torch.allclose(v.cpu(), ref, rtol=rt, atol=2e-3), k
            elif "num_batches_tracked" in k:
                assert int(v) == int(ref), k
    assert _rel(y, yo) < 3e-3
    assert _rel(x.grad, dxo) < 1.5e-2
    # merged-layout gradients of the CUDA path
    cfg2, P2 = ob.extract(blk)  # tensors are references: read .grad through the modules

    def cat_grads(convs, dim):
        return torch.cat([c.weight.grad.flatten(1) for c in convs], dim)

    fused = hasattr(blk, "expand_conv")
    if fused:
        exp = [blk.expand_conv[0]] if blk.expand else []
        dws = [list(op.children())[-1] for op in blk.depth_ops]
        proj = [blk.project_conv[0]]
        bn1 = [blk.expand_conv[1]] if blk.expand else []
        bn3 = blk.project_conv[1]
    else:
        exp = [op[0][0] for op in blk.ops] if blk.expand else []
        bn1 = [op[0][1] for op in blk.ops] if blk.expand else []
        dws = [op[1] if blk.expand else op[0] for op in blk.ops]
        proj = [op[2] if blk.expand else op[1] for op in blk.ops]
        bn3 = blk.pw_bn
    if exp:
        assert _rel(cat_grads(exp, 0), G["w_exp"]) < 1.5e-2
        assert _rel(torch.cat([b.weight.grad for b in bn1]), G["bn1_g"]) < 1.5e-2
        assert _rel(torch.cat([b.bias.grad for b in bn1]), G["bn1_b"]) < 1.5e-2
    assert _rel(cat_grads(proj, 1), G["w_proj"]) < 1.5e-2
    for d, gref in zip(dws, G["w_dw"]):
        assert _rel(d[0].weight.grad[:, 0], gref) < 1.5e-2
    assert _rel(torch.cat([d[1].weight.grad for d in dws]), G["bn2_g"]) < 1.5e-2
    assert _rel(torch.cat([d[1].bias.grad for d in dws]), G["bn2_b"]) < 1.5e-2
    assert _rel(bn3.weight.grad, G["bn3_g"]) < 1.5e-2
    assert _rel(bn3.bias.grad, G["bn3_b"]) < 1.5e-2
    if "w_nl" in G:
        nl = blk.nl_op
        assert _rel(nl.depthwise_conv.weight.grad[:, 0], G["w_nl"]) < 1.5e-2
        assert _rel(nl.bn.weight.grad, G["bn4_g"]) < 1.5e-2
        assert _rel(nl.bn.bias.grad, G["bn4_b"]) < 1.5e-2
    if "se_wr" in G:
        se = blk.se_op
        assert _rel(se.se_reduce.weight.grad.flatten(1), G["se_wr"]) < 1.5e-2
        assert _rel(se.se_reduce.bias.grad, G["se_br"]) < 1.5e-2
        assert _rel(se.se_expand.weight.grad.flatten(1), G["se_we"]) < 1.5e-2
        assert _rel(se.se_expand.bias.grad, G["se_be"]) < 1.5e-2


def test_block_rejects_cpu_input(built_lib):
    """No CPU fallback: the product path fails loudly off-GPU."""
    from yet_another_mobilenet_series_b200 import native as nat
    rec = GOLD["v2_s2_relu"]
    blk = _build(rec, "cpu")
    with pytest.raises(nat.NativeError):
        blk(rec["x"])


def test_bn_calibration_cumulative(built_lib):
    """bn_calibration semantics (reference utils/common.py:175-187): block in eval mode, BN
    children in train mode with momentum=None -> cumulative average of batch statistics."""
    from oracle import ir_block as ob
    rec = GOLD["v2_res_relu6"]
    dev = torch.device("cuda")
    blk = _build(rec, dev).eval()
    ref = _build(rec, "cpu")
    cfg, P = ob.extract(ref)
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.train()
            m.momentum = None
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(rec["x"].shape, generator=g) for _ in range(3)]
    means = []
    with torch.no_grad():
        for xx in xs:
            blk(xx.to(dev))
            _, S = ob.forward(xx, cfg, P, training=True, quant=True)
            means.append(S["bn3_mean"])
    torch.cuda.synchronize()
    bn3 = blk.pw_bn
    assert int(bn3.num_batches_tracked) == 3
    want = torch.stack(means).mean(0)
    assert torch.allclose(bn3.running_mean.cpu(), want, rtol=5e-3, atol=2e-3)


@pytest.mark.parametrize("cls,act,extra", [
    ("InvertedResidualChannels", "nn.ReLU6", {}),
    ("InvertedResidualChannelsFused", "nn.Swish", {"se_ratio": 0.5}),
])
def test_odd_width_block_runs_through_padded_shadow(built_lib, cls, act, extra):
    """AtomNAS-style hidden widths (not multiples of 8; reference
    apps/searched/models/atomnas_c.yml:14) go through the zero-padded shadow block; results are
    compared with the oracle evaluated on the REAL (unpadded) parameters."""
    from oracle import ir_block as ob
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb
    torch.manual_seed(5)
    bn = {"momentum": 0.01, "eps": 1e-3}
    args = (24, 24, 1, [15, 23, 13], [3, 5, 7], True)
    blk = getattr(mb, cls)(*args, active_fn=mb.get_active_fn(act), batch_norm_kwargs=bn, **extra)
    blk.apply(mb.init_weights_mnas)
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.3)
    ref = getattr(mb, cls)(*args, active_fn=mb.get_active_fn(act), batch_norm_kwargs=bn, **extra)
    ref.load_state_dict(blk.state_dict())
    x = torch.randn(4, 24, 12, 12)
    dy = torch.randn(4, 24, 12, 12)
    blk = blk.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = blk(xg)
    y.backward(dy.cuda().to(y.dtype))
    torch.cuda.synchronize()
    cfg, P = ob.extract(ref)
    yo, S = ob.forward(x, cfg, P, training=True, quant=True)
    dxo, G = ob.backward(dy, cfg, P, S, training=True, quant=True)
    assert _rel(y, yo) < 3e-3
    assert _rel(xg.grad, dxo) < 1.5e-2
    for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.shape, k
    fused = hasattr(blk, "expand_conv")
    proj = [blk.project_conv[0]] if fused else [op[2] for op in blk.ops]
    got = torch.cat([c.weight.grad.flatten(1) for c in proj], 1)
    assert _rel(got, G["w_proj"]) < 1.5e-2
    # running statistics came back un-padded
    bn3 = blk.project_conv[1] if fused else blk.pw_bn
    rm, _, _ = ob.bn_running_update(P["bn3_rm"], P["bn3_rv"], S["bn3_mean"], S["bn3_var"],
                                    S["count_out"], cfg.momentum, 0)
    assert torch.allclose(bn3.running_mean.cpu(), rm, rtol=5e-3, atol=2e-3)
    assert int(bn3.num_batches_tracked) == 1
