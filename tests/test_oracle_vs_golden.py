"""CPU: pin the oracle (oracle/ir_block.py, oracle/optim.py) against the golden vectors produced
by the LIVE reference (oracle/make_golden.py) and against the reference's own known-answer
tests for the neighbouring services (tests/utils/optim_test.py of the reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import ir_block as ob
from oracle import optim as oo


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _gold(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _build(rec):
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb
    blk = getattr(mb, rec["cls"])(*rec["args"], active_fn=mb.get_active_fn(rec["act"]),
                                  batch_norm_kwargs=rec["bn"], **rec["extra"])
    blk.load_state_dict(rec["state"])
    return blk


CASES = ["v2_s2_relu", "v2_res_relu6", "v2_noexpand", "multi_k357", "fused_se_swish",
         "fused_s2_k5", "fused_plain"]
NL_CASES = ["nl_B_res", "nl_A_small_map", "nl_sub2_odd_map", "nl_s2_nores"]


@pytest.mark.parametrize("name", CASES + NL_CASES)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_oracle_block_matches_reference(golden_dir, name, mode):
    rec = _gold(golden_dir, "blocks_nl.pt" if name in NL_CASES else "blocks.pt")[name]
    blk = _build(rec)
    cfg, P = ob.extract(blk)
    tr = mode == "train"
    y, S = ob.forward(rec["x"], cfg, P, training=tr)
    dx, G = ob.backward(rec[mode]["dy"], cfg, P, S, training=tr)
    gold = rec[mode]
    assert _rel(y, gold["y"]) < 2e-6
    assert _rel(dx, gold["dx"]) < 5e-6
    # parameter gradients, mapped from the merged layout back to the module names
    names = dict(blk.named_parameters())
    fused = hasattr(blk, "expand_conv")
    got = {}
    if fused:
        if cfg.expand:
            got["expand_conv.0.weight"] = G["w_exp"]
            got["expand_conv.1.weight"], got["expand_conv.1.bias"] = G["bn1_g"], G["bn1_b"]
        c0 = 0
        for i, c in enumerate(cfg.channels):
            got["depth_ops.%d.1.0.weight" % i] = G["w_dw"][i]
            got["depth_ops.%d.1.1.weight" % i] = G["bn2_g"][c0:c0 + c]
            got["depth_ops.%d.1.1.bias" % i] = G["bn2_b"][c0:c0 + c]
            c0 += c
        got["project_conv.0.weight"] = G["w_proj"]
        got["project_conv.1.weight"], got["project_conv.1.bias"] = G["bn3_g"], G["bn3_b"]
        for k in ("se_wr", "se_br", "se_we", "se_be"):
            if k in G:
                tgt = {"se_wr": "se_op.se_reduce.weight", "se_br": "se_op.se_reduce.bias",
                       "se_we": "se_op.se_expand.weight", "se_be": "se_op.se_expand.bias"}[k]
                got[tgt] = G[k]
        if "w_nl" in G:   # non-local block: depthwise 3x3 + (ZeroInit)BN, :142-156
            got["nl_op.depthwise_conv.weight"] = G["w_nl"]
            got["nl_op.bn.weight"], got["nl_op.bn.bias"] = G["bn4_g"], G["bn4_b"]
    else:
        c0 = 0
        for i, c in enumerate(cfg.channels):
            j = 0
            if cfg.expand:
                got["ops.%d.0.0.weight" % i] = G["w_exp"][c0:c0 + c]
                got["ops.%d.0.1.weight" % i] = G["bn1_g"][c0:c0 + c]
                got["ops.%d.0.1.bias" % i] = G["bn1_b"][c0:c0 + c]
                j = 1
            got["ops.%d.%d.0.weight" % (i, j)] = G["w_dw"][i]
            got["ops.%d.%d.1.weight" % (i, j)] = G["bn2_g"][c0:c0 + c]
            got["ops.%d.%d.1.bias" % (i, j)] = G["bn2_b"][c0:c0 + c]
            got["ops.%d.%d.weight" % (i, j + 1)] = G["w_proj"][:, c0:c0 + c]
            c0 += c
        got["pw_bn.weight"], got["pw_bn.bias"] = G["bn3_g"], G["bn3_b"]
    assert set(got) == set(names)
    for k, g in got.items():
        assert _rel(g.reshape(gold["grads"][k].shape), gold["grads"][k]) < 2e-5, k
    if tr:  # running statistics after one training forward
        pairs = [("bn3", "pw_bn" if not fused else "project_conv.1")]
        if "w_nl" in G:
            pairs.append(("bn4", "nl_op.bn"))
        for pfx, key in pairs:
            rm, rv, _ = ob.bn_running_update(P[pfx + "_rm"], P[pfx + "_rv"], S[pfx + "_mean"],
                                             S[pfx + "_var"], S["count_out"], cfg.momentum, 0)
            assert _rel(rm, gold["state_after"][key + ".running_mean"]) < 1e-5
            assert _rel(rv, gold["state_after"][key + ".running_var"]) < 1e-5


def test_oracle_quant_mode_is_close_to_fp32(golden_dir):
    """bf16 rounding points move the result by about the bf16 budget, not more."""
    rec = _gold(golden_dir, "blocks.pt")["v2_res_relu6"]
    cfg, P = ob.extract(_build(rec))
    y32, _ = ob.forward(rec["x"], cfg, P, training=True)
    yq, _ = ob.forward(rec["x"], cfg, P, training=True, quant=True)
    assert 1e-4 < _rel(yq, y32) < 1e-2


@pytest.mark.parametrize("tag", ["mnas", "plain", "wd"])
def test_oracle_rmsprop_matches_reference(golden_dir, tag):
    rec = _gold(golden_dir, "optim.pt")[tag]
    kw = dict(rec["kw"])
    p = rec["p0"].numpy().copy()
    sq = np.zeros_like(p)
    mom = np.zeros_like(p)
    for i in range(rec["grads"].shape[0]):
        p, sq, mom, _ = oo.rmsprop_step(p, rec["grads"][i].numpy(), sq, mom, **kw)
        np.testing.assert_allclose(p, rec["ps"][i].numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(sq, rec["square_avg"].numpy(), rtol=2e-6, atol=1e-9)


def test_oracle_ema_matches_reference_and_known_answers(golden_dir):
    rec = _gold(golden_dir, "optim.pt")["ema"]
    sh = rec["v0"].numpy().copy()
    for t in range(rec["xs"].shape[0]):
        sh = oo.ema_update(sh, rec["xs"][t].numpy(), rec["decay"], t + 1)
        np.testing.assert_allclose(sh, rec["shadows"][t].numpy(), rtol=1e-6, atol=1e-7)
    # known answers of the reference's tests/utils/optim_test.py:113-128,161-172 (ported from TF):
    # decay 0.25 without num_updates; with num_updates=1 the effective decay is 2/11 = 0.181818
    assert oo.ema_momentum(0.25) == 0.25
    assert abs(oo.ema_momentum(0.25, 1) - 0.181818) < 1e-6
    v = oo.ema_update(np.array([10.0, 11.0], np.float32), np.array([20.0, 22.0], np.float32), 0.25)
    np.testing.assert_allclose(v, [10 * 0.25 + 20 * 0.75, 11 * 0.25 + 22 * 0.75], rtol=1e-6)
    # adjust_momentum equivalence (optim_test.py:191-215): 0.9999^(1/(4096/256))
    assert abs(oo.adjust_momentum(0.9999, 4096 / 256.0) - 0.9999 ** (256 / 4096.0)) < 1e-12


def test_oracle_l2_mask_and_grad(golden_dir):
    rec = _gold(golden_dir, "optim.pt")["l2"]
    mask = oo.l2_decay_mask([(k, tuple(v.shape)) for k, v in rec["params"].items()])
    assert mask == {"conv.weight": True, "dw.weight": True, "bn.weight": False, "bn.bias": False,
                    "classifier.weight": True, "classifier.bias": True}
    for k, p in rec["params"].items():
        want = rec["grads"][k].numpy()
        got = oo.l2_grad(p.numpy(), rec["wd"]) if mask[k] else np.zeros_like(want)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)


def test_label_smooth_ce_closed_form(golden_dir):
    """Reference known answer (tests/utils/optim_test.py:14-32): uniform logits, K classes,
    smoothing eps -> loss = log K for every sample; and the golden value of the live class."""
    rec = _gold(golden_dir, "optim.pt")["ce"]
    logp = torch.log_softmax(rec["logits"], 1)
    K = logp.shape[1]
    tgt = torch.zeros_like(logp).scatter_(1, rec["target"][:, None], 1) * 0.9 + 0.1 / K
    assert _rel(-(tgt * logp).sum(1), rec["loss"]) < 1e-6
