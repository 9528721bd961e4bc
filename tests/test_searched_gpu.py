"""Model-level parity of a searched network (reference models/searched_network.py with
InvertedResidualChannelsFused: several kernel sizes per block, Squeeze-and-Excitation, Swish, an
AtomNAS-style hidden width that is not a multiple of 8) against the reference's stock-torch graph:
eval logits and the first training steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# c, n, s, kernel sizes, hidden widths per kernel size, expand
ROWS = [[16, 1, 1, [3], [32], False],
        [24, 2, 2, [3, 5], [48, 24], True],
        [40, 2, 2, [3, 5, 7], [64, 36, 20], True],     # 36 / 20: not multiples of 8 (padded shadow)
        [80, 1, 2, [3], [160], True],
        [96, 1, 2, [3, 5], [96, 64], True]]
KW = dict(inverted_residual_setting=ROWS, block="InvertedResidualChannelsFused", se_ratio=0.5,
          active_fn="nn.Swish", batch_norm_momentum=0.01, batch_norm_epsilon=1e-3, num_classes=50,
          dropout_ratio=0.0, last_channel=320)
# the unfused packing (one expand/depthwise/project branch per kernel size, AtomNAS supernet)
KW_UNFUSED = dict(inverted_residual_setting=ROWS, block="InvertedResidualChannels",
                  active_fn="nn.ReLU6", batch_norm_momentum=0.01, batch_norm_epsilon=1e-3,
                  num_classes=50, dropout_ratio=0.0, last_channel=320)


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _model(size, kw=None):
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb, searched_network as sn
    torch.manual_seed(7)
    m = sn.Model(**(kw or KW), input_size=size)
    m.apply(mb.init_weights_mnas)
    return m


@pytest.mark.parametrize("kw", [KW, KW_UNFUSED], ids=["fused_se_swish", "unfused_relu6"])
def test_searched_eval_logits(built_lib, kw):
    from oracle import torch_model as tm
    m = _model(64, kw)
    g = torch.Generator().manual_seed(3)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1, generator=g)
            mod.running_var.uniform_(0.7, 1.3, generator=g)
    ref = tm.as_reference(m).eval()
    x = torch.randn(8, 3, 64, 64, generator=g)
    with torch.no_grad():
        want = ref(x)
        got = m.cuda().eval()(x.cuda())
    assert _rel(got, want) < 3e-2      # bf16 activations through 6 blocks + SE gates


@pytest.mark.parametrize("kw", [KW, KW_UNFUSED], ids=["fused_se_swish", "unfused_relu6"])
def test_searched_trains(built_lib, kw):
    from oracle import torch_model as tm
    from yet_another_mobilenet_series_b200.trainer import TrainStep
    B = 16
    m = _model(64, kw)
    ref = tm.as_reference(m).train()
    trainer = tm.RefTrainer(ref, B)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 64, 64, generator=g)
    t = torch.randint(0, 50, (B,), generator=g)
    m = m.cuda()
    ts = TrainStep(m, B, image_size=64)
    losses, losses_ref = [], []
    for _ in range(4):
        l2 = float(tm.l2_loss_mnas(ref, 1e-5))
        losses_ref.append(trainer.step(x, t) - l2)
        losses.append(float(ts(x.to(torch.bfloat16), t)))
    torch.cuda.synchronize()
    assert all(l == l for l in losses)                                   # no NaN
    assert abs(losses[0] - losses_ref[0]) < 3e-2 * abs(losses_ref[0]), (losses, losses_ref)
    assert losses[-1] < 0.7 * losses[0], losses                          # it trains
