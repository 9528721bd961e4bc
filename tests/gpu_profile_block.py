"""Profiling driver (run under ncu on the GPU box): forward+backward of MobileNetV2 blocks at the
bench shapes (N=256), eager (no graph), 2 warm-up iterations then one profiled iteration between
cudaProfilerStart/Stop."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yet_another_mobilenet_series_b200 import mobilenet_base as mb  # noqa: E402

# (inp, oup, stride, hidden, H)
SHAPES = {"b1": (32, 16, 1, 32, 112), "b2": (16, 24, 2, 96, 112), "b3": (24, 24, 1, 144, 56),
          "b6": (32, 32, 1, 192, 28), "b8": (64, 64, 1, 384, 14), "b9": (64, 64, 1, 384, 14),
          "b13": (96, 96, 1, 576, 14), "b15": (160, 160, 1, 960, 7), "b16": (160, 160, 1, 960, 7)}


def main_eval(names):
    """YAMB_PROFILE_EVAL=1: eval-mode forward of the same blocks under no_grad (the one-launch
    kernel of csrc/block_eval.cu where it applies)."""
    N = int(os.environ.get("YAMB_N", "256"))
    dev = torch.device("cuda")
    bn = {"momentum": 0.01, "eps": 1e-3}
    blocks = []
    for n in names:
        inp, oup, s, hid, H = SHAPES[n]
        torch.manual_seed(0)
        blk = mb.InvertedResidualChannels(inp, oup, s, [hid], [3], hid != inp,
                                          mb.get_active_fn("nn.ReLU"), bn).to(dev)
        blk.apply(mb.init_weights_mnas)
        blk.eval()
        x = torch.randn(N, inp, H, H, device=dev).to(torch.bfloat16).contiguous(
            memory_format=torch.channels_last)
        blocks.append((blk, x))
    with torch.no_grad():
        for it in range(3):
            if it == 2:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
            for blk, x in blocks:
                blk(x)
            if it == 2:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
    print("done")


def main():
    names = sys.argv[1:] or ["b2", "b3"]
    if os.environ.get("YAMB_PROFILE_EVAL"):
        return main_eval(names)
    N = int(os.environ.get("YAMB_N", "256"))
    dev = torch.device("cuda")
    bn = {"momentum": 0.01, "eps": 1e-3}
    blocks = []
    for n in names:
        inp, oup, s, hid, H = SHAPES[n]
        torch.manual_seed(0)
        blk = mb.InvertedResidualChannels(inp, oup, s, [hid], [3], hid != inp,
                                          mb.get_active_fn("nn.ReLU"), bn).to(dev).train()
        blk.apply(mb.init_weights_mnas)
        x = torch.randn(N, inp, H, H, device=dev).to(torch.bfloat16).contiguous(
            memory_format=torch.channels_last).requires_grad_(True)
        blocks.append((blk, x))
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        for blk, x in blocks:
            y = blk(x)
            y.backward(torch.ones_like(y))
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
