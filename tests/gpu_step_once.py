"""Profiling driver: one full MobileNetV2 training step (eager, N=256) between
cudaProfilerStart/Stop, after 2 warm-up steps.  Used for the ncu launch list in profiles/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from yet_another_mobilenet_series_b200.trainer import TrainStep  # noqa: E402


def main():
    B = int(os.environ.get("YAMB_N", "256"))
    model = bench.build_model().cuda()
    ts = TrainStep(model, B, use_graph=False)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    ts.load(x, t)
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        ts.run()
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    print("loss", float(ts.loss))


if __name__ == "__main__":
    main()
