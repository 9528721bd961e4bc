"""GPU driver: per-launch device time of one eager training step (MobileNetV2-1.0, N=256), in
launch order: forward blocks 1..17 (expand, dw, project, bn_apply), then backward 17..1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

ge.build()
from yet_another_mobilenet_series_b200 import engine  # noqa: E402
from yet_another_mobilenet_series_b200.trainer import TrainStep  # noqa: E402

B = int(os.environ.get("YAMB_N", "256"))
dev = torch.device("cuda")
model = bench.build_model().to(dev)
ts = TrainStep(model, B, use_graph=False)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 1000, (B,), generator=g)
ts.load(x, t)
for _ in range(3):
    ts.run()
torch.cuda.synchronize()
acc = None
R = 3
for _ in range(R):
    engine.PROFILE = []
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2.0e8))
    ts.run()
    torch.cuda.synchronize()
    rows = [(tag, nb, a.elapsed_time(b)) for tag, nb, fl, a, b in engine.PROFILE]
    if acc is None:
        acc = [[tag, nb, ms] for tag, nb, ms in rows]
    else:
        for r, (tag, nb, ms) in zip(acc, rows):
            r[2] += ms
engine.PROFILE = None
tot = 0.0
for i, (tag, nb, ms) in enumerate(acc):
    ms /= R
    tot += ms
    print("%3d %-18s %8.1f us  %8.1f MB  %7.1f GB/s" % (i, tag, ms * 1e3, nb / 1e6, nb / ms / 1e6 if ms else 0))
print("total of our launches: %.3f ms" % tot)
