"""GPU driver: is the train-mode forward reproducible run to run, model to model?  Prints per-block
relative L2 difference of the block outputs (expect ~0 or isolated bf16 ulps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_model_gpu import _model


def run(m, x):
    outs = []
    hs = [f.register_forward_hook(lambda mod, i, o: outs.append(o.detach().float().clone()))
          for f in m.features]
    m.train()
    with torch.no_grad():
        y = m(x)
    for h in hs:
        h.remove()
    torch.cuda.synchronize()
    return outs + [y.detach().float().clone()]


def cmp(a, b, tag):
    worst = 0.0
    for i, (p, q) in enumerate(zip(a, b)):
        d = float((p - q).norm() / (q.norm() + 1e-20))
        worst = max(worst, d)
        if d > 0:
            print("%s: output %2d shape %s rel %.3e max|d| %.3e" % (tag, i, tuple(p.shape), d,
                                                                    float((p - q).abs().max())))
    print("%s: worst rel %.3e" % (tag, worst))


for size, B in ((64, 8), (96, 16)):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, size, size, generator=g).to(torch.bfloat16).cuda().contiguous(
        memory_format=torch.channels_last)
    m1 = _model(size).cuda()
    a = run(m1, x)
    b = run(m1, x)
    cmp(a, b, "size %d same model twice" % size)
    junk = [torch.randn(1 << 22, device="cuda") for _ in range(8)]
    del junk
    m2 = _model(size).cuda()
    c = run(m2, x)
    cmp(c, a, "size %d second model" % size)
