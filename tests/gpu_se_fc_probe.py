"""GPU driver: the SE fully-connected kernels at an AtomNAS-C+ size (run under
`ncu --metrics gpu__time_duration.sum` to see the three kernels' durations separately)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yet_another_mobilenet_series_b200 import native as nat  # noqa: E402

N, Cc, R = int(os.environ.get("N", 256)), int(os.environ.get("C", 2320)), int(os.environ.get("R", 112))
dev = "cuda"
lib = nat.lib()
t = lambda *s: torch.randn(*s, device=dev)
pooled, wr, br, we, be = t(N, Cc), t(R, Cc) * 0.02, t(R), t(Cc, R) * 0.05, t(Cc)
u, v, gate = t(N, R), t(N, R), t(N, Cc)
f = nat.SeFc()
f.N, f.C, f.R, f.act = N, Cc, R, 3
f.pooled, f.w_r, f.b_r, f.w_e, f.b_e = (x.data_ptr() for x in (pooled, wr, br, we, be))
f.u, f.v, f.gate = u.data_ptr(), v.data_ptr(), gate.data_ptr()
b = nat.SeFcBwd()
b.N, b.C, b.R, b.act, b.inv_hw = N, Cc, R, 3, 1.0 / 49
dgate, dpool, dt, du = t(N, Cc), t(N, Cc), t(N, Cc), t(N, R)
g = [torch.zeros_like(x) for x in (wr, br, we, be)]
b.dgate, b.gate, b.u, b.v, b.pooled = (x.data_ptr() for x in (dgate, gate, u, v, pooled))
b.w_r, b.w_e, b.dpool, b.dt, b.du = (x.data_ptr() for x in (wr, we, dpool, dt, du))
b.g_wr, b.g_br, b.g_we, b.g_be = (x.data_ptr() for x in g)
st = nat.stream_handle()
for _ in range(3):
    nat.check(lib.yamb_se_fc_fwd(C.byref(f), st))
    nat.check(lib.yamb_se_fc_bwd(C.byref(b), st))
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(10):
    nat.check(lib.yamb_se_fc_fwd(C.byref(f), st))
e1.record()
for _ in range(10):
    nat.check(lib.yamb_se_fc_bwd(C.byref(b), st))
e2.record()
torch.cuda.synchronize()
print("se_fc fwd %.1f us, bwd (sample + param kernels) %.1f us" % (e0.elapsed_time(e1) * 100, e1.elapsed_time(e2) * 100))
