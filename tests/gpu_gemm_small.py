"""GPU driver: phase timers (YAMB_GEMM_DEBUG=512) of the GEMM roles on the small late layers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402,F401
ge.build()
from gpu_microbench_gemm import run  # noqa: E402

for dbg in [int(a) for a in sys.argv[1:]] or [0, 512]:
    os.environ["YAMB_GEMM_DEBUG"] = str(dbg)
    print("---- YAMB_GEMM_DEBUG=%d" % dbg)
    sys.stdout.flush()
    it = 1 if dbg & 512 else 20
    run("expand b16 +stats", 12544, 960, 160, stats=True, iters=it)
    run("project b16 +xform+stats", 12544, 160, 960, stats=True, xform=1, iters=it)
    run("wgrad b16 +xform", 960, 160, 12544, xform=1, a_mn=1, b_mn=1, epi=2, iters=it)
    run("expand b9 +stats", 50176, 384, 64, stats=True, iters=it)
    run("project b9 +xform+stats", 50176, 64, 384, stats=True, xform=1, iters=it)
    run("wgrad b9 +xform", 384, 64, 50176, xform=1, a_mn=1, b_mn=1, epi=2, iters=it)
