"""GPU driver: phase timers (YAMB_GEMM_DEBUG=512) of the GEMM roles on representative shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402,F401
os.environ.setdefault("YAMB_GEMM_TIMERS", "1")
ge.build(force=True)   # phase timers are a compile-time option
from gpu_microbench_gemm import run  # noqa: E402

for dbg in [int(a) for a in sys.argv[1:]] or [512]:
    os.environ["YAMB_GEMM_DEBUG"] = str(dbg)
    print("---- YAMB_GEMM_DEBUG=%d" % dbg)
    sys.stdout.flush()
    run("expand b3 +stats", 802816, 144, 24, stats=True, iters=1)
    run("project b3 +xform+stats", 802816, 24, 144, stats=True, xform=1, iters=1)
    run("wgrad b3 +xform", 144, 24, 802816, xform=1, a_mn=1, b_mn=1, epi=2, iters=1)
    run("project b12 +xform+stats", 50176, 96, 576, stats=True, xform=1, iters=1)
