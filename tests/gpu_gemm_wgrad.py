"""GPU driver: project / expand weight-gradient GEMM shapes with the phase timers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pass
import __graft_entry__ as ge  # noqa: E402,F401
ge.build()
from gpu_microbench_gemm import run  # noqa: E402

for dbg in [int(a) for a in sys.argv[1:]] or [0, 512]:
    os.environ["YAMB_GEMM_DEBUG"] = str(dbg)
    print("---- YAMB_GEMM_DEBUG=%d" % dbg)
    sys.stdout.flush()
    it = 1 if dbg & 512 else 10
    run("project wgrad b2 (24x96)", 24, 96, 802816, a_mn=1, b_mn=1, epi=2, wgrad=True, iters=it)
    run("project wgrad b3 (24x144)", 24, 144, 802816, a_mn=1, b_mn=1, epi=2, wgrad=True, iters=it)
    run("project wgrad b9 (64x384)", 64, 384, 50176, a_mn=1, b_mn=1, epi=2, wgrad=True, iters=it)
