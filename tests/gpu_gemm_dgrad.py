"""GPU driver: project-dgrad GEMM shapes (epilogue kind 1) with and without phase timers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402,F401
os.environ.setdefault("YAMB_GEMM_TIMERS", "1")
ge.build(force=True)   # phase timers are a compile-time option
from gpu_microbench_gemm import run  # noqa: E402

for dbg in [int(a) for a in sys.argv[1:]] or [0, 512]:
    os.environ["YAMB_GEMM_DEBUG"] = str(dbg)
    print("---- YAMB_GEMM_DEBUG=%d" % dbg)
    sys.stdout.flush()
    it = 1 if dbg & 512 else 10
    run("project dgrad b2", 802816, 96, 24, b_mn=1, dgrad=True, iters=it)
    run("project dgrad b3", 802816, 144, 24, b_mn=1, dgrad=True, iters=it)
    run("project dgrad b9", 50176, 384, 64, b_mn=1, dgrad=True, iters=it)
