"""GPU driver: one GEMM shape a few times (target of an ncu capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402,F401
ge.build()
from gpu_microbench_gemm import run  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "project"
if which == "project":
    run("project b3 +xform+stats", 802816, 24, 144, stats=True, xform=1, iters=2)
elif which == "expand":
    run("expand b3 +stats", 802816, 144, 24, stats=True, iters=2)
else:
    run("wgrad b3 +xform", 144, 24, 802816, xform=1, a_mn=1, b_mn=1, epi=2, iters=2)
