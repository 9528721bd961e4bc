"""GPU driver: time output-heavy GEMMs with the store-path experiment bits of YAMB_GEMM_DEBUG
(64: no store at all, 128: coalesced st.global instead of the TMA store)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402,F401
ge.build()
from gpu_microbench_gemm import run  # noqa: E402

for dbg in (512, 512 + 128, 512 + 64):
    os.environ["YAMB_GEMM_DEBUG"] = str(dbg)
    print("---- YAMB_GEMM_DEBUG=%d" % dbg)
    run("expand b3 plain", 802816, 144, 24, iters=1)
    run("expand b3 +stats", 802816, 144, 24, stats=True, iters=1)
    run("expand b2 plain", 3211264, 96, 16, iters=1)
    run("project b3 +xform+stats", 802816, 24, 144, stats=True, xform=1, iters=1)
