"""GPU driver (round 2): yamb_pointwise_gemm on the MobileNetV2 N=256 shapes of every class, so
that knobs (YAMB_GEMM_DEBUG bits, the -DYAMB_GEMM_TIMERS build via YAMB_LIB_PATH) can be compared
per shape.  Usage: python tests/gpu_microbench_r2.py [filter-substring]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_microbench_gemm import run  # noqa: E402

B = 256
SHAPES = {  # block: (pixels_in, pixels_out, Cin, Chid, Cout)
    "b1": (B * 112 * 112, B * 112 * 112, 32, 32, 16),
    "b2": (B * 112 * 112, B * 56 * 56, 16, 96, 24),
    "b3": (B * 56 * 56, B * 56 * 56, 24, 144, 24),
    "b5": (B * 28 * 28, B * 28 * 28, 32, 192, 32),
    "b8": (B * 14 * 14, B * 14 * 14, 64, 384, 64),
    "b12": (B * 14 * 14, B * 14 * 14, 96, 576, 96),
    "b15": (B * 7 * 7, B * 7 * 7, 160, 960, 160),
}


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, (mi, mo, cin, chid, cout) in SHAPES.items():
        cases = []
        if name != "b1":
            cases.append(("expand_fwd", dict(M=mi, N=chid, K=cin, stats=True)))
        cases.append(("project_fwd", dict(M=mo, N=cout, K=chid, stats=True, xform=1)))
        cases.append(("project_dgrad", dict(M=mo, N=chid, K=cout, b_mn=1, dgrad=True)))
        cases.append(("project_wgrad", dict(M=cout, N=chid, K=mo, a_mn=1, b_mn=1, epi=2, wgrad=True)))
        for tag, kw in cases:
            full = "%s %s" % (name, tag)
            if flt and flt not in full:
                continue
            run(full, **kw)


if __name__ == "__main__":
    main()
