import sys, os
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import gpu_microbench_gemm as g
M = 256*56*56
g.run("project b3 plain", M, 24, 144)
g.run("project b3 +xform dbg=%s" % os.environ.get("YAMB_GEMM_DEBUG"), M, 24, 144, xform=1)
