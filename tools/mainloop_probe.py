"""Steady-state mainloop period of the tensor-core kernel for one big GEMM (CTA 0 timeline).
BM_TC_TIMELINE=1 BM_TC_REPS=2 python tools/mainloop_probe.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boltzmann-machines_b200'))
import numpy as np
from boltzmann_machines import _native
rng = np.random.RandomState(0)
M, K = 4096, 1536
cases = [(1024, False, 256, 2), (1024, False, 128, 2), (1024, False, 64, 2), (1024, True, 256, 2), (1024, True, 128, 2),
         (1024, False, 256, 1), (1024, False, 128, 1)]
for (N, b_t, bn, c) in cases:
    A = rng.rand(M, K).astype(np.float32)
    B = (rng.randn(K, N) if b_t else rng.randn(N, K)).astype(np.float32)
    sys.stderr.write("== M%d N%d K%d b_t=%s bn=%d cluster=%d\n" % (M, N, K, b_t, bn, c))
    _native.debug_tc_gemm(A, B, b_t=b_t, force_bn=bn, force_cluster=c)
