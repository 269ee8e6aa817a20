"""Prints the in-kernel timeline (globaltimer) of CTA 0 for a few GEMM shapes/tile configs.
BM_TC_TIMELINE=1 [BM_TC_REPS=n] python tools/tc_timeline.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boltzmann-machines_b200'))
import numpy as np
from boltzmann_machines import _native
rng = np.random.RandomState(0)
for (M, N, K, b_t, bn, c) in [(4096, 1024, 784, True, 256, 2), (4096, 784, 1024, False, 208, 2), (4096, 1024, 784, True, 256, 1)]:
    A = rng.rand(M, K).astype(np.float32)
    B = (rng.randn(K, N) if b_t else rng.randn(N, K)).astype(np.float32)
    for rep in range(2):
        sys.stderr.write("M%d N%d K%d b_t=%s bn=%d c=%d rep%d\n" % (M, N, K, b_t, bn, c, rep))
        C = _native.debug_tc_gemm(A, B, b_t=b_t, force_bn=bn, force_cluster=c)
