"""Epilogue duration of one 256x256 pair tile (CTA 0 timeline) for the hot output modes.
BM_TC_TIMELINE=1 BM_TC_REPS=1 BM_TC_EPI=1|2|3 [BM_TC_PROBE=1|2|3] python tools/epilogue_probe.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boltzmann-machines_b200'))
import numpy as np
from boltzmann_machines import _native
rng = np.random.RandomState(0)
M, N, K = 4096, 1024, 784
A = rng.rand(M, K).astype(np.float32)
B = rng.randn(N, K).astype(np.float32)
_native.debug_tc_gemm(A, B, b_t=False, force_bn=256, force_cluster=2)
