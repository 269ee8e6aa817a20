"""BM_TC_PROGRAM_TIMELINE=n python tools/program_timeline.py : per-unit timeline of CTA 0 for the first n program launches."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boltzmann-machines_b200'))
import numpy as np
import bench
from boltzmann_machines import _native
X = bench.synth_mnist(bench.B * 4)
eng = _native.CudaRBM(bench.model_cfg('bf16'))
eng.init_normal_W(0.01, 1337)
eng.set_data(X)
for i in range(4):
    eng.train_step_at((i % 4) * bench.B, bench.B, bench.LR, bench.MOMENTUM, bench.K_GIBBS, 1, i)
_native.Context.default().sync()
