"""Host-side wall time of bm_rbm_train_epoch[_u8] calls on the bench workload (diagnostic)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boltzmann-machines_b200'))
import numpy as np
import bench
from boltzmann_machines import _native
B, NB = bench.B, 20
X = bench.synth_mnist(B * NB)
eng = _native.CudaRBM(bench.model_cfg('bf16'))
eng.init_normal_W(0.01, 1337)
ctx = _native.Context.default()
for name in ('u8', 'f32', 'u8', 'f32'):
    Xh = eng.pin(X) if name == 'u8' else _native.pinned_copy(X)
    ts = []
    for rep in range(4):
        ctx.sync(); t0 = time.perf_counter()
        eng.train_epoch(Xh, B, bench.LR, bench.MOMENTUM, bench.K_GIBBS, 7, rep * NB, metrics=('msre',), every=1)
        ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3 / NB)
    print(name, Xh.dtype, 'ms/step per epoch call:', ['%.3f' % t for t in ts], flush=True)
    _native.pinned_free(Xh)
