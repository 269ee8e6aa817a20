"""Fixtures that hold OUTPUTS OF THE ORACLE (not of the reference): they exist because the oracle needs minutes for them and the GPU
box should spend its time on the GPU.  Each one is re-derived here, in part, so that the file cannot drift from oracle/."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_benchmark_shape_ais_fixture_is_the_float64_oracles():
    """tests/golden/ais_benchmark_shape_oracle.json: chains are independent (Philox counters carry the chain index), so two chains
    from the middle of the ladder, recomputed alone, must reproduce the stored float64 log-weights bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_ais_benchmark_oracle', os.path.join(HERE, 'golden', 'make_ais_benchmark_oracle.py'))
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)
    fx = json.load(open(os.path.join(HERE, 'golden', 'ais_benchmark_shape_oracle.json')))
    assert fx['shape'] == {k: (list(v) if isinstance(v, tuple) else v) for k, v in M.SHAPE.items()}
    assert len(fx['log_weights']) == fx['n_runs'] == 256
    got = M.log_weights(2, first_run=200)
    np.testing.assert_array_equal(got, np.asarray(fx['log_weights'][200:202]))
    # and the parameters are the ones the GPU test hands to the engines
    import sys
    sys.path.insert(0, HERE)
    d = M.weights()
    rng = np.random.RandomState(0)
    assert np.array_equal(d['vb'], (0.1 * rng.randn(784)).astype(np.float32))
