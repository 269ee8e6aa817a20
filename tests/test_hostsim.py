"""The HOST side of libbm.so executed without a GPU: the library's own objects linked against a stand-in CUDA runtime
(tests/hostsim/fake_cudart.cpp: device memory is host memory, kernel launches are recorded and skipped).

Every entry point of the RBM and DBM engines -- including the tensor-core DBM engine and its persistent-program variants,
which have not run on a B200 yet -- is driven through the C-ABI on benchmark-sized and ragged shapes.  What this checks:
every BM_REQUIRE on the way, every tensor map the host encodes (validated like the driver does, and its whole view must
lie inside one device allocation), every memcpy / memset range against the allocation it touches, launch configurations,
program construction (dependencies, op counts).  What it cannot check: anything a kernel computes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libbm_hostsim.so')
OBJ = os.path.join(ROOT, 'boltzmann-machines_b200', 'build')


@pytest.fixture(scope='module')
def sim():
    if not (os.path.isdir(OBJ) and any(f.endswith('.o') for f in os.listdir(OBJ))):
        pytest.skip('library objects not built (run build.sh / __graft_entry__.build())')
    subprocess.check_call(['bash', os.path.join(ROOT, 'tests', 'hostsim', 'build.sh')], stdout=subprocess.DEVNULL)
    from boltzmann_machines import _native
    lib = _native.load_library(SIM)
    lib.fakecuda_violation.restype = C.c_char_p
    old_lib, old_ctx = _native._lib, dict(_native.Context._default)
    _native._lib = lib
    _native.Context._default.clear()
    yield lib
    _native.Context._default.clear()
    _native.Context._default.update(old_ctx)
    _native._lib = old_lib


def clean(lib):
    v = lib.fakecuda_violation().decode()
    assert v == '', v


RBM_SHAPES = [(784, 1024, 4096, 5), (784, 16, 32, 1), (130, 70, 65, 2), (3072, 5000, 512, 1), (37, 29, 19, 3), (784, 4096, 300, 25)]


@pytest.mark.parametrize('V,H,B,k', RBM_SHAPES)
@pytest.mark.parametrize('kind', ['bernoulli', 'gaussian'])
@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_rbm_entry_points(sim, V, H, B, k, kind, compute):
    from boltzmann_machines import _native
    if compute == 'fp32' and V * H > 2000000:
        pytest.skip('same host path as the smaller shapes')
    cfg = dict(n_visible=V, n_hidden=H, dtype='float32', compute=compute, l2=1e-5, sample_v=False, sample_h=True, max_batch=B,
               v_kind=kind, h_kind='bernoulli', dropout=None if kind == 'bernoulli' else 0.9)
    if kind == 'gaussian':
        cfg['sigma'] = np.ones(V)
    rng = np.random.RandomState(0)
    X = (rng.rand(2 * B + 3, V) < 0.2).astype(np.float32)
    sim.fakecuda_reset()
    eng = _native.CudaRBM(cfg)
    eng.init_normal_W(0.01, 1337)
    eng.train_step(X[:B], 0.05, 0.5, k, 7, 0, metrics=('msre', 'pll', 'free_energy', 'l2_loss'))
    eng.train_step(X[:max(1, B // 3)], 0.05, 0.5, k, 7, 1)                     # a ragged last batch
    eng.train_epoch(X, B, 0.05, 0.5, k, 7, 2, metrics=('msre',), every=1)
    if kind == 'bernoulli':
        P = eng.pin(X)
        assert P.dtype == np.uint8
        eng.train_epoch(P, B, 0.05, 0.5, k, 7, 5, metrics=('msre',), every=2)
        eng.unpin(P)
    eng.set_data(X)
    eng.train_step_at(3, B, 0.05, 0.5, k, 7, 9)
    eng.transform(X[:B], k, 7, 10)
    eng.metrics(X[:B], k, 7, 11, ('msre', 'pll', 'free_energy'))
    for name in ('h0_means', 'v_means', 'h_means'):
        eng.get_activation(name, B)
    eng.get_params()
    clean(sim)
    if compute == 'bf16':
        assert sim.fakecuda_launches(b'tc_program_kernel') >= 5      # (tensor maps are cached per buffer: no lower bound)
    else:
        assert sim.fakecuda_launches(b'tc_program_kernel') == 0
    eng.close()


def dbm_cfg(V, Hs, B, M, compute, gaussian=False, max_mf=6):
    L = len(Hs)
    cfg = dict(n_visible=V, n_hiddens=list(Hs), v_kind='gaussian' if gaussian else 'bernoulli', h_kinds=['bernoulli'] * L,
               h_n_samples=[100.] * L, dtype='float32', compute=compute, n_particles=M, batch_size=B, max_mf_updates=max_mf,
               mf_tol=-1.0,                     # kernels do not run here: a negative tolerance walks every sweep of the loop
               l2=1e-4, max_norm=3.0, sample_v=True, sample_h=[True] * L, sparsity_target=[0.2] * L, sparsity_cost=[0.01] * L,
               sparsity_damping=0.9)
    if gaussian:
        cfg['sigma'] = np.ones(V)
    return cfg


DBM_SHAPES = [(784, (512, 1024), 1024, 1024, 25), (30, (18, 11), 10, 12, 6), (30, (18, 11, 7), 10, 12, 7), (130, (70,), 33, 65, 3),
              (784, (4096,), 300, 300, 2), (9, (5, 4), 1, 1, 4)]


@pytest.mark.parametrize('V,Hs,B,M,max_mf', DBM_SHAPES)
@pytest.mark.parametrize('mode', ['fp32', 'bf16', 'bf16-programs'])
def test_dbm_entry_points(sim, monkeypatch, V, Hs, B, M, max_mf, mode):
    from boltzmann_machines import _native
    compute = 'fp32' if mode == 'fp32' else 'bf16'
    monkeypatch.setenv('BM_DBM_MF_CHUNK', '0')
    monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '0')
    if mode == 'bf16-programs':
        monkeypatch.setenv('BM_DBM_MF_CHUNK', '4')
        monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '1')
    gaussian = (V == 130)
    cfg = dbm_cfg(V, Hs, B, M, compute, gaussian, max_mf)
    rng = np.random.RandomState(1)
    X = (rng.rand(B, V) < 0.3).astype(np.float32)
    sim.fakecuda_reset()
    eng = _native.CudaDBM(cfg)
    assert eng.compute == compute
    sizes = [V] + list(Hs)
    init = {'vb': np.zeros(V, np.float32)}
    for i in range(len(Hs)):
        s = '' if i == 0 else '_%d' % i
        init['W' + s] = (0.1 * rng.randn(sizes[i], sizes[i + 1])).astype(np.float32)
        init['hb' + s] = np.zeros(sizes[i + 1], np.float32)
    eng.set_params(init)
    eng.init_particles(4242)
    for k in (1, 3):
        got = eng.train_step(X, 0.01, 0.5, k, 99, k, metrics=('msre', 'n_mf_updates'))
        assert got['n_mf_updates'] == max_mf            # every sweep of the E-step was issued
    eng.train_step(X[:max(1, B // 2)], 0.01, 0.5, 1, 99, 5)
    eng.val_metrics(X, 2, 99, 6)
    assert eng.transform(X).shape == (B, Hs[-1])
    assert eng.reconstruct(X).shape == (B, V)
    assert eng.sample_v(2, 99, 7).shape == (M, V)
    if len(Hs) == 2 and not gaussian:
        assert eng.log_proba(X).shape == (B,)
        n_runs = 20000 if V == 784 else 13
        assert eng.ais(n_runs, 5, 2, 2222).shape == (n_runs,)
        assert eng.ais(7, 5, 1, 2222, first_run=6).shape == (7,)
    eng.get_params()
    clean(sim)
    n_prog = sim.fakecuda_launches(b'tc_program_kernel')
    if mode == 'fp32':
        assert n_prog == 0
    else:
        assert n_prog > 0
    eng.close()


def test_the_simulation_catches_a_tensor_view_past_its_allocation(sim):
    """Self-test of the checker: a tensor map whose view is larger than the allocation behind it must be refused."""
    from boltzmann_machines import _native
    lib = sim
    sim.fakecuda_reset()
    ctx = _native.Context.default()
    A = np.ones((64, 64), np.float32)
    # the debug GEMM hook allocates exactly M x K; ask it for a healthy product first
    out = _native.debug_tc_gemm(A, A, a_t=False, b_t=False)
    assert out.shape == (64, 64)
    clean(sim)
    # now corrupt: call the raw C entry with K larger than what the host buffers were sized for is not possible through the
    # binding (it sizes the buffers itself), so check the primitive directly: free of an unknown pointer is reported
    rc = C.CDLL(SIM).cudaFree(C.c_void_p(12345678))
    assert rc != 0 and b'not a live allocation' in lib.fakecuda_violation()
    lib.fakecuda_reset()


# ---------------------------------------------------------------------------------------------------------------------
# Numerics of the tensor-core DBM engine with its kernels INTERPRETED on the CPU (tests/hostsim/kernels_cpu.cpp: the program
# kernel from its descriptors, the small kernels restated from their sources) against the bf16 oracle emulation.  This is the
# check of the engine's wiring -- operand orientations, scales, sites, buffers, program dependencies -- that does not need a GPU.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def executing(sim):
    sim.fakecuda_skipped.restype = C.c_char_p
    sim.fakecuda_reset()
    sim.fakecuda_set_execute(1)
    yield sim
    sim.fakecuda_set_execute(0)


def tc_pair(cfg, scale=0.3, seed=0):
    from boltzmann_machines import _native
    from oracle.dbm_bf16 import OracleDBMbf16
    eng, emu = _native.CudaDBM(cfg), OracleDBMbf16(cfg)
    assert eng.compute == 'bf16'
    rng = np.random.RandomState(seed)
    sizes = [cfg['n_visible']] + cfg['n_hiddens']
    d = {'vb': (0.1 * rng.randn(sizes[0])).astype(np.float32)}
    for i in range(len(cfg['n_hiddens'])):
        s = '' if i == 0 else '_%d' % i
        d['W' + s] = (scale * rng.randn(sizes[i], sizes[i + 1])).astype(np.float32)
        d['hb' + s] = (0.1 * rng.randn(sizes[i + 1])).astype(np.float32)
    for e in (eng, emu):
        e.set_params(d)
        e.init_particles(4242)
    return eng, emu


def small_cfg(Hs=(18, 11), V=30, gaussian=False, **kw):
    cfg = dbm_cfg(V, Hs, 10, 12, 'bf16', gaussian, 6)
    cfg['mf_tol'] = 1e-6
    if gaussian:
        cfg['sigma'] = np.linspace(0.7, 1.3, V)
    cfg.update(kw)
    return cfg


def no_skips(sim):
    clean(sim)
    assert sim.fakecuda_skipped() == b'', sim.fakecuda_skipped()


@pytest.mark.parametrize('Hs', [(18,), (18, 11), (18, 11, 7), (70, 130)])
@pytest.mark.parametrize('programs', [False, True])
@pytest.mark.parametrize('mixed', ['0', '1'])
def test_tc_dbm_queries_equal_the_bf16_emulation(executing, monkeypatch, Hs, programs, mixed):
    monkeypatch.setenv('BM_DBM_TC_MIXED', mixed)        # 0: transposed second shadow (default); 1: one op, two B layouts
    monkeypatch.setenv('BM_DBM_MF_CHUNK', '0')
    monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '0')
    if programs:
        monkeypatch.setenv('BM_DBM_MF_CHUNK', '4')
        monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '1')
    cfg = small_cfg(Hs)
    eng, emu = tc_pair(cfg)
    g, w = eng.get_params(), emu.get_params()
    for k in ('v', 'h'):
        np.testing.assert_array_equal(g[k], w[k], err_msg='initial ' + k)
    X = (np.random.RandomState(9).rand(7, cfg['n_visible']) < 0.3).astype(np.float32)
    tol = dict(rtol=2.0 ** -7, atol=1e-6)
    np.testing.assert_allclose(eng.transform(X), emu.transform(X), err_msg='transform', **tol)
    np.testing.assert_allclose(eng.reconstruct(X), emu.reconstruct(X), err_msg='reconstruct', **tol)
    if len(Hs) == 2:
        np.testing.assert_allclose(eng.log_proba(X), emu.log_proba(X), rtol=1e-3, atol=2e-2)
    a, b = eng.val_metrics(X, 2, 7, 3), emu.val_metrics(X, 2, 7, 3)
    assert a['n_mf_updates'] == b['n_mf_updates'] and a['msre'] == pytest.approx(b['msre'], rel=1e-3)
    np.testing.assert_allclose(eng.sample_v(3, 11, 4), emu.sample_v(3, 11, 4), err_msg='sample_v', **tol)
    g, w = eng.get_params(), emu.get_params()
    for k in w:
        if k.startswith(('v', 'h', 'mu')) and not k.startswith(('hb', 'vb', 'mu_means')):
            np.testing.assert_allclose(g[k], w[k], err_msg=k, **tol)
    no_skips(executing)
    eng.close()


@pytest.mark.parametrize('Hs,gaussian', [((18,), False), ((18, 11), False), ((18, 11, 7), False), ((18, 11), True)])
@pytest.mark.parametrize('programs', [False, True])
@pytest.mark.parametrize('mixed', ['0', '1'])
def test_tc_dbm_training_steps_equal_the_bf16_emulation(executing, monkeypatch, Hs, gaussian, programs, mixed):
    monkeypatch.setenv('BM_DBM_TC_MIXED', mixed)
    monkeypatch.setenv('BM_DBM_MF_CHUNK', '0')
    monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '0')
    if programs:
        monkeypatch.setenv('BM_DBM_MF_CHUNK', '3')
        monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '1')
    cfg = small_cfg(Hs, gaussian=gaussian)
    eng, emu = tc_pair(cfg)
    rng = np.random.RandomState(5)
    for it in range(3):
        X = rng.randn(10, cfg['n_visible']).astype(np.float32) if gaussian else (rng.rand(10, cfg['n_visible']) < 0.3).astype(np.float32)
        a = eng.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        b = emu.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        assert a['n_mf_updates'] == b['n_mf_updates'], it
        assert a['msre'] == pytest.approx(b['msre'], rel=1e-3), it
    g, w = eng.get_params(), emu.get_params()
    for k in w:
        np.testing.assert_allclose(g[k], w[k], rtol=2.0 ** -7, atol=2e-5, err_msg=k)
    no_skips(executing)
    eng.close()


@pytest.mark.parametrize('k,fused', [(1, 'epilogue'), (3, 'epilogue'), (1, '1'), (1, '0'), (3, '1')])
@pytest.mark.parametrize('mixed', ['0', '1'])
def test_tc_dbm_ais_equals_the_bf16_emulation(executing, monkeypatch, k, fused, mixed):
    monkeypatch.setenv('BM_DBM_TC_MIXED', mixed)
    monkeypatch.setenv('BM_DBM_AIS_EPILOGUE', '1' if fused == 'epilogue' else '0')
    monkeypatch.setenv('BM_DBM_AIS_FUSED', '0' if fused == '0' else '1')
    cfg = small_cfg((5, 4), V=7, n_particles=4, batch_size=4)
    eng, emu = tc_pair(cfg)
    a = eng.ais(16, 60, k, 2222)
    b = emu.ais(16, 60, k, 2222)
    np.testing.assert_allclose(a, b, rtol=0, atol=2e-4)                 # same chains: a different one would differ by ~0.1
    np.testing.assert_allclose(eng.ais(5, 60, k, 2222, first_run=9), b[9:14], rtol=0, atol=2e-4)
    no_skips(executing)
    eng.close()


@pytest.mark.parametrize('kind', ['bernoulli', 'gaussian'])
@pytest.mark.parametrize('V,H,B,k', [(37, 29, 19, 2), (130, 72, 65, 1), (784, 16, 32, 3)])
def test_the_interpreter_reproduces_the_gpu_verified_rbm_program(executing, kind, V, H, B, k):
    """Credential of the interpreter: the RBM tensor-core path is verified on the GPU against the bf16-rounding oracle
    (tests/test_tc_gpu.py); run through the interpreter instead of the GPU it must land on the same oracle -- whole CD-k
    programs with their split-K dW ops, spare-lane scheduling and both weight-update variants included."""
    from boltzmann_machines import _native
    from oracle.rbm import OracleRBM
    cfg = dict(n_visible=V, n_hidden=H, dtype='float32', compute='bf16', l2=1e-4, max_batch=B, sample_v=False, sample_h=True,
               sparsity_cost=0.01, sparsity_target=0.2, v_kind=kind, h_kind='bernoulli', dropout=0.9)
    if kind == 'gaussian':
        cfg['sigma'] = np.linspace(0.5, 1.5, V)
    rng = np.random.RandomState(0)
    init = dict(W=(0.1 * rng.randn(V, H)).astype(np.float32), vb=(0.1 * rng.randn(V)).astype(np.float32), hb=(0.1 * rng.randn(H)).astype(np.float32))
    eng, ora = _native.CudaRBM(cfg), OracleRBM(cfg)
    eng.set_params(init), ora.set_params(init)
    for it in range(3):
        X = rng.randn(B, V).astype(np.float32) if kind == 'gaussian' else (rng.rand(B, V) < 0.3).astype(np.float32)
        a = eng.train_step(X, 0.05, 0.5, k, 0xABCDEF, it, metrics=('msre',))
        b = ora.train_step(X, 0.05, 0.5, k, 0xABCDEF, it, metrics=('msre',))
        assert a['msre'] == pytest.approx(b['msre'], rel=2e-3)
    g, w = eng.get_params(), ora.get_params()
    for name in ('W', 'vb', 'hb', 'dW', 'dvb', 'dhb', 'q_means'):
        np.testing.assert_allclose(g[name], w[name], atol=2e-3, err_msg=name)
    np.testing.assert_allclose(eng.transform(X, k, 5, 9), ora.transform(X, k, 5, 9), atol=2e-2)
    no_skips(executing)
    eng.close()


@pytest.mark.parametrize('Hs,gaussian', [((18,), False), ((18, 11), False), ((18, 11, 7), False), ((18, 11), True)])
def test_fp32_dbm_engine_equals_the_pinned_oracle_under_the_interpreter(executing, Hs, gaussian):
    """The default (storage-precision) DBM engine, whose AIS was re-cut into slices and whose step became shard-aware in this
    round, against oracle/dbm.py with its kernels interpreted on the CPU -- the same comparison tests/test_dbm_gpu.py makes
    on the GPU, available before the GPU is."""
    from boltzmann_machines import _native
    from oracle.dbm import OracleDBM
    cfg = small_cfg(Hs, gaussian=gaussian, compute='fp32')
    eng, ora = _native.CudaDBM(cfg), OracleDBM(cfg)
    assert eng.compute == 'fp32'
    rng = np.random.RandomState(0)
    sizes = [cfg['n_visible']] + cfg['n_hiddens']
    d = {'vb': (0.1 * rng.randn(sizes[0])).astype(np.float32)}
    for i in range(len(Hs)):
        s = '' if i == 0 else '_%d' % i
        d['W' + s] = (0.3 * rng.randn(sizes[i], sizes[i + 1])).astype(np.float32)
        d['hb' + s] = (0.1 * rng.randn(sizes[i + 1])).astype(np.float32)
    for e in (eng, ora):
        e.set_params(d)
        e.init_particles(4242)
    for it in range(3):
        X = rng.randn(10, cfg['n_visible']).astype(np.float32) if gaussian else (rng.rand(10, cfg['n_visible']) < 0.3).astype(np.float32)
        a = eng.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        b = ora.train_step(X, 0.05, 0.5, 2, 99, it, metrics=('msre', 'n_mf_updates'))
        assert a['n_mf_updates'] == b['n_mf_updates'] and a['msre'] == pytest.approx(b['msre'], rel=1e-4)
    g, w = eng.get_params(), ora.get_params()
    for k in w:
        np.testing.assert_allclose(g[k], w[k], atol=5e-5, err_msg=k)
    Xq = X[:7]
    np.testing.assert_allclose(eng.transform(Xq), ora.transform(Xq), atol=2e-5)
    np.testing.assert_allclose(eng.reconstruct(Xq), ora.reconstruct(Xq), atol=2e-5)
    np.testing.assert_allclose(eng.sample_v(3, 11, 4), ora.sample_v(3, 11, 4), atol=2e-5)
    if len(Hs) == 2 and not gaussian:
        np.testing.assert_allclose(eng.log_proba(Xq), ora.log_proba(Xq), rtol=1e-5, atol=1e-4)
        full = eng.ais(13, 40, 2, 99)
        np.testing.assert_allclose(full, ora.ais(13, 40, 2, 99), atol=5e-3)
        parts = np.concatenate([eng.ais(5, 40, 2, 99, first_run=0), eng.ais(1, 40, 2, 99, first_run=5), eng.ais(7, 40, 2, 99, first_run=6)])
        np.testing.assert_allclose(full, parts, rtol=0, atol=1e-6)
    no_skips(executing)
    eng.close()


def test_boundary_refuses_bad_arguments_without_crashing(sim):
    """Error behaviour of the C-ABI (include/bm.h: every function returns BM_OK or a negative code and bm_last_error() says
    why): wrong sizes, unknown names, out-of-range rows and unsupported models are refused with a message, and the handle
    stays usable."""
    from boltzmann_machines import _native
    sim.fakecuda_reset()
    cfg = dict(n_visible=12, n_hidden=8, dtype='float32', compute='bf16', max_batch=4)
    eng = _native.CudaRBM(cfg)
    with pytest.raises(RuntimeError, match='unknown|size'):
        _native.check(sim.bm_rbm_set_param(eng.handle, b'nope', np.zeros(3, np.float32).ctypes.data, 12))
    with pytest.raises(RuntimeError, match='size'):
        _native.check(sim.bm_rbm_set_param(eng.handle, b'W', np.zeros(5, np.float32).ctypes.data, 20))
    with pytest.raises(ValueError):
        eng.train_step(np.zeros((4, 11), np.float32), 0.1, 0.5, 1, 1, 0)           # wrong width: refused by the binding
    with pytest.raises(RuntimeError):
        eng.train_step(np.zeros((4, 12), np.float32), 0.1, 0.5, 0, 1, 0)           # k = 0
    with pytest.raises(RuntimeError, match='resident|range|data'):
        eng.train_step_at(0, 4, 0.1, 0.5, 1, 1, 0)                                 # no resident dataset
    eng.train_step(np.zeros((4, 12), np.float32), 0.1, 0.5, 1, 1, 0)               # still usable
    eng.close()
    dcfg = dbm_cfg(9, (5, 4, 3), 4, 4, 'fp32')
    dbm = _native.CudaDBM(dcfg)
    with pytest.raises(RuntimeError, match='batch_size'):
        dbm.train_step(np.zeros((5, 9), np.float32), 0.1, 0.5, 1, 1, 0)            # more rows than batch_size
    with pytest.raises(RuntimeError, match='2-layer|2 hidden'):
        dbm.ais(4, 5, 1, 1)                                                        # AIS is defined for two hidden layers
    with pytest.raises(RuntimeError, match='2 hidden'):
        dbm.log_proba(np.zeros((4, 9), np.float32))
    with pytest.raises(RuntimeError, match='unknown'):
        _native.check(sim.bm_dbm_get_param(dbm.handle, b'W_7', np.zeros(4, np.float32).ctypes.data, 16))
    dbm.train_step(np.zeros((4, 9), np.float32), 0.1, 0.5, 1, 1, 0)
    dbm.close()
    bad = dict(dbm_cfg(9, (5, 4), 4, 4, 'fp32'), h_kinds=['bernoulli', 'gaussian'])
    with pytest.raises((RuntimeError, KeyError)):
        d2 = _native.CudaDBM(bad)
        d2.train_step(np.zeros((4, 9), np.float32), 0.1, 0.5, 1, 1, 0)             # gaussian hidden layers are not supported
    clean(sim)


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_long_ais_ladders_go_in_chunks(sim, compute):
    """More runs than grid.y can index (65535): bm_dbm_ais cuts the ladder into chunks of 32768 runs; launch configurations
    stay legal (kernels not interpreted here: 70000 runs)."""
    from boltzmann_machines import _native
    sim.fakecuda_reset()
    eng = _native.CudaDBM(dbm_cfg(9, (5, 4), 4, 4, compute))
    out = eng.ais(70000, 3, 1, 5)
    assert out.shape == (70000,)
    clean(sim)
    eng.close()


def test_chunked_ais_equals_the_unchunked_ladder(executing, monkeypatch):
    """Same chains whatever the chunking: a 13-run ladder against its slices was checked above; here the engine's own
    chunk boundary is crossed on purpose by comparing runs [32760, 32776) computed inside one 40000-run call with the same
    runs computed alone."""
    from boltzmann_machines import _native
    cfg = small_cfg((5, 4), V=7, n_particles=4, batch_size=4, compute='fp32')
    eng = _native.CudaDBM(cfg)
    rng = np.random.RandomState(0)
    eng.set_params({'W': (0.3 * rng.randn(7, 5)).astype(np.float32), 'W_1': (0.3 * rng.randn(5, 4)).astype(np.float32)})
    alone = eng.ais(16, 6, 1, 77, first_run=32760)
    inside = eng.ais(40000, 6, 1, 77)[32760:32776]
    np.testing.assert_allclose(inside, alone, rtol=0, atol=1e-9)
    eng.close()


def test_tall_batches_and_particle_sets_draw_by_global_row(executing):
    """Batches / particle sets with more rows than grid.y can index go in slabs of 32768 rows; a row's draws depend on its
    index in the whole batch: rows around the slab boundary equal the same rows of a reference computed row by row."""
    from boltzmann_machines import _native
    from oracle import philox as P
    n, V = 33000, 8
    dbm = _native.CudaDBM(dbm_cfg(V, (4,), 4, n, 'fp32'))
    dbm.init_particles(4242)
    v = dbm.get_params(['v'])['v']
    want = P.uniform_at(n, V, 4242, P.SITE_PARTICLE_INIT, 0, 0)
    np.testing.assert_array_equal(v[32760:32776], want[32760:32776])
    np.testing.assert_array_equal(v[:8], want[:8])
    dbm.close()
    # dropout mask of a tall batch: the prepared input is X / keep * floor(keep + u(row, col))
    cfg = dict(n_visible=V, n_hidden=4, dtype='float32', compute='fp32', max_batch=n, dropout=0.7, sample_v=False, sample_h=False)
    rbm = _native.CudaRBM(cfg)
    X = np.ones((n, V), np.float32)
    rbm.train_step(X, 0.0, 0.0, 1, 99, 3)
    Xp = rbm.get_activation('X', n)
    u = P.uniform_at(n, V, 99, P.SITE_DROPOUT, 0, 3)
    want = np.float32(1.0) / np.float32(0.7) * np.floor(np.float32(0.7) + u)
    np.testing.assert_allclose(Xp[32760:32776], want[32760:32776], rtol=1e-6)
    rbm.close()
    no_skips(executing)


def test_structural_cost_of_the_hot_entry_points(sim):
    """What a step costs in runtime calls, counted on the stand-in runtime at the benchmark's shape (kernels not interpreted):
    the whole-epoch entry point blocks the host ONCE per epoch whatever the number of batches, moves exactly one byte per
    visible unit and row to the device and 64 bytes per step back, and launches 4 kernels per batch (program, update tail, MSRE on the
    compute stream; the byte -> bf16 conversion on the copy stream); a dataset-resident step is 2
    launches, no copy and no host synchronisation -- bench.py's `gpu_launches`, `h2d_bytes_per_step`, `d2h_bytes_per_step`."""
    from boltzmann_machines import _native
    V, H, B = 784, 1024, 4096
    cfg = dict(n_visible=V, n_hidden=H, dtype='float32', compute='bf16', l2=1e-5, sample_v=False, sample_h=True, max_batch=B)
    eng = _native.CudaRBM(cfg)
    eng.init_normal_W(0.01, 1)
    X = (np.random.RandomState(0).rand(6 * B, V) < 0.13).astype(np.float32)
    P = eng.pin(X)
    assert P.dtype == np.uint8
    eng.train_epoch(P, B, 0.05, 0.5, 5, 1, 0, metrics=('msre',), every=1)            # first call: programs, staging buffers
    for nb in (3, 6):
        sim.fakecuda_reset()
        eng.train_epoch(P[:nb * B], B, 0.05, 0.5, 5, 1, 100, metrics=('msre',), every=1)
        assert sim.fakecuda_syncs() == 1, (nb, sim.fakecuda_syncs())
        assert sim.fakecuda_launches(b'') == 4 * nb
        assert sim.fakecuda_launches(b'tc_program_kernel') == nb
        assert sim.fakecuda_h2d_bytes() == nb * B * V and sim.fakecuda_d2h_bytes() == 64 * nb
    eng.unpin(P)
    eng.set_data(X)
    eng.train_step_at(0, B, 0.05, 0.5, 5, 1, 0)                                      # first resident step builds its program
    sim.fakecuda_reset()
    for i in range(6):
        eng.train_step_at(i * B, B, 0.05, 0.5, 5, 1, 1 + i)
    assert sim.fakecuda_launches(b'') == 2 * 6 and sim.fakecuda_syncs() == 0 and sim.fakecuda_h2d_bytes() == 0
    clean(sim)
    eng.close()


def test_random_tc_dbm_configurations_equal_the_bf16_emulation(executing, monkeypatch):
    """tools/fuzz_dbm_tc_hostsim.py over a fixed range of seeds: ragged widths, 1-3 hidden layers, Gaussian visibles, every
    combination of the engine's switches; one training step at a time from identical states, then queries and AIS."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('fuzz_dbm_tc_hostsim', os.path.join(ROOT, 'tools', 'fuzz_dbm_tc_hostsim.py'))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    for k in ('BM_DBM_TC_MIXED', 'BM_DBM_AIS_FUSED', 'BM_DBM_MF_CHUNK', 'BM_DBM_PCD_PROGRAM'):
        monkeypatch.setenv(k, '0')                       # restored afterwards (the tool sets them per configuration)
    outcomes = {}
    for seed in range(900000, 900150):
        cfg, env, run = fuzz.draw(np.random.RandomState(seed))
        r = fuzz.one(cfg, env, run, seed, executing)
        outcomes[r] = outcomes.get(r, 0) + 1
    assert outcomes.get("ok", 0) >= 145 and not outcomes.get("FAIL"), outcomes   # (mf-count, sample-flip: rounding events, see the tool)
    assert executing.fakecuda_hazard_launches() > 100          # the programs' dataflow was checked on the way


def test_program_dataflow_hazards_are_checked_and_a_missing_dependency_is_caught(executing, monkeypatch):
    """On the GPU the units of a program run concurrently, ordered only by the dependencies the host declares; the interpreter
    runs the ops in program order, so it shadows every element a launch touches and reports reads / overwrites no dependency
    path covers (row block by row block).  Clean: the GPU-verified RBM step program and the tensor-core DBM engine's mean-field
    and particle programs on three 256-row blocks.  The checker itself: with any one needed dependency taken away it speaks up."""
    from boltzmann_machines import _native
    executing.fakecuda_hazard_launches.restype = C.c_long
    rng = np.random.RandomState(0)

    def rbm_step(drop=None):
        executing.fakecuda_reset()
        eng = _native.CudaRBM(dict(n_visible=70, n_hidden=40, dtype='float32', compute='bf16', l2=1e-5, sample_v=False, sample_h=True,
                                   max_batch=600))
        eng.init_normal_W(0.01, 1)
        X = (rng.rand(600, 70) < 0.3).astype(np.float32)
        if drop:
            executing.fakecuda_drop_dependency(*drop)
        before = executing.fakecuda_hazard_launches()
        eng.train_step(X, 0.05, 0.5, 2, 7, 0)
        executing.fakecuda_drop_dependency(-1, -1)
        checked, v = executing.fakecuda_hazard_launches() - before, executing.fakecuda_violation().decode()
        eng.close()
        return checked, v

    def dbm_step(Hs, drop=None):
        executing.fakecuda_reset()
        cfg = dbm_cfg(70, Hs, 600, 520, 'bf16', False, 6)
        cfg['mf_tol'] = 1e-6
        eng = _native.CudaDBM(cfg)
        eng.init_particles(1)
        X = (rng.rand(600, 70) < 0.3).astype(np.float32)
        if drop:
            executing.fakecuda_drop_dependency(*drop)
        before = executing.fakecuda_hazard_launches()
        eng.train_step(X, 0.05, 0.5, 2, 7, 0)
        executing.fakecuda_drop_dependency(-1, -1)
        checked, v = executing.fakecuda_hazard_launches() - before, executing.fakecuda_violation().decode()
        eng.close()
        return checked, v

    checked, v = rbm_step()
    assert checked == 1 and v == '', v
    # ops of the step program: h0, sum_rows X (no dependency), positive dW, v1, h1, v2, h2, sum_rows v_k, negative dW
    for op in (2, 3, 4, 5, 6, 7):
        _, v = rbm_step(drop=(op, 0))
        assert 'without a dependency path' in v, (op, v)
    monkeypatch.setenv('BM_DBM_MF_CHUNK', '3')
    monkeypatch.setenv('BM_DBM_PCD_PROGRAM', '1')
    for mixed in ('0', '1'):
        monkeypatch.setenv('BM_DBM_TC_MIXED', mixed)
        for Hs in ((40,), (40, 24), (40, 24, 16)):
            checked, v = dbm_step(Hs)
            assert checked >= 2 and v == '', (Hs, mixed, v)
    for op in (1, 2, 3):
        _, v = dbm_step((40, 24), drop=(op, 0))
        assert 'without a dependency path' in v, (op, v)
    # the kernel's wait loop skips the unit-level wait of every same-row-block dependency of an op that follows its producer
    # granule by granule: launch_tc_program only lets an op do so when that producer is its single such dependency
    assert executing.fakecuda_unhonoured_dependencies() == 0
    executing.fakecuda_reset()


@pytest.mark.parametrize('selection', [
    ['tests/test_plugin_gpu.py'],
    ['tests/test_tc_gpu.py', '-k', 'multinomial and not 784'],
    ['tests/test_tc_gpu.py', '-k', 'longer_than_one_program'],
    ['tests/test_rbm_gpu.py', '-k', 'bfloat16_feed or byte_valued or fit_takes'],
])
def test_gpu_tests_of_round_two_features_pass_on_the_interpreter(selection):
    """The GPU tests written in round 2 -- user-defined layers on the plugin engine, multinomial layers on the tensor-core path,
    chains cut into several program launches, the packed feeds -- dry-run here through the library's own host code with interpreted
    kernels (BM_HOSTSIM=1, see conftest.py): their host side is covered by the CPU suite, their kernels by the B200 runs."""
    import subprocess
    import sys
    env = dict(os.environ, BM_HOSTSIM='1')
    res = subprocess.run([sys.executable, '-m', 'pytest', '-m', 'gpu', '-q', '-x', '-p', 'no:cacheprovider'] + selection,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-3000:]
    assert 'runtime violations: none' in res.stdout and 'kernels without a CPU restatement: none' in res.stdout, res.stdout[-1500:]


def test_a_device_that_cannot_hold_every_cluster_of_a_program_is_refused(sim):
    """The dataflow programs' CTAs wait for each other: launching one on a device that keeps fewer clusters resident than the grid
    has (MPS / MIG share, other work on the SMs) must be an error at the boundary, not a kernel that traps after its 4-second
    guard.  Fresh process: the runtime is asked once per device."""
    import sys
    import textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path[:0] = [%r, %r]
        from boltzmann_machines import _native
        lib = _native.load_library(%r)
        _native._lib = lib
        lib.fakecuda_set_max_active_clusters(10)
        eng = _native.CudaRBM(dict(n_visible=64, n_hidden=64, compute='bf16', max_batch=256))
        eng.set_params({'W': np.zeros((64, 64), np.float32)})
        try:
            eng.train_step((np.random.rand(256, 64) < 0.5).astype(np.float32), 0.1, 0.5, 1, 1, 0)
        except RuntimeError as e:
            print('REFUSED:', e)
        else:
            print('LAUNCHED')
    ''') % (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200'), os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libbm_hostsim.so'))
    res = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert 'REFUSED:' in res.stdout and 'cannot hold all CTA clusters' in res.stdout, res.stdout[-2000:]


def _stream_order_probe(code_body):
    """runs `code_body` in a fresh process on the stand-in runtime with launches interpreted and the stream-order check on; returns
    the process's stdout"""
    import sys
    import textwrap
    head = textwrap.dedent('''
        import ctypes as C, sys, numpy as np
        sys.path[:0] = [%r, %r]
        from boltzmann_machines import _native
        lib = _native.load_library(%r)
        lib.fakecuda_violation.restype = C.c_char_p
        _native._lib = lib
        lib.fakecuda_set_execute(1)
        lib.fakecuda_set_stream_order_check(1)
    ''') % (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200'), SIM)
    res = subprocess.run([sys.executable, '-c', head + textwrap.dedent(code_body)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    return res.stdout


EPOCH_PROBE = '''
    V, H, B = 100, 64, 64            # (a half of the double-buffered operand = 64 rows x 192 columns x 2 bytes = 6 whole pages)
    rng = np.random.RandomState(0)
    eng = _native.CudaRBM(dict(n_visible=V, n_hidden=H, dtype='float32', compute=%r, l2=1e-4, max_batch=B, **%r))
    eng.init_normal_W(0.01, 1)
    X = %s
    P = eng.pin(X)
    for it in range(2):
        eng.train_epoch(P, B, 0.05, 0.5, 2, 1, 10 * it, metrics=('msre',), every=1)
    eng.unpin(P)
    eng.close()
    print('OPS', lib.fakecuda_stream_order_ops(), 'RACES', lib.fakecuda_stream_races())
    print('VIOLATION:', lib.fakecuda_violation().decode())
'''
BYTES = "(rng.rand(7 * 64 + 10, 100) < 0.3).astype(np.float32)"
REALS = "rng.rand(7 * 64 + 10, 100).astype(np.float32)"


@pytest.mark.parametrize('compute,data,extra', [
    ('bf16', BYTES, {}), ('bf16', REALS, {}), ('fp32', BYTES, {}), ('fp32', REALS, {}),
    ('bf16', REALS, dict(v_kind='gaussian', sigma=1.0)),                 # not `plain`: fp32 staging + prepare_input on the compute stream
    ('bf16', BYTES, dict(dropout=0.8)), ('bf16', BYTES, dict(h_kind='multinomial', h_n_samples=5))])
def test_epoch_loops_are_ordered_between_the_copy_and_the_compute_stream(sim, compute, data, extra):
    """bm_rbm_train_epoch[_u8|_bf16]: the copy stream uploads (and, for the bf16 engine, converts) batch i+1 into one half of a double
    buffer while the compute stream works on batch i from the other.  The stand-in runtime knows exactly which pages every copy and
    every interpreted kernel touches (page protection) and which operations are ordered by streams, events and host
    synchronisation: no two conflicting accesses may be unordered."""
    out = _stream_order_probe(EPOCH_PROBE % (compute, extra, data))
    assert 'RACES 0' in out and 'VIOLATION: \n' in out + '\n', out[-3000:]
    assert int(out.split('OPS')[1].split()[0]) > 50, out[-3000:]


def test_the_stream_order_check_sees_a_missing_event_wait(sim):
    """the same epoch with cudaStreamWaitEvent turned into a no-op: the copy stream now overwrites a staging buffer the compute
    stream may still read (and the compute stream reads a batch that may not have arrived) -- the checker must say so"""
    out = _stream_order_probe("\n    lib.fakecuda_ignore_event_waits(1)" + EPOCH_PROBE % ('bf16', {}, BYTES))
    assert 'stream race' in out and 'RACES 0' not in out, out[-3000:]
