"""ctypes binding of ``libbm.so`` (C-ABI in ``include/bm.h``) and the engine
objects the model classes talk to.

There is deliberately no fallback here: if the shared library is missing, was
not built for this GPU, or no GPU is visible, construction raises
``RuntimeError`` with the library's own message.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbm.so')

UNIT_KINDS = {'bernoulli': 0, 'multinomial': 1, 'gaussian': 2}
DTYPES = {'float32': 0, 'float64': 1}
COMPUTE = {'fp32': 0, 'bf16': 1}
METRIC_BITS = {'l2_loss': 1, 'msre': 2, 'pll': 4, 'free_energy': 8}
METRIC_SLOTS = {'l2_loss': 0, 'msre': 1, 'pll': 2, 'free_energy': 3}

# every symbol include/bm.h declares (tests check the library exports all of them)
EXPORTS = (
    'bm_version', 'bm_last_error', 'bm_device_count', 'bm_ctx_create', 'bm_ctx_destroy', 'bm_ctx_sync',
    'bm_ctx_timer_start', 'bm_ctx_timer_stop', 'bm_ctx_flush_l2', 'bm_host_alloc', 'bm_host_free',
    'bm_host_pack_u8', 'bm_host_pack_bf16',
    'bm_ctx_launch_count', 'bm_ctx_profile_tc', 'bm_ctx_profile_read', 'bm_comm_unique_id', 'bm_ctx_comm_init',
    'bm_rbm_create', 'bm_rbm_destroy', 'bm_rbm_set_param', 'bm_rbm_get_param', 'bm_rbm_init_weights',
    'bm_rbm_train_step', 'bm_rbm_set_data', 'bm_rbm_train_step_at', 'bm_rbm_train_epoch', 'bm_rbm_train_epoch_u8', 'bm_rbm_train_epoch_bf16', 'bm_rbm_transform', 'bm_rbm_metrics',
    'bm_rbm_get_activation', 'bm_debug_tc_gemm',
    'bm_dbm_create', 'bm_dbm_destroy', 'bm_dbm_set_param', 'bm_dbm_get_param', 'bm_dbm_init_particles',
    'bm_dbm_train_step', 'bm_dbm_val_metrics', 'bm_dbm_transform', 'bm_dbm_reconstruct', 'bm_dbm_log_proba',
    'bm_dbm_sample_v', 'bm_dbm_ais', 'bm_dbm_ais_rows',
)


class RbmCfg(C.Structure):
    _fields_ = [
        ('n_visible', C.c_int32), ('n_hidden', C.c_int32),
        ('v_kind', C.c_int32), ('h_kind', C.c_int32),
        ('dtype', C.c_int32), ('compute', C.c_int32),
        ('sample_v', C.c_int32), ('sample_h', C.c_int32),
        ('max_batch', C.c_int32), ('reserved0', C.c_int32),
        ('l2', C.c_double), ('dropout_keep', C.c_double),
        ('sparsity_target', C.c_double), ('sparsity_cost', C.c_double), ('sparsity_damping', C.c_double),
        ('propup_mult', C.c_double), ('propdown_mult', C.c_double),
        ('v_n_samples', C.c_double), ('h_n_samples', C.c_double),
        ('sigma', C.POINTER(C.c_double)),
    ]


class DbmCfg(C.Structure):
    _fields_ = [
        ('n_layers', C.c_int32), ('n_visible', C.c_int32),
        ('n_hiddens', C.POINTER(C.c_int32)),
        ('v_kind', C.c_int32),
        ('h_kinds', C.POINTER(C.c_int32)),
        ('h_n_samples', C.POINTER(C.c_double)),
        ('dtype', C.c_int32), ('compute', C.c_int32),
        ('n_particles', C.c_int32), ('batch_size', C.c_int32), ('max_mf_updates', C.c_int32),
        ('sample_v', C.c_int32),
        ('sample_h', C.POINTER(C.c_int32)),
        ('mf_tol', C.c_double), ('l2', C.c_double), ('max_norm', C.c_double), ('sparsity_damping', C.c_double),
        ('sparsity_target', C.POINTER(C.c_double)),
        ('sparsity_cost', C.POINTER(C.c_double)),
        ('sigma', C.POINTER(C.c_double)),
    ]


_lib = None


def load_library(path=None):
    """dlopen libbm.so and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise RuntimeError("native library not found at '{0}': build it with "
                           "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)".format(p))
    lib = C.CDLL(p)
    vp, i32, u32, u64, i64, dbl, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_int64, C.c_double, C.c_size_t
    lib.bm_version.restype = C.c_char_p
    lib.bm_last_error.restype = C.c_char_p
    protos = {
        'bm_device_count': [C.POINTER(C.c_int)],
        'bm_ctx_create': [C.c_int, C.POINTER(vp)],
        'bm_ctx_sync': [vp], 'bm_ctx_timer_start': [vp], 'bm_ctx_timer_stop': [vp, C.POINTER(C.c_float)],
        'bm_ctx_flush_l2': [vp], 'bm_host_alloc': [C.POINTER(vp), sz], 'bm_host_free': [vp],
        'bm_host_pack_u8': [vp, i32, sz, vp, C.POINTER(i32)], 'bm_host_pack_bf16': [vp, sz, vp],
        'bm_ctx_launch_count': [vp, C.POINTER(u64)],
        'bm_ctx_profile_tc': [vp, C.c_int],
        'bm_ctx_profile_read': [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(u64)],
        'bm_comm_unique_id': [vp], 'bm_ctx_comm_init': [vp, vp, C.c_int, C.c_int],
        'bm_rbm_create': [vp, C.POINTER(RbmCfg), C.POINTER(vp)],
        'bm_rbm_set_param': [vp, C.c_char_p, vp, sz], 'bm_rbm_get_param': [vp, C.c_char_p, vp, sz],
        'bm_rbm_init_weights': [vp, dbl, u64],
        'bm_rbm_train_step': [vp, vp, i32, dbl, dbl, i32, u64, u32, u32, C.POINTER(dbl)],
        'bm_rbm_set_data': [vp, vp, i64],
        'bm_rbm_train_step_at': [vp, i64, i32, dbl, dbl, i32, u64, u32, u32, C.POINTER(dbl)],
        'bm_rbm_train_epoch': [vp, vp, i64, i32, dbl, dbl, i32, u64, u32, u32, i32, i64, vp],
        'bm_rbm_train_epoch_u8': [vp, vp, i64, i32, dbl, dbl, i32, u64, u32, u32, i32, i64, vp],
        'bm_rbm_train_epoch_bf16': [vp, vp, i64, i32, dbl, dbl, i32, u64, u32, u32, i32, i64, vp],
        'bm_rbm_transform': [vp, vp, i32, i32, u64, u32, vp],
        'bm_rbm_metrics': [vp, vp, i32, i32, u64, u32, u32, C.POINTER(dbl)],
        'bm_rbm_get_activation': [vp, C.c_char_p, vp, sz],
        'bm_debug_tc_gemm': [vp, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp, i32, i32, i32, i32, vp],
        'bm_dbm_create': [vp, C.POINTER(DbmCfg), C.POINTER(vp)],
        'bm_dbm_set_param': [vp, C.c_char_p, vp, sz], 'bm_dbm_get_param': [vp, C.c_char_p, vp, sz],
        'bm_dbm_init_particles': [vp, u64],
        'bm_dbm_train_step': [vp, vp, i32, dbl, dbl, i32, u64, u32, i32, C.POINTER(dbl)],
        'bm_dbm_val_metrics': [vp, vp, i32, i32, u64, u32, C.POINTER(dbl)],
        'bm_dbm_transform': [vp, vp, i32, vp], 'bm_dbm_reconstruct': [vp, vp, i32, vp],
        'bm_dbm_log_proba': [vp, vp, i32, vp], 'bm_dbm_sample_v': [vp, i32, u64, u32, vp],
        'bm_dbm_ais': [vp, i32, i32, i32, u64, vp],
        'bm_dbm_ais_rows': [vp, i32, i32, i32, u64, u32, vp],
    }
    for name, argtypes in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.bm_ctx_destroy.argtypes = [vp]
    lib.bm_ctx_destroy.restype = None
    lib.bm_rbm_destroy.argtypes = [vp]
    lib.bm_rbm_destroy.restype = None
    lib.bm_dbm_destroy.argtypes = [vp]
    lib.bm_dbm_destroy.restype = None
    if path is None:
        _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError('libbm: ' + load_library().bm_last_error().decode('utf-8', 'replace'))


def device_count():
    n = C.c_int(0)
    check(load_library().bm_device_count(C.byref(n)))
    return n.value


class Context(object):
    """One GPU + compute stream (+ optional NCCL communicator)."""
    _default = {}

    def __init__(self, device=None):
        lib = load_library()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', os.environ.get('BM_DEVICE', '0')))
        self.device = device
        h = C.c_void_p()
        check(lib.bm_ctx_create(device, C.byref(h)))
        self.handle = h
        self.rank, self.nranks = 0, 1

    @classmethod
    def default(cls, device=None):
        key = device
        if key not in cls._default:
            cls._default[key] = cls(device)
        return cls._default[key]

    def sync(self):
        check(load_library().bm_ctx_sync(self.handle))

    def timer_start(self):
        check(load_library().bm_ctx_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float(0)
        check(load_library().bm_ctx_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(load_library().bm_ctx_flush_l2(self.handle))

    def launch_count(self):
        n = C.c_uint64(0)
        check(load_library().bm_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def profile_tc(self, enable):
        check(load_library().bm_ctx_profile_tc(self.handle, int(bool(enable))))

    def profile_read(self):
        """(algorithmic FLOPs, device ms, launches) of the tensor-core kernel since profiling was enabled."""
        f, ms, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        check(load_library().bm_ctx_profile_read(self.handle, C.byref(f), C.byref(ms), C.byref(n)))
        return f.value, ms.value, n.value

    def comm_init(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(load_library().bm_ctx_comm_init(self.handle, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        check(load_library().bm_comm_unique_id(buf))
        return buf.raw


# Page-locking is expensive (about 0.3 ms per MB): blocks released by `pinned_free` are kept in a small per-process pool and
# handed out again (what torch's caching host allocator does), so that the second `fit()` of a pipeline -- RBM #2 after RBM #1,
# the DBM after both -- or the next epoch's staging does not pay for it again.  `pinned_pool_clear()` returns them to the system.
_PINNED_POOL = {}            # bytes -> [addresses]
_PINNED_POOL_MAX = 2 << 30
_pinned_sizes = {}           # address -> bytes


def pinned_empty(shape, dtype=np.float32):
    """numpy array over page-locked host memory (for the per-batch feed path)."""
    n = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    free = _PINNED_POOL.get(n)
    if free:
        addr = free.pop()
    else:
        p = C.c_void_p()
        check(load_library().bm_host_alloc(C.byref(p), n))
        addr = p.value
    _pinned_sizes[addr] = n
    buf = (C.c_char * n).from_address(addr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


def pinned_free(arr):
    """Release an array made by `pinned_empty` / `pinned_copy` (the array must not be used afterwards)."""
    addr = np.asarray(arr).ctypes.data
    n = _pinned_sizes.pop(addr, None)
    if n is not None and sum(k * len(v) for k, v in _PINNED_POOL.items()) + n <= _PINNED_POOL_MAX:
        _PINNED_POOL.setdefault(n, []).append(addr)
        return
    check(load_library().bm_host_free(C.c_void_p(addr)))


def pinned_pool_clear():
    for n, addrs in list(_PINNED_POOL.items()):
        while addrs:
            check(load_library().bm_host_free(C.c_void_p(addrs.pop())))
    _PINNED_POOL.clear()


def pinned_copy(X):
    """Page-locked copy of a host array: lets `train_epoch` overlap the batch uploads with compute."""
    P = pinned_empty(X.shape, X.dtype)
    P[...] = X
    return P


def as_bytes(X, out=None):
    """X (float32 / float64) as uint8 if that is lossless (all values integers in 0..255), else None.  One multi-threaded
    native pass (bm_host_pack_u8); `out`: optional destination (e.g. page-locked memory) of X's shape."""
    if X.size == 0:
        return None
    if X.dtype not in (np.float32, np.float64) or not X.flags['C_CONTIGUOUS']:
        lo, hi = float(X.min()), float(X.max())
        if not (lo >= 0.0 and hi <= 255.0):          # also false for NaN
            return None
        Xb = X.astype(np.uint8)
        return Xb if np.array_equal(Xb, X) else None
    Xb = out if out is not None else np.empty(X.shape, dtype=np.uint8)
    exact = C.c_int32(0)
    check(load_library().bm_host_pack_u8(X.ctypes.data, DTYPES[X.dtype.name], X.size, Xb.ctypes.data, C.byref(exact)))
    return Xb if exact.value else None


class Bf16Array(np.ndarray):
    """uint16 array whose elements are bfloat16 bit patterns (a real-valued training set packed for
    bm_rbm_train_epoch_bf16).  `widen()` gives the float32 values back (exactly the rounded ones)."""
    def widen(self):
        return (np.asarray(self).astype(np.uint32) << 16).view(np.float32)


def as_bf16(X, out=None):
    """float32 X rounded to bfloat16 (round to nearest even, as the engine's own conversion), as a Bf16Array."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    P = out if out is not None else np.empty(X.shape, dtype=np.uint16)
    check(load_library().bm_host_pack_bf16(X.ctypes.data, X.size, P.ctypes.data))
    return P.view(Bf16Array)


def debug_tc_gemm(A, B, a_t=False, b_t=False, A2=None, B2=None, neg2=False, splits=1, ctx=None,
                  force_bn=0, force_cluster=0):
    """Raw tensor-core GEMM (test hook): C = A' B'^T (+/- A2' B2'^T) with the operands stored as
    given: A is [M,K] (or [K,M] if a_t), B is [N,K] (or [K,N] if b_t)."""
    ctx = ctx or Context.default()
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = (A.shape[1], A.shape[0]) if a_t else A.shape
    N = B.shape[1] if b_t else B.shape[0]
    K2 = 0
    p2a = p2b = None
    if A2 is not None:
        A2 = np.ascontiguousarray(A2, dtype=np.float32)
        B2 = np.ascontiguousarray(B2, dtype=np.float32)
        K2 = A2.shape[0] if a_t else A2.shape[1]
        p2a, p2b = A2.ctypes.data, B2.ctypes.data
    Cout = np.empty((M, N), dtype=np.float32)
    check(load_library().bm_debug_tc_gemm(ctx.handle, M, N, K, A.ctypes.data, int(a_t), B.ctypes.data, int(b_t),
                                          K2, p2a, p2b, int(neg2), int(splits), int(force_bn), int(force_cluster),
                                          Cout.ctypes.data))
    return Cout


def _mask(names):
    m = 0
    for n in names:
        m |= METRIC_BITS[n]
    return m


class CudaRBM(object):
    """Engine object for ``BaseRBM``: a ``bm_rbm`` handle."""

    def __init__(self, cfg, ctx=None):
        self._lib = load_library()
        self.ctx = ctx or Context.default()
        self.cfg = dict(cfg)
        self.V, self.H = int(cfg['n_visible']), int(cfg['n_hidden'])
        self.dt = np.dtype(cfg.get('dtype', 'float32'))
        c = RbmCfg()
        c.n_visible, c.n_hidden = self.V, self.H
        c.v_kind = UNIT_KINDS[cfg.get('v_kind', 'bernoulli')]
        c.h_kind = UNIT_KINDS[cfg.get('h_kind', 'bernoulli')]
        c.dtype = DTYPES[self.dt.name]
        # float32 models run on the tensor cores (bf16 operands, fp32 accumulate) unless told
        # otherwise; float64 models always use the storage-precision CUDA-core path
        compute = 'fp32'
        if self.dt == np.float32:
            compute = cfg.get('compute') or os.environ.get('BM_COMPUTE') or 'bf16'
        self.compute = compute
        c.compute = COMPUTE[compute]
        c.sample_v, c.sample_h = int(cfg.get('sample_v', False)), int(cfg.get('sample_h', True))
        c.max_batch = int(cfg.get('max_batch', 0))
        c.l2 = float(cfg.get('l2', 0.))
        keep = cfg.get('dropout', None)
        c.dropout_keep = -1.0 if keep is None else float(keep)
        c.sparsity_target = float(cfg.get('sparsity_target', 0.1))
        c.sparsity_cost = float(cfg.get('sparsity_cost', 0.))
        c.sparsity_damping = float(cfg.get('sparsity_damping', 0.9))
        c.propup_mult = 2.0 if cfg.get('dbm_first', False) else 1.0
        c.propdown_mult = 2.0 if cfg.get('dbm_last', False) else 1.0
        c.v_n_samples = float(cfg.get('v_n_samples', 100))
        c.h_n_samples = float(cfg.get('h_n_samples', 100))
        self._sigma = None
        if cfg.get('sigma', None) is not None:
            self._sigma = np.ascontiguousarray(
                np.broadcast_to(np.asarray(cfg['sigma'], dtype=np.float64), (self.V,)))
            c.sigma = self._sigma.ctypes.data_as(C.POINTER(C.c_double))
        h = C.c_void_p()
        check(self._lib.bm_rbm_create(self.ctx.handle, C.byref(c), C.byref(h)))
        self.handle = h
        self._names = ['W', 'vb', 'hb', 'dW', 'dvb', 'dhb', 'q_means'] + (['sigma'] if self._sigma is not None else [])
        self._shapes = {'W': (self.V, self.H), 'dW': (self.V, self.H), 'vb': (self.V,), 'dvb': (self.V,),
                        'hb': (self.H,), 'dhb': (self.H,), 'q_means': (self.H,), 'sigma': (self.V,)}
        self._metrics_buf = (C.c_double * 4)()

    # ---- variables ---------------------------------------------------------------
    def set_params(self, d):
        for k, v in d.items():
            if k == 'sigma':
                continue                      # fixed at construction (rbm/rbm.py:101-107)
            a = np.ascontiguousarray(v, dtype=self.dt).reshape(self._shapes[k])
            check(self._lib.bm_rbm_set_param(self.handle, k.encode(), a.ctypes.data, a.nbytes))

    def get_params(self, names=None):
        out = {}
        for k in (names or self._names):
            a = np.empty(self._shapes[k], dtype=self.dt)
            check(self._lib.bm_rbm_get_param(self.handle, k.encode(), a.ctypes.data, a.nbytes))
            out[k] = a
        return out

    def init_normal_W(self, stddev, op_seed):
        check(self._lib.bm_rbm_init_weights(self.handle, float(stddev), int(op_seed) & (2 ** 64 - 1)))

    # ---- compute -------------------------------------------------------------------
    def _batch(self, X):
        if isinstance(X, Bf16Array):
            X = X.widen()
        X = np.ascontiguousarray(X, dtype=self.dt)
        if X.ndim != 2 or X.shape[1] != self.V:
            raise ValueError('batch has shape {0}, expected (rows, {1})'.format(X.shape, self.V))
        return X

    def _collect(self, names):
        return {n: float(self._metrics_buf[METRIC_SLOTS[n]]) for n in names}

    def train_step(self, X, lr, momentum, k, seed, tick, metrics=()):
        X = self._batch(X)
        check(self._lib.bm_rbm_train_step(self.handle, X.ctypes.data, X.shape[0], lr, momentum, int(k),
                                          int(seed), int(tick), _mask(metrics), self._metrics_buf))
        return self._collect(metrics) if metrics else None

    def train_epoch(self, X, batch, lr, momentum, k, seed, tick0, metrics=(), every=0, iter0=0):
        """One pass over the host dataset X in mini-batches (uploads overlap compute).  Returns
        {metric: [value of every reporting batch]}; batch i uses tick0 + i, exactly like
        train_step(X[i*batch:(i+1)*batch], ..., tick=tick0 + i)."""
        byte_valued = getattr(X, 'dtype', None) == np.uint8 and not isinstance(X, Bf16Array)   # see as_bytes(): exact, 4x less traffic
        packed_bf16 = isinstance(X, Bf16Array)                  # see pin(): real-valued data for the bf16 engine, 2x less traffic
        if byte_valued or packed_bf16:
            X = np.ascontiguousarray(X)
            if X.ndim != 2 or X.shape[1] != self.V:
                raise ValueError('batch has shape {0}, expected (rows, {1})'.format(X.shape, self.V))
        else:
            X = self._batch(X)
        nb = (X.shape[0] + batch - 1) // batch
        out = np.zeros((nb, 4), dtype=np.float64)
        fn = (self._lib.bm_rbm_train_epoch_u8 if byte_valued else
              self._lib.bm_rbm_train_epoch_bf16 if packed_bf16 else self._lib.bm_rbm_train_epoch)
        check(fn(self.handle, X.ctypes.data, X.shape[0], int(batch), lr, momentum, int(k),
                 int(seed), int(tick0), _mask(metrics), int(every), int(iter0), out.ctypes.data))
        rep = [i for i in range(nb) if every and (iter0 + i + 1) % every == 0]
        return {m: [float(out[i, METRIC_SLOTS[m]]) for i in rep] for m in metrics}

    def pin(self, X):
        """Page-locked copy of the training set for `train_epoch`.  Data whose every value is an integer in
        0..255 (binarised MNIST: {0, 1}) is kept as one byte per unit: the engine widens it exactly on the
        device, so results do not change and each epoch moves a quarter of the bytes over PCIe."""
        X = self._batch(X)
        if X.size == 0:
            return pinned_copy(X)
        P = pinned_empty(X.shape, np.uint8)
        if as_bytes(X, out=P) is not None:
            return P
        pinned_free(P)
        # real-valued data: the bf16 engine rounds its input to bfloat16 before the first GEMM, so feeding the rounded
        # values is bit-identical and halves the copy (not with dropout / sigma scaling, which act on the fp32 input first)
        if (self.compute == 'bf16' and self.dt == np.float32 and self.cfg.get('v_kind', 'bernoulli') == 'bernoulli' and
                self.cfg.get('h_kind', 'bernoulli') == 'bernoulli' and self.cfg.get('dropout') is None):
            return as_bf16(X, out=pinned_empty(X.shape, np.uint16))
        return pinned_copy(X)

    def unpin(self, P):
        pinned_free(np.asarray(P))

    def set_data(self, X):
        X = self._batch(X)
        check(self._lib.bm_rbm_set_data(self.handle, X.ctypes.data, X.shape[0]))

    def train_step_at(self, first_row, rows, lr, momentum, k, seed, tick, metrics=()):
        check(self._lib.bm_rbm_train_step_at(self.handle, int(first_row), int(rows), lr, momentum, int(k),
                                             int(seed), int(tick), _mask(metrics), self._metrics_buf))
        return self._collect(metrics) if metrics else None

    def transform(self, X, k, seed, tick):
        X = self._batch(X)
        Hout = np.empty((X.shape[0], self.H), dtype=self.dt)
        check(self._lib.bm_rbm_transform(self.handle, X.ctypes.data, X.shape[0], int(k), int(seed), int(tick),
                                         Hout.ctypes.data))
        return Hout

    def metrics(self, X, k, seed, tick, names):
        X = self._batch(X)
        check(self._lib.bm_rbm_metrics(self.handle, X.ctypes.data, X.shape[0], int(k), int(seed), int(tick),
                                       _mask(names), self._metrics_buf))
        return self._collect(names)

    def get_activation(self, name, rows):
        n = self.V if name in ('X', 'v_means', 'v_states') else self.H
        dt = np.float32 if self.compute == 'bf16' else self.dt
        a = np.empty((rows, n), dtype=dt)
        check(self._lib.bm_rbm_get_activation(self.handle, name.encode(), a.ctypes.data, a.nbytes))
        return a

    def close(self):
        if getattr(self, 'handle', None):
            self._lib.bm_rbm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _sfx(i):
    return '' if i == 0 else '_{0}'.format(i)


class CudaDBM(object):
    """Engine object for ``DBM``: a ``bm_dbm`` handle."""

    def __init__(self, cfg, ctx=None):
        self._lib = load_library()
        self.ctx = ctx or Context.default()
        self.cfg = dict(cfg)
        self.V = int(cfg['n_visible'])
        self.Hs = [int(h) for h in cfg['n_hiddens']]
        self.L = len(self.Hs)
        self.M, self.B = int(cfg['n_particles']), int(cfg['batch_size'])
        self.dt = np.dtype(cfg.get('dtype', 'float32'))
        i32arr = lambda xs: (C.c_int32 * len(xs))(*[int(x) for x in xs])
        dblarr = lambda xs: (C.c_double * len(xs))(*[float(x) for x in xs])
        c = DbmCfg()
        c.n_layers, c.n_visible = self.L, self.V
        self._keep = [i32arr(self.Hs), i32arr([UNIT_KINDS[k] for k in cfg.get('h_kinds', ['bernoulli'] * self.L)]),
                      dblarr(cfg.get('h_n_samples', [100.] * self.L)), i32arr(cfg.get('sample_h', [True] * self.L)),
                      dblarr(cfg.get('sparsity_target', [0.1] * self.L)), dblarr(cfg.get('sparsity_cost', [0.] * self.L))]
        c.n_hiddens, c.h_kinds, c.h_n_samples, c.sample_h, c.sparsity_target, c.sparsity_cost = self._keep
        c.v_kind = UNIT_KINDS[cfg.get('v_kind', 'bernoulli')]
        c.dtype = DTYPES[self.dt.name]
        # float32 models with Bernoulli hidden layers (binary or Gaussian visibles) run on the tensor-core engine
        # (csrc/bm_dbm_tc.cuh: bf16 operands, fp32 accumulation) -- the same rule as DbmTC::supports(); compute='fp32'
        # (or BM_DBM_COMPUTE / BM_COMPUTE = fp32) selects the storage-precision CUDA-core engine, which every other
        # model uses anyway.
        compute = 'fp32'
        tc_ok = (self.dt == np.float32 and cfg.get('v_kind', 'bernoulli') in ('bernoulli', 'gaussian') and
                 all(k == 'bernoulli' for k in cfg.get('h_kinds', ['bernoulli'] * self.L)))
        if tc_ok:
            compute = cfg.get('compute') or os.environ.get('BM_DBM_COMPUTE') or os.environ.get('BM_COMPUTE') or 'bf16'
        self.compute = compute
        c.compute = COMPUTE[compute]
        c.n_particles, c.batch_size = self.M, self.B
        c.max_mf_updates = int(cfg.get('max_mf_updates', 10))
        c.sample_v = int(cfg.get('sample_v', True))
        c.mf_tol = float(cfg.get('mf_tol', 1e-7))
        c.l2 = float(cfg.get('l2', 0.))
        mn = float(cfg.get('max_norm', np.inf))
        c.max_norm = mn if np.isfinite(mn) else 1e300
        c.sparsity_damping = float(cfg.get('sparsity_damping', 0.9))
        self._sigma = None
        if cfg.get('sigma', None) is not None:
            self._sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(cfg['sigma'], dtype=np.float64), (self.V,)))
            c.sigma = self._sigma.ctypes.data_as(C.POINTER(C.c_double))
        h = C.c_void_p()
        check(self._lib.bm_dbm_create(self.ctx.handle, C.byref(c), C.byref(h)))
        self.handle = h
        self._shapes = {'vb': (self.V,), 'dvb': (self.V,), 'v': (self.M, self.V)}
        sizes = [self.V] + self.Hs
        for i in range(self.L):
            s = _sfx(i)
            self._shapes.update({'W' + s: (sizes[i], sizes[i + 1]), 'dW' + s: (sizes[i], sizes[i + 1]),
                                 'hb' + s: (self.Hs[i],), 'dhb' + s: (self.Hs[i],), 'mu' + s: (self.B, self.Hs[i]),
                                 'q_means' + s: (self.Hs[i],), 'mu_means' + s: (self.Hs[i],), 'h' + s: (self.M, self.Hs[i])})
        self._out2 = (C.c_double * 2)()

    def set_params(self, d):
        for k, v in d.items():
            if k == 'sigma':
                continue
            a = np.ascontiguousarray(v, dtype=self.dt).reshape(self._shapes[k])
            check(self._lib.bm_dbm_set_param(self.handle, k.encode(), a.ctypes.data, a.nbytes))

    def get_params(self, names=None):
        out = {}
        for k in (names or list(self._shapes)):
            a = np.empty(self._shapes[k], dtype=self.dt)
            check(self._lib.bm_dbm_get_param(self.handle, k.encode(), a.ctypes.data, a.nbytes))
            out[k] = a
        return out

    def init_particles(self, seed):
        check(self._lib.bm_dbm_init_particles(self.handle, int(seed) & (2 ** 64 - 1)))

    def _batch(self, X):
        X = np.ascontiguousarray(X, dtype=self.dt)
        if X.ndim != 2 or X.shape[1] != self.V:
            raise ValueError('batch has shape {0}, expected (rows, {1})'.format(X.shape, self.V))
        return X

    def train_step(self, X, lr, momentum, k, seed, tick, metrics=()):
        X = self._batch(X)
        check(self._lib.bm_dbm_train_step(self.handle, X.ctypes.data, X.shape[0], lr, momentum, int(k), int(seed), int(tick),
                                          1 if metrics else 0, self._out2))
        return {'msre': float(self._out2[0]), 'n_mf_updates': float(self._out2[1])} if metrics else None

    def val_metrics(self, X, k, seed, tick):
        X = self._batch(X)
        check(self._lib.bm_dbm_val_metrics(self.handle, X.ctypes.data, X.shape[0], int(k), int(seed), int(tick), self._out2))
        return {'msre': float(self._out2[0]), 'n_mf_updates': float(self._out2[1])}

    def transform(self, X):
        X = self._batch(X)
        out = np.empty((X.shape[0], self.Hs[-1]), dtype=self.dt)
        check(self._lib.bm_dbm_transform(self.handle, X.ctypes.data, X.shape[0], out.ctypes.data))
        return out

    def reconstruct(self, X):
        X = self._batch(X)
        out = np.empty((X.shape[0], self.V), dtype=self.dt)
        check(self._lib.bm_dbm_reconstruct(self.handle, X.ctypes.data, X.shape[0], out.ctypes.data))
        return out

    def log_proba(self, X):
        X = self._batch(X)
        out = np.empty(X.shape[0], dtype=np.float64)
        check(self._lib.bm_dbm_log_proba(self.handle, X.ctypes.data, X.shape[0], out.ctypes.data))
        return out

    def sample_v(self, k, seed, tick):
        out = np.empty((self.M, self.V), dtype=self.dt)
        check(self._lib.bm_dbm_sample_v(self.handle, int(k), int(seed), int(tick), out.ctypes.data))
        return out

    def ais(self, n_runs, n_betas, k, seed, first_run=None):
        """log Z estimates of ``n_runs`` AIS runs.  ``first_run=None``: the whole ladder (sharded over the ranks of
        the context's communicator, if any); an integer: runs [first_run, first_run + n_runs) of it, computed here."""
        out = np.empty(int(n_runs), dtype=np.float64)
        if first_run is None:
            check(self._lib.bm_dbm_ais(self.handle, int(n_runs), int(n_betas), int(k), int(seed), out.ctypes.data))
        else:
            check(self._lib.bm_dbm_ais_rows(self.handle, int(n_runs), int(n_betas), int(k), int(seed), int(first_run),
                                            out.ctypes.data))
        return out

    def close(self):
        if getattr(self, 'handle', None):
            self._lib.bm_dbm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_factory(kind):
    load_library()
    if kind == 'dbm':
        return CudaDBM
    if kind == 'rbm':
        return CudaRBM
    raise RuntimeError("no native engine for '{0}'".format(kind))
