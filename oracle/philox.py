"""Philox-4x32-10 and the TF-1.x bit->float conversions, in numpy.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.

What this file restates (the arithmetic lives in the un-vendored dependency
``tensorflow-gpu~=1.3.0``, /root/reference/requirements.txt:11):

* ``tf.random_normal(..., seed=random_seed)`` used for the weight init at
  /root/reference/boltzmann_machines/rbm/base_rbm.py:277-279.  TF-1.3 semantics
  (``tensorflow/core/lib/random/philox_random.h`` and
  ``random_distributions.h``): Philox-4x32-10, key = (seed_lo, seed_hi),
  counter = (0, 0, seed2_lo, seed2_hi), ``tf.get_seed(op_seed)`` with no graph
  seed -> (87654321, op_seed); float normals by Box-Muller on pairs of words.
  PINNED by the reference's own known-answer test
  /root/reference/boltzmann_machines/rbm/tests/test_rbm.py:65,67
  (W[0][0] == -0.0094548017 in float32, -0.0077341544416 in float64);
  see tests/test_oracle_philox.py.

* the engine's own counter layout for the Gibbs-chain draws (``site_block`` /
  ``uniform_at``), shared bit-for-bit with the CUDA kernels
  (boltzmann-machines_b200/csrc/bm_rng.cuh).  The reference's sampling stream is
  not reproducible without TF itself, so this layout is ours.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# draw sites (bits 0..7 of counter word 2); bits 8..31 carry the Gibbs index t
SITE_DROPOUT = 0
SITE_H0 = 1
SITE_V = 2
SITE_H = 3
SITE_PLL = 4
SITE_MULTINOMIAL_FE = 5
SITE_PARTICLE_INIT = 6
SITE_AIS_INIT = 7
SITE_AIS_V = 8
SITE_AIS_H2 = 9
SITE_AIS_H1 = 10
SITE_DBM_V = 11
SITE_DBM_H = 16          # + layer index


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32-10.  Counters are array-likes (broadcastable),
    keys python ints.  Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0 = hi1 ^ c1 ^ np.uint64(k0)
        c1 = lo1
        c2 = hi0 ^ c3 ^ np.uint64(k1)
        c3 = lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def u32_to_float(x):
    """TF ``Uint32ToFloat``: 23 mantissa bits, exponent 127, minus 1 -> [0,1)."""
    x = np.asarray(x, dtype=np.uint32)
    bits = (x & np.uint32(0x7FFFFF)) | np.uint32(127 << 23)
    return bits.view(np.float32) - np.float32(1.0)


def u64_to_double(x0, x1):
    """TF ``Uint64ToDouble``: 52 mantissa bits ((x0 & 0xfffff) << 32 | x1)."""
    x0 = np.asarray(x0, dtype=np.uint64)
    x1 = np.asarray(x1, dtype=np.uint64)
    bits = ((x0 & np.uint64(0xFFFFF)) << np.uint64(32)) | x1 | (np.uint64(1023) << np.uint64(52))
    return bits.view(np.float64) - 1.0


def box_muller_float(x0, x1):
    """TF ``BoxMullerFloat``: two uint32 -> two float32 normals (sin first)."""
    eps = np.float32(1.0e-7)
    u1 = np.maximum(u32_to_float(x0), eps)
    v1 = np.float32(2.0 * np.pi) * u32_to_float(x1)
    u2 = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    return (np.sin(v1).astype(np.float32) * u2, np.cos(v1).astype(np.float32) * u2)


def box_muller_double(x0, x1, x2, x3):
    eps = 1.0e-7
    u1 = np.maximum(u64_to_double(x0, x1), eps)
    v1 = 2.0 * np.pi * u64_to_double(x2, x3)
    u2 = np.sqrt(-2.0 * np.log(u1))
    return np.sin(v1) * u2, np.cos(v1) * u2


def tf_random_normal(shape, stddev, op_seed, dtype='float32', graph_seed=87654321):
    """``tf.random_normal(shape, 0, stddev, seed=op_seed)`` with no graph-level
    seed set (the situation of ``init()``; base_rbm.py:277-279, tf_model.py:168).

    Block i of the op uses counter (i, 0, seed2_lo, seed2_hi); float32 takes 4
    normals per block, float64 takes 2.  Only element 0 is pinned by the
    reference KAT; the rest follows TF's documented fill order.
    """
    n = int(np.prod(shape))
    k0, k1 = graph_seed & 0xFFFFFFFF, (graph_seed >> 32) & 0xFFFFFFFF
    s2lo, s2hi = op_seed & 0xFFFFFFFF, (op_seed >> 32) & 0xFFFFFFFF
    if dtype == 'float32':
        nb = (n + 3) // 4
        x0, x1, x2, x3 = philox4x32_10(np.arange(nb), 0, s2lo, s2hi, k0, k1)
        a, b = box_muller_float(x0, x1)
        c, d = box_muller_float(x2, x3)
        out = np.stack([a, b, c, d], axis=1).reshape(-1)[:n]
        return (out * np.float32(stddev)).reshape(shape).astype(np.float32)
    elif dtype == 'float64':
        nb = (n + 1) // 2
        x0, x1, x2, x3 = philox4x32_10(np.arange(nb), 0, s2lo, s2hi, k0, k1)
        a, b = box_muller_double(x0, x1, x2, x3)
        out = np.stack([a, b], axis=1).reshape(-1)[:n]
        return (out * float(stddev)).reshape(shape)
    raise ValueError(dtype)


# --------------------------------------------------------------------------
# Engine counter layout (shared with csrc/bm_rng.cuh)
# --------------------------------------------------------------------------
_clib = None


def _load_clib():
    """Optional C accelerator (oracle/philox_c.c -> oracle/_build/liboracle.so),
    built by __graft_entry__.build().  Pure numpy is the fallback."""
    global _clib
    if _clib is not None:
        return _clib
    import ctypes, os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'liboracle.so')
    if os.path.exists(p):
        try:
            lib = ctypes.CDLL(p)
            lib.oracle_site_words.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint64]
            lib.oracle_site_words.restype = None
            _clib = lib
            return lib
        except OSError:
            pass
    _clib = False
    return _clib


def site_words(rows, cols, seed, site, t, tick, row0=0, use_c=True):
    """uint32 words for every element (r, c) of a [rows, cols] draw.

    element (r, c): Philox counter (c // 4, row0 + r, site | t << 8, tick),
    key (seed & 0xffffffff, seed >> 32), word lane c % 4.
    """
    c2 = (int(site) & 0xFF) | ((int(t) & 0xFFFFFF) << 8)
    lib = _load_clib() if use_c else False
    if lib:
        out = np.empty((rows, cols), dtype=np.uint32)
        if out.size:
            lib.oracle_site_words(out.ctypes.data, rows, cols, row0, c2, int(tick) & 0xFFFFFFFF,
                                  int(seed) & 0xFFFFFFFFFFFFFFFF)
        return out
    nb = (cols + 3) // 4
    r = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
    b = np.arange(nb, dtype=np.uint64)[None, :]
    x = philox4x32_10(b, r, c2, int(tick) & 0xFFFFFFFF,
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    w = np.stack(x, axis=2).reshape(rows, nb * 4)
    return np.ascontiguousarray(w[:, :cols])


def uniform_at(rows, cols, seed, site, t, tick, row0=0):
    """float32 uniforms in [0,1) for a [rows, cols] draw (23-bit, TF style)."""
    return u32_to_float(site_words(rows, cols, seed, site, t, tick, row0))


def normal_at(rows, cols, seed, site, t, tick, row0=0):
    """float32 standard normals: lanes (0,1) Box-Muller of words (0,1), lanes
    (2,3) of words (2,3) of the element's block."""
    cp = (cols + 3) // 4 * 4
    w = site_words(rows, cp, seed, site, t, tick, row0).reshape(rows, cp // 4, 4)
    a, b = box_muller_float(w[:, :, 0], w[:, :, 1])
    c, d = box_muller_float(w[:, :, 2], w[:, :, 3])
    out = np.stack([a, b, c, d], axis=2).reshape(rows, cp)
    return np.ascontiguousarray(out[:, :cols])
