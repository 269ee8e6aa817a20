// Fused "layer op" on the 5th-generation tensor cores (sm_100a only).
//
//   C[M,N] = sum_p (+/-) A_p * B_p^T          bf16 operands, fp32 accumulation in TMEM
//   out    = sample(act(acc_scale * C * sigma + bias_scale * bias))
//
// replaces, per Gibbs half-step, the reference's tf.matmul + 3 element-wise kernels + the
// random_uniform/Less/Cast sampling chain (rbm/base_rbm.py:329-365, layers.py:34-51), and
// per training step the two dW GEMMs (base_rbm.py:447-448) as ONE GEMM over the concatenated
// batch dimension with the negative phase subtracted by the MMA's a_negate bit.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d, 128B-swizzled tiles, 4-stage mbarrier ring
//   warp 1   MMA issuer:   one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16)
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages of up to 256 fp32 columns)
//   warps 4-7 epilogue:    tcgen05.ld 32x32b (thread = one accumulator row), bias/activation,
//            Philox-4x32-10 in registers, bf16/fp32 stores; overlaps the next tile's MMAs.
// Operand layouts: both K-major and MN-major shared-memory descriptors are used so that a single
// bf16 copy of W serves v->h (W as MN-major B), h->v (W as K-major B) and no activation is ever
// transposed in memory (dW takes X and h as MN-major A and B).
#include "bm_tc.h"
#include <cuda.h>
#include <map>
#include <tuple>
#include <mutex>

namespace bm {

constexpr int BM = 128;            // rows per tile = TMEM lanes
constexpr int BK = 64;             // K per pipeline stage = one 128-byte swizzle atom of bf16
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;      // TMEM columns per accumulator stage
constexpr int A_BYTES = BM * BK * 2;           // 16 KiB
constexpr int B_BYTES = 256 * BK * 2;          // 32 KiB (BN <= 256)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 256;

struct TcParams {
    int M, N, BN;
    int m_tiles, n_tiles, splits;
    int n_pairs;
    int chunks[2];                 // K chunks (of BK) per pair
    int a_mn[2], b_mn[2], a_neg[2];
    int a_row0[2], a_k0[2];
    unsigned long long split_stride;
    float acc_scale, bias_scale;
    const float* bias;
    const float* sigma;
    const float* noise_sigma;
    int act, sample;
    RngKey rng;
    __nv_bfloat16* out_mean_bf;  int ld_mean_bf;
    __nv_bfloat16* out_state_bf; int ld_state_bf;
    float* out_f32;              int ld_f32;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address [0,14),
// leading byte offset [16,30), stride byte offset [32,46) (all >> 4), version=1 at bit 46,
// layout type SWIZZLE_128B = 2 at bits [61,64).
//   K-major tile  [rows][64 k] : 8-row groups are 1024 B apart (SBO); LBO unused.
//   MN-major tile [64 k][64 mn] boxes of 8 KiB: 64-wide MN blocks are 8192 B apart (LBO),
//                 8-k groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int mn_major) {
    const uint64_t lbo = mn_major ? (8192u >> 4) : 1u;
    const uint64_t sbo = 1024u >> 4;
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_softplus(float x) { return fmaxf(x, 0.f) + __logf(1.0f + __expf(-fabsf(x))); }

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB0);
        if (p.n_pairs > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_chunks = p.chunks[0] + (p.n_pairs > 1 ? p.chunks[1] : 0);
    const int units = p.m_tiles * p.n_tiles * p.splits;

    if (warp == 0) {
        // ================================ TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx_bytes = A_BYTES + (uint32_t)p.BN * BK * 2;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int split = unit % p.splits;
                const int tile = unit / p.splits;
                const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
                const int c_begin = (int)(((long long)total_chunks * split) / p.splits);
                const int c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
                for (int c = c_begin; c < c_end; ++c) {
                    const int pr = (c >= p.chunks[0]) ? 1 : 0;
                    const int kc = (pr ? c - p.chunks[0] : c) * BK;
                    const CUtensorMap* mA = pr ? &tmA1 : &tmA0;
                    const CUtensorMap* mB = pr ? &tmB1 : &tmB0;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * STAGE_BYTES;
                    uint8_t* sB = sA + A_BYTES;
                    mbar_expect_tx(&full[stage], tx_bytes);
                    if (!p.a_mn[pr]) {
                        tma_load_2d(sA, mA, &full[stage], kc, p.a_row0[pr] + m_blk * BM);
                    } else {
                        tma_load_2d(sA, mA, &full[stage], m_blk * BM, p.a_k0[pr] + kc);
                        tma_load_2d(sA + 8192, mA, &full[stage], m_blk * BM + 64, p.a_k0[pr] + kc);
                    }
                    if (!p.b_mn[pr]) {
                        tma_load_2d(sB, mB, &full[stage], kc, n_blk * p.BN);
                    } else {
                        for (int j = 0; j < p.BN / 64; ++j)
                            tma_load_2d(sB + j * 8192, mB, &full[stage], n_blk * p.BN + j * 64, kc);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ========================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=bf16 [7,10), B=bf16 [10,13),
            // a_negate 13, a_major 15, b_major 16, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int split = unit % p.splits;
                const int c_begin = (int)(((long long)total_chunks * split) / p.splits);
                const int c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                uint32_t accumulate = 0;
                for (int c = c_begin; c < c_end; ++c) {
                    const int pr = (c >= p.chunks[0]) ? 1 : 0;
                    const uint32_t idesc = idesc_base | ((uint32_t)p.a_neg[pr] << 13) | ((uint32_t)p.a_mn[pr] << 15) | ((uint32_t)p.b_mn[pr] << 16);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t aaddr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t baddr = aaddr + A_BYTES;
                    const uint32_t a_step = p.a_mn[pr] ? 2048u : 32u;    // bytes per K=16 slice
                    const uint32_t b_step = p.b_mn[pr] ? 2048u : 32u;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = make_smem_desc(aaddr + k * a_step, p.a_mn[pr]);
                        const uint64_t bdesc = make_smem_desc(baddr + k * b_step, p.b_mn[pr]);
                        umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                        accumulate = 1;
                    }
                    umma_commit(&empty[stage]);          // smem slot is free once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);                // accumulator complete -> epilogue
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ==========================================
        const int ew = warp - 4;                         // TMEM lane quarter this warp may read
        const int row = ew * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int split = unit % p.splits;
            const int tile = unit / p.splits;
            const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
            const int m = m_blk * BM + row;
            const bool row_ok = m < p.M;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            float* out_f32 = p.out_f32 ? p.out_f32 + (size_t)split * p.split_stride : nullptr;
            const int n_chunks = p.BN / 16;
            for (int ch = 0; ch < n_chunks; ++ch) {
                uint32_t v[16];
                __syncwarp();                            // tcgen05.ld is warp-collective (.sync.aligned)
                tmem_ld16(t_row + (uint32_t)(ch * 16), v);
                tmem_ld_wait();
                if (ch == n_chunks - 1) {
                    // all of this warp's reads of the accumulator are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
                const int n0 = n_blk * p.BN + ch * 16;
                if (n0 >= p.N) continue;
                float mean[16], state[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    U4 w{0, 0, 0, 0};
                    if (p.sample != SMP_NONE) w = site_block(p.rng, (uint32_t)m, (uint32_t)((n0 >> 2) + q));
                    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
                    float g[4] = {0.f, 0.f, 0.f, 0.f};
                    if (p.sample == SMP_GAUSSIAN) {
                        const float u1a = fmaxf(u32_to_unit_float(w.x), 1.0e-7f), u1b = fmaxf(u32_to_unit_float(w.z), 1.0e-7f);
                        const float ra = sqrtf(-2.0f * __logf(u1a)), rb = sqrtf(-2.0f * __logf(u1b));
                        float sa, ca, sb, cb;
                        __sincosf(6.2831853071795864769f * u32_to_unit_float(w.y), &sa, &ca);
                        __sincosf(6.2831853071795864769f * u32_to_unit_float(w.w), &sb, &cb);
                        g[0] = sa * ra; g[1] = ca * ra; g[2] = sb * rb; g[3] = cb * rb;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = q * 4 + j;
                        const int n = n0 + e;
                        const bool col_ok = n < p.N;
                        float x = p.acc_scale * __uint_as_float(v[e]);
                        if (p.sigma && col_ok) x *= p.sigma[n];
                        if (p.bias && col_ok) x = fmaf(p.bias_scale, p.bias[n], x);
                        float mu = x;
                        if (p.act == ACT_SIGMOID) mu = fast_sigmoid(x);
                        else if (p.act == ACT_SOFTPLUS) mu = fast_softplus(x);
                        float st = mu;
                        if (p.sample == SMP_BERNOULLI) st = (u32_to_unit_float(words[j]) < mu) ? 1.0f : 0.0f;
                        else if (p.sample == SMP_GAUSSIAN) st = mu + (p.noise_sigma && col_ok ? p.noise_sigma[n] : 1.0f) * g[j];
                        mean[e] = mu; state[e] = st;
                    }
                }
                if (!row_ok) continue;
                const bool full_chunk = (n0 + 16 <= p.N);
                if (p.out_mean_bf) {
                    __nv_bfloat16* dst = p.out_mean_bf + (size_t)m * p.ld_mean_bf + n0;
                    if (full_chunk) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(mean[2 * e], mean[2 * e + 1]);
                            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else {
                        for (int e = 0; e < 16 && n0 + e < p.N; ++e) dst[e] = __float2bfloat16_rn(mean[e]);
                    }
                }
                if (p.out_state_bf) {
                    __nv_bfloat16* dst = p.out_state_bf + (size_t)m * p.ld_state_bf + n0;
                    if (full_chunk) {
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(state[2 * e], state[2 * e + 1]);
                            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else {
                        for (int e = 0; e < 16 && n0 + e < p.N; ++e) dst[e] = __float2bfloat16_rn(state[e]);
                    }
                }
                if (out_f32) {
                    float* dst = out_f32 + (size_t)m * p.ld_f32 + n0;
                    if (full_chunk && (p.ld_f32 & 3) == 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            reinterpret_cast<float4*>(dst)[e] = make_float4(mean[4 * e], mean[4 * e + 1], mean[4 * e + 2], mean[4 * e + 3]);
                    } else {
                        for (int e = 0; e < 16 && n0 + e < p.N; ++e) dst[e] = mean[e];
                    }
                }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time -> no libcuda link dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) throw Error(BM_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    return fn;
}

// 2-D map over a row-major bf16 matrix: dim0 = columns (contiguous), dim1 = rows; box = box0 x box1;
// 128-byte swizzle; out-of-range elements read as zero (ragged M/N/K need no padding in memory).
static CUtensorMap make_map(const TcMat& m, int box0, int box1) {
    typedef std::tuple<const void*, int, int, int, int, int> Key;
    static std::map<Key, CUtensorMap> cache;
    static std::mutex mu;
    Key key(m.ptr, m.rows, m.cols, m.ld, box0, box1);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    BM_REQUIRE(m.ptr != nullptr && (reinterpret_cast<uintptr_t>(m.ptr) & 15) == 0, "tensor-core operand must be 16-byte aligned");
    BM_REQUIRE(m.ld % 8 == 0 && m.ld >= m.cols, "tensor-core operand leading dimension must be a multiple of 8");
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)m.ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(m.ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw Error(BM_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
    if (cache.size() > 4096) cache.clear();
    cache[key] = tm;
    return tm;
}

static int pick_bn(int N, bool need64, int m_tiles, int sms) {
    // candidates: multiples of 16 (64 for MN-major B) up to 256; minimise waves * BN (time), then padding
    const int step = need64 ? 64 : 16;
    int best = step; double best_cost = 1e30;
    for (int bn = step; bn <= 256; bn += step) {
        const int nt = (N + bn - 1) / bn;
        const long tiles = (long)nt * m_tiles;
        const long waves = (tiles + sms - 1) / sms;
        // smaller N tiles pay relatively more for the A operand's shared-memory reads
        const double eff = bn >= 128 ? 1.0 : (bn >= 64 ? 1.15 : 1.5);
        const double cost = (double)waves * bn * eff + 1e-3 * nt * bn;
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

void launch_tc_gemm(Ctx* ctx, const TcGemm& g) {
    BM_REQUIRE(g.M > 0 && g.N > 0 && g.n_pairs >= 1 && g.n_pairs <= 2, "bad tensor-core GEMM shape");
    static bool attr_set = false;
    if (!attr_set) {
        BM_CUDA(cudaFuncSetAttribute(tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    TcParams p{};
    p.M = g.M; p.N = g.N; p.n_pairs = g.n_pairs;
    p.m_tiles = (g.M + BM - 1) / BM;
    bool need64 = false;
    for (int i = 0; i < g.n_pairs; ++i) need64 = need64 || g.b_t[i];
    p.BN = pick_bn(g.N, need64, p.m_tiles * (g.splits > 0 ? g.splits : 1), ctx->sm_count);
    p.n_tiles = (g.N + p.BN - 1) / p.BN;
    p.splits = g.splits > 0 ? g.splits : 1;
    p.split_stride = g.split_stride;
    CUtensorMap maps[4];
    for (int i = 0; i < 2; ++i) {
        const int j = i < g.n_pairs ? i : 0;
        p.chunks[i] = i < g.n_pairs ? (g.K[j] + BK - 1) / BK : 0;
        p.a_mn[i] = g.a_t[j]; p.b_mn[i] = g.b_t[j]; p.a_neg[i] = g.neg[j];
        p.a_row0[i] = g.a_row0[j]; p.a_k0[i] = g.a_k0[j];
        // K-major: box = 64 k x (128 | BN) rows; MN-major: box = 64 mn x 64 k
        maps[2 * i] = g.a_t[j] ? make_map(g.A[j], 64, 64) : make_map(g.A[j], 64, BM);
        maps[2 * i + 1] = g.b_t[j] ? make_map(g.B[j], 64, 64) : make_map(g.B[j], 64, p.BN);
        BM_REQUIRE(i >= g.n_pairs || g.K[j] > 0, "tensor-core GEMM pair with K == 0");
    }
    const int total_chunks = p.chunks[0] + (g.n_pairs > 1 ? p.chunks[1] : 0);
    BM_REQUIRE(p.splits <= total_chunks, "more K splits than K chunks");
    BM_REQUIRE(p.splits == 1 || (g.out_f32 && !g.out_mean_bf && !g.out_state_bf), "split-K writes fp32 partials only");
    p.acc_scale = g.acc_scale; p.bias_scale = g.bias_scale;
    p.bias = g.bias; p.sigma = g.sigma; p.noise_sigma = g.noise_sigma;
    p.act = g.act; p.sample = g.sample; p.rng = g.rng;
    p.out_mean_bf = g.out_mean_bf; p.ld_mean_bf = g.ld_mean_bf;
    p.out_state_bf = g.out_state_bf; p.ld_state_bf = g.ld_state_bf;
    p.out_f32 = g.out_f32; p.ld_f32 = g.ld_f32;
    BM_REQUIRE(!g.out_mean_bf || (g.ld_mean_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
    BM_REQUIRE(!g.out_state_bf || (g.ld_state_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
    const int units = p.m_tiles * p.n_tiles * p.splits;
    const int grid = units < ctx->sm_count ? units : ctx->sm_count;
    if (ctx->profile_tc) BM_CUDA(cudaEventRecord(profile_event(ctx), ctx->stream));
    tc_layer_kernel<<<grid, TC_THREADS, SMEM_BYTES, ctx->stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    BM_CUDA(cudaGetLastError());
    if (ctx->profile_tc) {
        BM_CUDA(cudaEventRecord(profile_event(ctx), ctx->stream));
        double k_total = 0.0;
        for (int i = 0; i < g.n_pairs; ++i) k_total += g.K[i];
        ctx->prof_flops += 2.0 * g.M * g.N * k_total;       // algorithmic FLOPs (no padding counted)
        ctx->prof_launches++;
    }
    count_launch(ctx);
}

// ------------------------------------------------------------------------------------------
// small helpers on bf16 activations
// ------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd, int rows, int cols) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int r = blockIdx.y;
    if (c >= cols) return;
    const float a = src[(size_t)r * lds + c];
    const float b = (c + 1 < cols) ? src[(size_t)r * lds + c + 1] : 0.f;
    if (c + 1 < cols || c + 1 < ldd)
        *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)r * ldd + c) = __floats2bfloat162_rn(a, b);
    else
        dst[(size_t)r * ldd + c] = __float2bfloat16_rn(a);
}
void launch_f32_to_bf16(Ctx* ctx, const float* src, int lds, __nv_bfloat16* dst, int ldd, int rows, int cols) {
    if (rows <= 0) return;
    dim3 grid(((cols + 1) / 2 + 127) / 128, rows);
    f32_to_bf16_kernel<<<grid, 128, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols);
    count_launch(ctx);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < cols) dst[(size_t)r * ldd + c] = __bfloat162float(src[(size_t)r * lds + c]);
}
void launch_bf16_to_f32(Ctx* ctx, const __nv_bfloat16* src, int lds, float* dst, int ldd, int rows, int cols) {
    if (rows <= 0) return;
    dim3 grid((cols + 255) / 256, rows);
    bf16_to_f32_kernel<<<grid, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols);
    count_launch(ctx);
}

__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ P, int ldp, const __nv_bfloat16* __restrict__ Q, int ldq,
                                   int rows, int cols, float s1, float s2, float* __restrict__ out) {
    // 32 columns x 32 row-lanes per block; fp32 partial sums combined in a fixed order (deterministic)
    __shared__ float part[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float a = 0.f;
    if (c < cols) {
        for (int r = threadIdx.y; r < rows; r += 32) {
            float v = s1 * __bfloat162float(P[(size_t)r * ldp + c]);
            if (Q) v = fmaf(s2, __bfloat162float(Q[(size_t)r * ldq + c]), v);
            a += v;
        }
    }
    part[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += part[i][threadIdx.x];
        out[c] = s;
    }
}
void launch_colsum_bf16(Ctx* ctx, const __nv_bfloat16* P, int ldp, const __nv_bfloat16* Q, int ldq,
                        int rows, int cols, float s1, float s2, float* out) {
    if (cols <= 0) return;
    colsum_bf16_kernel<<<(cols + 31) / 32, dim3(32, 32), 0, ctx->stream>>>(P, ldp, Q, ldq, rows, cols, s1, s2, out);
    count_launch(ctx);
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, size_t stride, int splits, float* __restrict__ G, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        float4 a = *reinterpret_cast<const float4*>(partial + i);
        for (int s = 1; s < splits; ++s) {
            const float4 b = *reinterpret_cast<const float4*>(partial + s * stride + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4*>(G + i) = a;
    } else {
        for (size_t j = i; j < n; ++j) {
            float a = partial[j];
            for (int s = 1; s < splits; ++s) a += partial[s * stride + j];
            G[j] = a;
        }
    }
}
void launch_reduce_partials(Ctx* ctx, const float* partial, size_t stride, int splits, float* G, size_t n) {
    const size_t threads = (n + 3) / 4;
    reduce_partials_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(partial, stride, splits, G, n);
    count_launch(ctx);
}

}  // namespace bm
