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
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d, 128B-swizzled tiles, 4-stage mbarrier ring
//   warp 1   MMA issuer:   one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16)
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages of up to 256 fp32 columns)
//   warps 4-11 epilogue:   tcgen05.ld 32x32b.x32 (thread = one accumulator row, 32 columns), bias,
//            sigmoid, Philox-4x32-10 in registers, bf16/fp32 stores; two warps per TMEM lane
//            quarter alternate over the column chunks; overlaps the next tile's MMAs.
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
constexpr int MAX_STAGES = 6;       // 4 x 48 KiB (one CTA per tile) or 6 x 32 KiB (CTA pair: each CTA holds half of B)
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;      // TMEM columns per accumulator stage
constexpr int A_BYTES = BM * BK * 2;           // 16 KiB
constexpr int B_BYTES = 256 * BK * 2;          // 32 KiB (BN <= 256)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int RING_BYTES = 4 * STAGE_BYTES;   // == 6 * (A_BYTES + B_BYTES / 2)
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (4 + EPI_WARPS);

struct TcParams {
    int M, N, BN;
    int m_tiles, n_tiles, splits;
    int cluster;                   // 1: one CTA per 128-row tile; 2: CTA pair, tcgen05 cta_group::2 (M = 256, each CTA holds half of B)
    int stages, stage_bytes;       // shared-memory ring geometry
    int m_groups;                  // ceil(m_tiles / cluster)
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
    unsigned long long* dbg;     // optional timeline (globaltimer ns) of CTA 0: see bm_debug_tc_timeline
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
// multicast variant: the box lands at the same shared-memory offset in every CTA of `mask`, and
// completes `bytes` on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// cta_group::2 loads: data lands in the issuing CTA, the transaction bytes are counted on the
// LEADER CTA's mbarrier (address with the peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {      // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DBG_MARK(slot) do { if (p.dbg && blockIdx.x == 0) p.dbg[(slot)] = (unsigned long long)clock64(); } while (0)
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
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
// Epilogue specialisations: MODE 0 reads act/sample/outputs from TcParams at run time (every
// combination); MODE 1..4 fix them at compile time for the hot CD-k shapes so that the per-element
// instruction count stays near the Philox + sigmoid minimum.
enum : int {
    MODE_GENERIC = 0,
    MODE_SIG_BERN_MEAN_STATE = 1,    // h0: sigmoid, Bernoulli draw, bf16 means + bf16 states
    MODE_SIG_BERN_STATE = 2,         // mid-chain hidden: states only
    MODE_SIG_MEAN = 3,               // probabilities only (visible means, last hidden means)
    MODE_RAW_F32 = 4                 // raw fp32 accumulators (dW partials, linear pre-activations)
};

template <int MODE> struct EpiCfg {
    static constexpr bool fixed = MODE != MODE_GENERIC;
    static constexpr int act = (MODE == MODE_RAW_F32) ? ACT_LINEAR : ACT_SIGMOID;
    static constexpr int sample = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_BERN_STATE) ? SMP_BERNOULLI : SMP_NONE;
    static constexpr bool mean_bf = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_MEAN);
    static constexpr bool state_bf = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_BERN_STATE);
    static constexpr bool f32 = (MODE == MODE_RAW_F32);
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// sigmoid(x) = 1 / (1 + 2^(-x log2 e)): ex2.approx + rcp.approx keep the *relative* error of small
// probabilities at ~1e-7 (tanh.approx would not)
__device__ __forceinline__ float sigmoid_from_neg_log2(float t) {   // t = -x * log2(e)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h2);
}

template <int MODE, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const TcParams p) {
    typedef EpiCfg<MODE> E;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { DBG_MARK(0); if (p.dbg && blockIdx.x == 0) p.dbg[6] = gtime(); }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB0);
        if (p.n_pairs > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    }
    if (warp == 1 && lane == 0) {
        // pair: the leader's `full` collects its own expect_tx-arrive and the peer's arrive; its `tempty`
        // collects the epilogue warps of both CTAs; `empty`/`tfull` get one multicast commit each
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], (uint32_t)CL); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], (uint32_t)(EPI_WARPS * CL)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CL == 2) {     // one warp of each CTA of the pair allocates collectively
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();   // peers' barriers must exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) DBG_MARK(1);

    const int total_chunks = p.chunks[0] + (p.n_pairs > 1 ? p.chunks[1] : 0);
    // work units are (row-block group, column block, K split); the CTAs of a cluster walk the same
    // sequence of units and take consecutive row blocks of the group
    const int units = p.m_groups * p.n_tiles * p.splits;
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int unit0 = (CL > 1) ? (int)cluster_id_x() : (int)blockIdx.x;
    const int unit_step = (CL > 1) ? (int)cluster_count_x() : (int)gridDim.x;
    constexpr bool pair = (CL == 2);
    const int half_bn = p.BN >> 1;

    if (warp == 0) {
        // ================================ TMA producer =====================================
        // The whole warp walks the K chunks; for each chunk lane 0 arms the barrier and lanes
        // 0..n_ops-1 issue one bulk-tensor copy each in the same warp instruction (a single thread
        // issuing the 2-6 boxes of a stage back to back was measured at ~150-300 cycles per box and
        // starved the tensor pipe).
        int stage = 0; uint32_t phase = 0;
        const uint32_t tx_bytes = pair ? 2u * (A_BYTES + (uint32_t)half_bn * BK * 2) : A_BYTES + (uint32_t)p.BN * BK * 2;
        const int b_cols = pair ? half_bn : p.BN;          // B columns (rows of a K-major B tile) this CTA fetches
        for (int unit = unit0; unit < units; unit += unit_step) {
            const int split = unit % p.splits;
            const int tile = unit / p.splits;
            const int m_blk = (tile / p.n_tiles) * CL + crank, n_blk = tile % p.n_tiles;
            const int c_begin = (int)(((long long)total_chunks * split) / p.splits);
            const int c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
            const int n_col0 = n_blk * p.BN + (pair ? crank * half_bn : 0);
            for (int c = c_begin; c < c_end; ++c) {
                const int pr = (c >= p.chunks[0]) ? 1 : 0;
                const int kc = (pr ? c - p.chunks[0] : c) * BK;
                const CUtensorMap* mA = pr ? &tmA1 : &tmA0;
                const CUtensorMap* mB = pr ? &tmB1 : &tmB0;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sA = smem + stage * p.stage_bytes;
                uint8_t* sB = sA + A_BYTES;
                if (lane == 0 && c - c_begin < 24) DBG_MARK(8 + (c - c_begin));
                const int nA = p.a_mn[pr] ? 2 : 1;
                const int nB = p.b_mn[pr] ? b_cols / 64 : 1;
                // this lane's copy: destination, map, coordinates
                void* dst = nullptr; const CUtensorMap* map = nullptr; int c0 = 0, c1 = 0;
                if (lane < nA) {
                    map = mA; dst = sA + lane * 8192;
                    if (!p.a_mn[pr]) { c0 = kc; c1 = p.a_row0[pr] + m_blk * BM; }
                    else { c0 = m_blk * BM + lane * 64; c1 = p.a_k0[pr] + kc; }
                } else if (lane < nA + nB) {
                    const int j = lane - nA;
                    map = mB; dst = sB + j * 8192;
                    if (!p.b_mn[pr]) { c0 = kc; c1 = n_col0; }
                    else { c0 = n_col0 + j * 64; c1 = kc; }
                }
                if constexpr (pair) {
                    // both CTAs load their own A rows and their half of the B tile; bytes are counted on the leader
                    const uint32_t lbar = smem_u32(&full[stage]) & 0xFEFFFFFFu;
                    if (map) tma_load_2d_2sm(dst, map, lbar, c0, c1);
                    if (lane == 0) { if (crank == 0) mbar_expect_tx(&full[stage], tx_bytes); else mbar_arrive_remote(&full[stage], 0); }
                } else {
                    if (lane == 0) mbar_expect_tx(&full[stage], tx_bytes);
                    __syncwarp();
                    if (map) tma_load_2d(dst, map, &full[stage], c0, c1);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (the pair's leader CTA only) ============
        if (lane == 0 && crank == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=bf16 [7,10), B=bf16 [10,13),
            // a_negate 13, a_major 15, b_major 16, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
            for (int unit = unit0; unit < units; unit += unit_step) {
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
                    if (c - c_begin < 24) DBG_MARK(32 + (c - c_begin));
                    tc_fence_after();
                    const uint32_t aaddr = smem_u32(smem + stage * p.stage_bytes);
                    const uint32_t baddr = aaddr + A_BYTES;
                    const uint32_t a_step = p.a_mn[pr] ? 2048u : 32u;    // bytes per K=16 slice
                    const uint32_t b_step = p.b_mn[pr] ? 2048u : 32u;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = make_smem_desc(aaddr + k * a_step, p.a_mn[pr]);
                        const uint64_t bdesc = make_smem_desc(baddr + k * b_step, p.b_mn[pr]);
                        if constexpr (pair) umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, accumulate); else umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                        accumulate = 1;
                    }
                    // the slot is free once these MMAs retire; with multicast every producer of the
                    // cluster writes into this CTA's slot, so every CTA's `empty` barrier is told
                    if constexpr (pair) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if constexpr (pair) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);   // accumulator complete -> epilogue
                DBG_MARK(2);
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue (8 warps) =================================
        // warp w may read TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter alternate over
        // the tile's 32-column chunks.
        const int ew = warp - 4;
        const int quarter = ew & 3, half = ew >> 2;
        const int row = quarter * 32 + lane;
        const int act = E::fixed ? E::act : p.act;
        const int smp = E::fixed ? E::sample : p.sample;
        __nv_bfloat16* const out_mean = (E::fixed && !E::mean_bf) ? nullptr : p.out_mean_bf;
        __nv_bfloat16* const out_state = (E::fixed && !E::state_bf) ? nullptr : p.out_state_bf;
        float* const out_f32_base = (E::fixed && !E::f32) ? nullptr : p.out_f32;
        const float kNegLog2e = -1.4426950408889634f;
        // fold the sigmoid's -log2(e) into the affine map of the accumulator
        const float a_s = (act == ACT_SIGMOID) ? p.acc_scale * kNegLog2e : p.acc_scale;
        const float b_s = (act == ACT_SIGMOID) ? p.bias_scale * kNegLog2e : p.bias_scale;
        const bool has_sigma = !E::fixed && p.sigma != nullptr;
        const int n_chunks32 = (p.BN + 31) / 32;
        int acc = 0; uint32_t acc_phase = 0;
        for (int unit = unit0; unit < units; unit += unit_step) {
            const int split = unit % p.splits;
            const int tile = unit / p.splits;
            const int m_blk = (tile / p.n_tiles) * CL + crank, n_blk = tile % p.n_tiles;
            const int m = m_blk * BM + row;
            const bool row_ok = m < p.M;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (threadIdx.x == 128) DBG_MARK(3);
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            float* out_f32 = out_f32_base ? out_f32_base + (size_t)split * p.split_stride : nullptr;
            int last_ch = -1;
            for (int ch = half; ch < n_chunks32; ch += 2) last_ch = ch;
            if (last_ch < 0) {          // this warp has no chunk in the tile: release the accumulator at once
                __syncwarp();
                if (lane == 0) { if (pair && crank == 1) { if constexpr (pair) mbar_arrive_remote(&tempty[acc], 0); } else mbar_arrive(&tempty[acc]); }
            }
            for (int ch = half; ch < n_chunks32; ch += 2) {
                uint32_t v[32];
                __syncwarp();                            // tcgen05.ld is warp-collective (.sync.aligned)
                if (ch * 32 + 32 <= p.BN) {
                    tmem_ld32(t_row + (uint32_t)(ch * 32), v);
                } else {                                 // BN is a multiple of 16: a trailing half chunk
                    uint32_t lo[16];
                    tmem_ld16(t_row + (uint32_t)(ch * 32), lo);
#pragma unroll
                    for (int e = 0; e < 16; ++e) { v[e] = lo[e]; v[16 + e] = 0u; }
                }
                tmem_ld_wait();
                if (ch == last_ch) {
                    // all of this warp's reads of the accumulator are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (pair && crank == 1) { if constexpr (pair) mbar_arrive_remote(&tempty[acc], 0); } else mbar_arrive(&tempty[acc]); }
                }
                const int n0 = n_blk * p.BN + ch * 32;
                if (n0 >= p.N || !row_ok) continue;
                const int n_valid = min(32, min(p.N, n_blk * p.BN + p.BN) - n0);
                const bool full_chunk = (n_valid == 32);
                uint32_t mean_pk[16], state_pk[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float bq[4] = {0.f, 0.f, 0.f, 0.f};
                    if (p.bias) {
                        if (full_chunk) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + q);
                            bq[0] = b4.x; bq[1] = b4.y; bq[2] = b4.z; bq[3] = b4.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (q * 4 + j < n_valid) bq[j] = p.bias[n0 + q * 4 + j];
                        }
                    }
                    U4 w{0, 0, 0, 0};
                    if (smp != SMP_NONE) w = site_block(p.rng, (uint32_t)m, (uint32_t)((n0 >> 2) + q));
                    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
                    float g[4] = {0.f, 0.f, 0.f, 0.f};
                    if (!E::fixed && smp == SMP_GAUSSIAN) {
                        const float u1a = fmaxf(u32_to_unit_float(w.x), 1.0e-7f), u1b = fmaxf(u32_to_unit_float(w.z), 1.0e-7f);
                        const float ra = sqrtf(-2.0f * __logf(u1a)), rb = sqrtf(-2.0f * __logf(u1b));
                        float sa, ca, sb, cb;
                        __sincosf(6.2831853071795864769f * u32_to_unit_float(w.y), &sa, &ca);
                        __sincosf(6.2831853071795864769f * u32_to_unit_float(w.w), &sb, &cb);
                        g[0] = sa * ra; g[1] = ca * ra; g[2] = sb * rb; g[3] = cb * rb;
                    }
                    float mu[4], st[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = q * 4 + j;
                        float x = a_s * __uint_as_float(v[e]);
                        if (has_sigma && e < n_valid) x *= p.sigma[n0 + e];
                        x = fmaf(b_s, bq[j], x);
                        float m_ = x;
                        if (act == ACT_SIGMOID) m_ = sigmoid_from_neg_log2(x);
                        else if (!E::fixed && act == ACT_SOFTPLUS) m_ = fast_softplus(x);
                        float s_ = m_;
                        if (smp == SMP_BERNOULLI) s_ = (u32_to_unit_float(words[j]) < m_) ? 1.0f : 0.0f;
                        else if (!E::fixed && smp == SMP_GAUSSIAN)
                            s_ = m_ + ((p.noise_sigma && e < n_valid) ? p.noise_sigma[n0 + e] : 1.0f) * g[j];
                        mu[j] = m_; st[j] = s_;
                    }
                    if (out_mean) { mean_pk[2 * q] = pack_bf16(mu[0], mu[1]); mean_pk[2 * q + 1] = pack_bf16(mu[2], mu[3]); }
                    if (out_state) { state_pk[2 * q] = pack_bf16(st[0], st[1]); state_pk[2 * q + 1] = pack_bf16(st[2], st[3]); }
                    if (out_f32) {
                        float* dst = out_f32 + (size_t)m * p.ld_f32 + n0 + q * 4;
                        if (full_chunk && (p.ld_f32 & 3) == 0) {
                            *reinterpret_cast<float4*>(dst) = make_float4(mu[0], mu[1], mu[2], mu[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (q * 4 + j < n_valid) dst[j] = mu[j];
                        }
                    }
                }
                if (out_mean) {
                    __nv_bfloat16* dst = out_mean + (size_t)m * p.ld_mean_bf + n0;
                    if (full_chunk) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<uint4*>(dst)[i] = make_uint4(mean_pk[4 * i], mean_pk[4 * i + 1], mean_pk[4 * i + 2], mean_pk[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (e < n_valid) reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)(mean_pk[e >> 1] >> ((e & 1) * 16));
                    }
                }
                if (out_state) {
                    __nv_bfloat16* dst = out_state + (size_t)m * p.ld_state_bf + n0;
                    if (full_chunk) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<uint4*>(dst)[i] = make_uint4(state_pk[4 * i], state_pk[4 * i + 1], state_pk[4 * i + 2], state_pk[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (e < n_valid) reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)(state_pk[e >> 1] >> ((e & 1) * 16));
                    }
                }
            }
            if (threadIdx.x == 128) DBG_MARK(4);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    // no CTA may exit while a peer can still multicast into its shared memory or signal its barriers
    if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    if (threadIdx.x == 0) { DBG_MARK(5); if (p.dbg && blockIdx.x == 0) p.dbg[7] = gtime(); }
}

typedef void (*TcKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);

template <int CL> static TcKernelFn tc_kernel_mode(int mode) {
    switch (mode) {
        case MODE_SIG_BERN_MEAN_STATE: return tc_layer_kernel<MODE_SIG_BERN_MEAN_STATE, CL>;
        case MODE_SIG_BERN_STATE: return tc_layer_kernel<MODE_SIG_BERN_STATE, CL>;
        case MODE_SIG_MEAN: return tc_layer_kernel<MODE_SIG_MEAN, CL>;
        case MODE_RAW_F32: return tc_layer_kernel<MODE_RAW_F32, CL>;
        default: return tc_layer_kernel<MODE_GENERIC, CL>;
    }
}
// kernels that contain cta_group::2 instructions must be launched as clusters of 2, so the
// single-CTA and CTA-pair variants are separate instantiations
static TcKernelFn tc_kernel_for(int mode, int cluster) { return cluster == 2 ? tc_kernel_mode<2>(mode) : tc_kernel_mode<1>(mode); }

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

struct TilePick { int bn, cluster; };

static TilePick pick_tile(int N, bool b_mn, int m_tiles, int splits, int chunks, int sms) {
    // Tile width BN and CTA grouping.  A single SM ingests ~55-64 B/clk from L2 (measured: the
    // 4-stage ring of a 128x256 tile refills at 53 B/clk), while tcgen05 at M=128 consumes
    // 8192*(1/128 + 1/BN) B/clk of operands: a lone CTA is ingest-bound (~55% MMA duty).  A CTA pair
    // (cta_group::2, M=256) halves the B bytes each SM needs -> 64 B/clk at BN=256.
    // Cycle model per CTA: waves * chunks * max(mma, ingest) + epilogue of the last tile.
    TilePick best{b_mn ? 64 : 16, 1};
    double best_cost = 1e30;
    for (int c = 1; c <= 2; ++c) {
        if (c == 2 && m_tiles < 2) break;
        const int step = b_mn ? 64 * c : 16;              // every CTA of a pair holds BN/2 columns of B (whole boxes / 8-row atoms)
        const int slots = sms / c;
        for (int bn = step; bn <= 256; bn += step) {
            const int nt = (N + bn - 1) / bn;
            const int m_groups = (m_tiles + c - 1) / c;
            const long units = (long)m_groups * nt * splits;
            const long waves = (units + slots - 1) / slots;
            const double mma = 4.0 * (bn / 2.0);                                  // cycles per K chunk (4 x K=16)
            const double ingest = (16384.0 + (bn / c) * 128.0) / 55.0;           // bytes per chunk per SM / (B/clk)
            const double cost = (double)waves * chunks * (mma > ingest ? mma : ingest) + 12.0 * bn;
            if (cost < best_cost - 1e-9) { best_cost = cost; best.bn = bn; best.cluster = c; }
        }
    }
    return best;
}

void launch_tc_gemm(Ctx* ctx, const TcGemm& g) {
    BM_REQUIRE(g.M > 0 && g.N > 0 && g.n_pairs >= 1 && g.n_pairs <= 2, "bad tensor-core GEMM shape");
    static bool attr_set = false;
    if (!attr_set) {
        for (int md = 0; md <= 4; ++md)
            for (int cl = 1; cl <= 2; ++cl)
                BM_CUDA(cudaFuncSetAttribute(tc_kernel_for(md, cl), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    TcParams p{};
    p.M = g.M; p.N = g.N; p.n_pairs = g.n_pairs;
    p.m_tiles = (g.M + BM - 1) / BM;
    bool need64 = false;
    for (int i = 0; i < g.n_pairs; ++i) need64 = need64 || g.b_t[i];
    int chunks_total = 0;
    for (int i = 0; i < g.n_pairs; ++i) chunks_total += (g.K[i] + BK - 1) / BK;
    const int nsplit = g.splits > 0 ? g.splits : 1;
    TilePick tp = pick_tile(g.N, need64, p.m_tiles, nsplit, (chunks_total + nsplit - 1) / nsplit, ctx->sm_count);
    if (g.force_bn > 0) { tp.bn = g.force_bn; tp.cluster = g.force_cluster > 0 ? g.force_cluster : 1; }
    BM_REQUIRE(tp.cluster == 1 || tp.cluster == 2, "cluster must be 1 or 2");
    BM_REQUIRE(tp.cluster == 1 || (need64 ? tp.bn % 128 == 0 : tp.bn % 16 == 0), "CTA-pair tiles need BN/2 on whole boxes / swizzle atoms");
    p.BN = tp.bn; p.cluster = tp.cluster;
    p.m_groups = (p.m_tiles + p.cluster - 1) / p.cluster;
    p.n_tiles = (g.N + p.BN - 1) / p.BN;
    p.stages = p.cluster == 2 ? 6 : 4;
    p.stage_bytes = p.cluster == 2 ? (A_BYTES + B_BYTES / 2) : STAGE_BYTES;
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
        maps[2 * i + 1] = g.b_t[j] ? make_map(g.B[j], 64, 64) : make_map(g.B[j], 64, p.BN / p.cluster);
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
    p.dbg = g.dbg;
    BM_REQUIRE(!g.out_mean_bf || (g.ld_mean_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
    BM_REQUIRE(!g.out_state_bf || (g.ld_state_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
    const int units = p.m_groups * p.n_tiles * p.splits;
    const int max_clusters = ctx->sm_count / p.cluster;
    const int n_clusters = units < max_clusters ? units : max_clusters;
    const int grid = n_clusters * p.cluster;
    if (ctx->profile_tc) BM_CUDA(cudaEventRecord(profile_event(ctx), ctx->stream));
    int mode = MODE_GENERIC;
    if (!g.sigma && !g.noise_sigma) {
        const bool mb = g.out_mean_bf != nullptr, sb = g.out_state_bf != nullptr, f = g.out_f32 != nullptr;
        if (g.act == ACT_SIGMOID && g.sample == SMP_BERNOULLI && mb && sb && !f) mode = MODE_SIG_BERN_MEAN_STATE;
        else if (g.act == ACT_SIGMOID && g.sample == SMP_BERNOULLI && !mb && sb && !f) mode = MODE_SIG_BERN_STATE;
        else if (g.act == ACT_SIGMOID && g.sample == SMP_NONE && mb && !sb && !f) mode = MODE_SIG_MEAN;
        else if (g.act == ACT_LINEAR && g.sample == SMP_NONE && !mb && !sb && f) mode = MODE_RAW_F32;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = SMEM_BYTES; lc.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    BM_CUDA(cudaLaunchKernelEx(&lc, tc_kernel_for(mode, p.cluster), maps[0], maps[1], maps[2], maps[3], p));
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

// ---- column statistics of bf16 activations: up to 3 jobs (dvb, dhb, q) in one pair of launches ----
struct ColsumJobs {
    const __nv_bfloat16* P[3]; int ldp[3];
    const __nv_bfloat16* Q[3]; int ldq[3];
    float s1[3], s2[3];
    float* out[3];
    int cols[3];
    int rows, n;
};
constexpr int CS_RSPLIT = 32;

__global__ void colsum_bf16_partial_kernel(ColsumJobs j, float* __restrict__ partial, int max_cols) {
    // block: 32 x 8 threads; 64 columns (2 per thread) x one row slab; fixed combine order
    __shared__ float2 part[8][33];
    const int job = blockIdx.z;
    const int cols = j.cols[job];
    const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
    const int slab = (j.rows + CS_RSPLIT - 1) / CS_RSPLIT;
    const int r0 = blockIdx.y * slab, r1 = min(j.rows, r0 + slab);
    float2 a = make_float2(0.f, 0.f);
    if (c < cols) {
        const __nv_bfloat16* P = j.P[job]; const __nv_bfloat16* Q = j.Q[job];
        const float s1 = j.s1[job], s2 = j.s2[job];
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(P + (size_t)r * j.ldp[job] + c));
            a.x = fmaf(s1, p.x, a.x); a.y = fmaf(s1, p.y, a.y);
            if (Q) {
                const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Q + (size_t)r * j.ldq[job] + c));
                a.x = fmaf(s2, q.x, a.x); a.y = fmaf(s2, q.y, a.y);
            }
        }
    }
    part[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s.x += part[i][threadIdx.x].x; s.y += part[i][threadIdx.x].y; }
        float* dst = partial + ((size_t)job * CS_RSPLIT + blockIdx.y) * max_cols + c;
        dst[0] = s.x;
        if (c + 1 < cols) dst[1] = s.y;
    }
}
__global__ void colsum_bf16_finish_kernel(ColsumJobs j, const float* __restrict__ partial, int max_cols) {
    const int job = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= j.cols[job]) return;
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < CS_RSPLIT; ++i) s += partial[((size_t)job * CS_RSPLIT + i) * max_cols + c];
    j.out[job][c] = s;
}
static float* colsum_scratch(Ctx* ctx, size_t floats) {
    static float* buf[64] = {nullptr};
    static size_t cap[64] = {0};
    if (cap[ctx->device] < floats) {
        if (buf[ctx->device]) { BM_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(buf[ctx->device]); }
        BM_CUDA(cudaMalloc(&buf[ctx->device], floats * sizeof(float)));
        cap[ctx->device] = floats;
    }
    return buf[ctx->device];
}
static void run_colsum_jobs(Ctx* ctx, const ColsumJobs& j) {
    int max_cols = 0;
    for (int i = 0; i < j.n; ++i) max_cols = j.cols[i] > max_cols ? j.cols[i] : max_cols;
    if (max_cols <= 0 || j.rows <= 0) return;
    max_cols = (max_cols + 1) & ~1;
    float* scratch = colsum_scratch(ctx, (size_t)3 * CS_RSPLIT * max_cols);
    colsum_bf16_partial_kernel<<<dim3((max_cols + 63) / 64, CS_RSPLIT, j.n), dim3(32, 8), 0, ctx->stream>>>(j, scratch, max_cols);
    count_launch(ctx);
    colsum_bf16_finish_kernel<<<dim3((max_cols + 255) / 256, j.n), 256, 0, ctx->stream>>>(j, scratch, max_cols);
    count_launch(ctx);
}
void launch_colsum_bf16(Ctx* ctx, const __nv_bfloat16* P, int ldp, const __nv_bfloat16* Q, int ldq,
                        int rows, int cols, float s1, float s2, float* out) {
    ColsumJobs j{};
    j.P[0] = P; j.ldp[0] = ldp; j.Q[0] = Q; j.ldq[0] = ldq; j.s1[0] = s1; j.s2[0] = s2; j.out[0] = out; j.cols[0] = cols;
    j.rows = rows; j.n = 1;
    run_colsum_jobs(ctx, j);
}
void launch_cd_statistics_bf16(Ctx* ctx, const __nv_bfloat16* X, int ldx, const __nv_bfloat16* v, int ldv,
                               const __nv_bfloat16* h0, const __nv_bfloat16* hk, int ldh, int rows, int V, int H,
                               float* dvb_sum, float* dhb_sum, float* q_sum) {
    ColsumJobs j{};
    j.P[0] = X;  j.ldp[0] = ldx; j.Q[0] = v;  j.ldq[0] = ldv; j.s1[0] = 1.f; j.s2[0] = -1.f; j.out[0] = dvb_sum; j.cols[0] = V;   // base_rbm.py:451
    j.P[1] = h0; j.ldp[1] = ldh; j.Q[1] = hk; j.ldq[1] = ldh; j.s1[1] = 1.f; j.s2[1] = -1.f; j.out[1] = dhb_sum; j.cols[1] = H;   // :453
    j.P[2] = hk; j.ldp[2] = ldh; j.Q[2] = nullptr; j.ldq[2] = 0; j.s1[2] = 1.f; j.s2[2] = 0.f; j.out[2] = q_sum; j.cols[2] = H;   // :457
    j.rows = rows; j.n = 3;
    run_colsum_jobs(ctx, j);
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
