// Fused "layer ops" on the 5th-generation tensor cores (sm_100a only), as a persistent
// multi-phase dataflow kernel.
//
//   C[M,N] = sum_p (+/-) A_p * B_p^T          bf16 operands, fp32 accumulation in TMEM
//   out    = sample(act(acc_scale * C * sigma + bias_scale * bias))
//
// One such op replaces, per Gibbs half-step, the reference's tf.matmul + 3 element-wise kernels +
// the random_uniform/Less/Cast sampling chain (rbm/base_rbm.py:329-365, layers.py:34-51); the dW op
// replaces the two gradient GEMMs (base_rbm.py:447-448) by ONE GEMM over the concatenated batch
// dimension whose negative phase is subtracted by the MMA's a_negate bit.  A *program* strings the
// 2k+1 half-steps of a CD-k chain and its dW into one launch: units (op, row-block pair, column
// block) are walked in a fixed global order by all CTA pairs and start as soon as the row blocks they
// read are complete (per-row-block counters in global memory), so nothing waits on a kernel boundary,
// and the epilogue of one half-step overlaps the MMAs of the next.  Activations pass between ops
// through L2-resident bf16 buffers (8 MB at batch 4096: never evicted to HBM by the 126 MB L2).
//
// Per CTA (384 threads, one CTA per SM; CL = 2: CTA pairs, tcgen05 cta_group::2, 256-row tiles):
//   warp 0     TMA producer: the warp's lanes issue the 2-6 bulk-tensor copies of a stage in one
//              instruction; 128B-swizzled tiles; mbarrier ring of 6 x 32 KiB (pair) / 4 x 48 KiB
//   warp 1     MMA issuer (pair leader only): one thread issues tcgen05.mma kind::f16, M=128/256, N<=256, K=16
//   warp 2     TMEM allocator (512 columns = 2 accumulator stages)
//   warps 4-11 epilogue: tcgen05.ld 32x32b.x32 (thread = one accumulator row, 32 columns), bias,
//              sigmoid by ex2/rcp, Philox-4x32-10 in registers, bf16/fp32 stores; then the unit's
//              completion is published (fence + atomic) for the ops that read it.
// Why pairs: one SM ingests only ~50 B/clk from L2 (measured), a 128x256 tile needs 96 B/clk at full
// MMA rate; with cta_group::2 each SM loads its 128 A rows and HALF of the B tile (64 B/clk).
// Operand layouts: K-major and MN-major shared-memory descriptors are both used so that a single
// bf16 copy of W serves v->h (W as MN-major B) and h->v (W as K-major B) and no activation is ever
// transposed in memory (dW takes X and h as MN-major A and B).
#include "bm_tc.h"
#include "bm_tc_desc.h"
#include <cuda.h>
#include <map>
#include <tuple>
#include <mutex>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <utility>

namespace bm {

constexpr int BM = 128;            // rows per CTA tile = TMEM lanes
constexpr int BK = 64;             // K per pipeline stage = one 128-byte swizzle atom of bf16
constexpr int MAX_STAGES = 10;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;      // TMEM columns per accumulator stage
constexpr int A_BYTES = BM * BK * 2;           // 16 KiB
constexpr int B_BYTES = 256 * BK * 2;          // 32 KiB (BN <= 256)
constexpr int RING_BYTES = 160 * 1024;                // 5 x 32 KiB pair stages / 3 x 48 KiB single-CTA stages
constexpr int N_OUT_BUF = 4;                          // bf16 output staging FIFO: 64-column x 128-row granules
constexpr int OUT_BUF_BYTES = BM * 64 * 2;            // 16 KiB each, 128B-swizzled like the operand tiles
constexpr int OUT_BYTES = N_OUT_BUF * OUT_BUF_BYTES;
constexpr int SMEM_BARRIER_BYTES = 320;   // full/empty ring, tfull/tempty, TMEM slot, unit/granule barriers
constexpr int GRAN_COLS = 64;             // bf16 outputs leave the SM (and are published) in 64-column granules
constexpr int EPI_WARPS = 16;        // 4 per TMEM lane quarter: the epilogue is latency-bound, it needs warps
constexpr int EPI_SUBS = EPI_WARPS / 4;
constexpr int CW = 16;               // accumulator columns per epilogue chunk (one tcgen05.ld.x16)
constexpr int TC_THREADS = 32 * (4 + EPI_WARPS);
// Warp roles.  The SM's issue arbiter prefers the highest warp id: the single-thread roles whose latency
// is on the critical path (TMA producer, MMA issuer, publishers) get the highest ids so that sixteen busy
// (or spinning) epilogue warps cannot starve them.
constexpr int W_PUB0 = EPI_WARPS, W_PUB1 = EPI_WARPS + 1, W_MMA = EPI_WARPS + 2, W_TMA = EPI_WARPS + 3;

enum : int {
    MODE_GENERIC = 0,                // every combination, flags read at run time
    MODE_SIG_BERN_MEAN_STATE = 1,    // h0: sigmoid, Bernoulli draw, bf16 means + bf16 states
    MODE_SIG_BERN_STATE = 2,         // mid-chain hidden: states only
    MODE_SIG_MEAN = 3,               // probabilities only (visible means, last hidden means)
    MODE_RAW_F32 = 4,                // raw fp32 accumulators (dW partials, linear pre-activations)
    MODE_AIS_UNITS = 5,              // AIS: importance-weight increment of the row + sampled unit updates of the next transition
    MODE_AIS_STATE = 6,              // AIS: sampled transition (sigmoid, Bernoulli, states) + the linear term of log p*
    MODE_AIS_UNITS_S = 7,            // AIS units op without a weight increment (first transition, sweeps > 0)
    MODE_AIS_UNITS_G = 8             // AIS units op, run-time flags (unsampled units, the ladder's last increment, far-apart temperatures)
};

constexpr int MAX_PHASES = 96;
// The series form of the AIS increment (chunk_body, MODE_AIS_UNITS) is used for ladders of >= 400 temperatures: with
// t = (b - a) z and d = (next - (a + b) / 2) z its errors are t^2/24 sigma''/sigma (midpoint rule) and d^4/24 sigma''''/sigma
// (third-order series) relative to the term -- < 3e-6 and < 1e-8 for |z| <= 10 at these spacings, and both vanish with
// sigma' for large |z|.  Coarser ladders take the closed form (MODE_AIS_UNITS_G).
constexpr float AIS_MAX_CT = 2.5e-3f, AIS_MAX_CD = 4.0e-3f;
constexpr int SBIAS_BYTES = ACC_STAGES * 256 * (int)sizeof(float);

constexpr int SMEM_BYTES = RING_BYTES + OUT_BYTES + SMEM_BARRIER_BYTES + SBIAS_BYTES;     // 231,744 of 232,448

// The ops of the running launch.  Constant memory on purpose: indexed by warp-uniform values it is read
// through the uniform datapath, so the TMA / MMA issue loops keep their descriptors, coordinates and
// trip counts in uniform registers -- with per-thread registers every UTMALDG / UTCHMMA is wrapped in
// an ELECT + R2UR "waterfall" loop of ~100 cycles, which made the issue loops the bottleneck.
__constant__ TcPhaseLite c_ph[MAX_PHASES];


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
// Every wait of the kernel is bounded: a wait that has not been satisfied after HANG_NS of wall time (a lost
// arrival, a peer that never launched, stale dataflow counters) traps -- the launch fails with an error instead
// of holding the GPU for ever.  The clock is only read every 4096 failed attempts.
constexpr unsigned long long HANG_NS = 4000000000ull;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void hang_guard(uint32_t& n, unsigned long long& t0) {
    if ((++n & 0xFFFu) == 0) {
        const unsigned long long t = gtime();
        if (t0 == 0) t0 = t;
        else if (t - t0 > HANG_NS) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    uint32_t n = 0; unsigned long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        hang_guard(n, t0);
    }
}
// for waits that are expected to be long (epilogue warps waiting for an accumulator or a staging
// buffer): sleep between attempts instead of spinning in the issue slots of the busy roles
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns = 128) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    uint32_t n = 0; unsigned long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
        hang_guard(n, t0);
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, uint64_t map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// cta_group::2 loads: data lands in the issuing CTA, the transaction bytes are counted on the
// LEADER CTA's mbarrier (address with the peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, uint64_t map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(uint64_t map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// release: everything this thread has observed (completed bulk stores, barrier-acquired stores of the
// epilogue warps) is visible to whoever acquires the counter
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
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
// One lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas knows that exactly one thread
// executes the guarded code and moves its operands to uniform registers without an ELECT/R2UR loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}
#define DBG_MARK(slot) do { if (L.dbg && blockIdx.x == 0) L.dbg[(slot)] = (unsigned long long)clock64(); } while (0)
// per-unit marks of CTA 0: slot 64 + ord * 8 + kind (ord = ordinal of the unit within this CTA, < 24)
#define DBG_UNIT(kind, ord) do { if (L.dbg && blockIdx.x == 0 && (ord) < 24) L.dbg[64 + (ord) * 8 + (kind)] = (unsigned long long)clock64(); } while (0)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// descriptors are passed as (lo, hi) halves: only `lo` (the start address) changes between MMAs
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {      // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
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

__device__ __forceinline__ float fast_softplus(float x) { return fmaxf(x, 0.f) + __logf(1.0f + __expf(-fabsf(x))); }
// sigmoid(x) = 1 / (1 + 2^(-x log2 e)): ex2.approx + rcp.approx keep the *relative* error of small
// probabilities at ~1e-7 (tanh.approx would not)
__device__ __forceinline__ float sigmoid_from_neg_log2(float t) {   // t = -x * log2(e)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
// asfloat((word & 0x7fffff) | 0x3f800000) in [1, 2): TF's Uint32ToFloat before its "- 1.0f"
__device__ __forceinline__ float u32_to_one_two(uint32_t w, uint32_t mant_mask, uint32_t one_bits) {
    uint32_t bits;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(bits) : "r"(w), "r"(mant_mask), "r"(one_bits));
    return __uint_as_float(bits);
}
// 256-bit store (sm_100+): a thread's 8 consecutive fp32 go out as one full 32-byte sector
__device__ __forceinline__ void stg256(float* dst, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h2);
}
// Polling uses relaxed loads: an acquire load makes ptxas emit CCTL.IVALL (a full L1 invalidation) on
// every iteration, which evicted the epilogue warps' bias/constant lines for as long as a producer
// warp was waiting.  One acquire fence follows the successful poll.
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// ------------------------------------------------------------------------------------------
// epilogue of one tile, specialised per output mode
// ------------------------------------------------------------------------------------------
template <int MODE> struct EpiCfg {
    static constexpr bool fixed = MODE != MODE_GENERIC;
    static constexpr bool ais_units = (MODE == MODE_AIS_UNITS || MODE == MODE_AIS_UNITS_S || MODE == MODE_AIS_UNITS_G);
    static constexpr bool ais = ais_units || MODE == MODE_AIS_STATE;
    static constexpr int act = (MODE == MODE_RAW_F32 || ais_units) ? ACT_LINEAR : ACT_SIGMOID;
    static constexpr int sample = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_BERN_STATE || MODE == MODE_AIS_STATE) ? SMP_BERNOULLI : SMP_NONE;
    static constexpr bool mean_bf = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_MEAN);
    static constexpr bool state_bf = (MODE == MODE_SIG_BERN_MEAN_STATE || MODE == MODE_SIG_BERN_STATE || MODE == MODE_AIS_STATE);
    static constexpr bool f32 = (MODE == MODE_RAW_F32);
};

// the epilogue's view of a phase, copied into registers once per unit (the descriptor itself lives in
// global memory; reading it inside the element loops would put L2 latency on every use)
struct EpiPhase {
    int M, N, BN, act, sample;
    float acc_scale, bias_scale;
    const float* __restrict__ bias;
    const float* __restrict__ sigma;
    const float* __restrict__ noise_sigma;
    __nv_bfloat16* __restrict__ out_mean_bf;  int ld_mean_bf;
    __nv_bfloat16* __restrict__ out_state_bf; int ld_state_bf;
    float* __restrict__ out_f32;              int ld_f32;
    unsigned long long split_stride;
    float ais_a, ais_b, ais_next, ais_lin;
    float ais_ct, ais_cd;          // (b - a) and (next - (a + b) / 2): z-multipliers of the series form of the increment
    double* ais_logw;
};
struct EpiCtx {
    EpiPhase p;
    RngKey rng;
    int m, n_blk, split;          // this thread's global row, the tile's column block and K split
    uint32_t t_row;               // TMEM address of this thread's lane, column 0 of the accumulator stage
    uint64_t* tempty;             // accumulator-free barrier (leader's in pair mode)
    int sub, lane;                // sub: which chunks (mod EPI_SUBS) this warp takes
    const float* sbias;           // shared memory: bias_scale * bias (x -log2 e for sigmoid) of the tile's columns
    bool remote_arrive;           // pair mode, peer CTA: signal the leader's barrier
    // bf16 outputs are staged in shared memory and leave the SM as TMA stores (publisher warp)
    uint32_t out_base;            // shared address of the staging FIFO
    uint32_t row_off;             // this thread's row inside a staging buffer: (row % 128) * 128 bytes
    uint32_t row_swz;             // row & 7: XOR pattern of the 128-byte swizzle
    uint64_t* out_full;           // [N_OUT_BUF] a staged granule is complete (16 warp arrivals)
    uint64_t* out_free;           // [N_OUT_BUF] its TMA store has read the buffer (publisher)
    uint32_t* seq;                // FIFO position (same sequence in every warp and in the publisher)
    uint32_t mant_mask, one_bits; // TcLaunch::mant_mask / one_bits
};

// One CW-column chunk of one accumulator row: activation, sampling, then
//   bf16 means / states -> this thread's row of the staging buffers (two 16-byte units each),
//   fp32 -> global memory directly (dW partials and raw pre-activations only).
// softplus(b z) - softplus(a z) = log1p(sigmoid(a z) * expm1((b - a) z)) without cancellation (b - a ~ 1e-3): short series
// where both arguments are small (the common case), the library functions elsewhere
__device__ __forceinline__ float ais_softplus_diff(float a, float b, float z) {
    const float sa = sigmoid_from_neg_log2(a * z * -1.4426950408889634f);
    const float t = (b - a) * z;
    float em;
    if (fabsf(t) < 0.03f) em = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.0f / 120.0f, 1.0f / 24.0f), 1.0f / 6.0f), 0.5f), 1.0f);
    else em = expm1f(t);
    const float y = sa * em;
    if (fabsf(y) < 0.03f) return y * fmaf(y, fmaf(y, fmaf(y, fmaf(y, 0.2f, -0.25f), 1.0f / 3.0f), -0.5f), 1.0f);
    return log1pf(y);
}

// AIS_UNITS only: the chunk's contribution to the row's log-weight; AIS_STATE: sum of state * pre-scaled bias; else 0
template <int MODE>
__device__ __forceinline__ float chunk_body(const EpiCtx& c, const uint32_t (&v)[CW], int ch, int n0, int n_valid,
                                            uint32_t smem_mean, uint32_t smem_state, bool store_ok) {
    typedef EpiCfg<MODE> E;
    const EpiPhase& p = c.p;
    float ais_part = 0.f;
    const int act = E::fixed ? E::act : p.act;
    const int smp = (MODE == MODE_AIS_UNITS || MODE == MODE_AIS_UNITS_S) ? SMP_BERNOULLI :
                    (E::fixed && MODE != MODE_AIS_UNITS_G) ? E::sample : p.sample;
    float* const out_f32 = ((E::fixed && !E::f32) || !p.out_f32) ? nullptr : p.out_f32 + (size_t)c.split * p.split_stride;
    // fold the sigmoid's -log2(e) into the affine map of the accumulator
    const float a_s = (act == ACT_SIGMOID) ? p.acc_scale * -1.4426950408889634f : p.acc_scale;
    const bool has_sigma = !E::fixed && p.sigma != nullptr;
    const int m = c.m;
    // in the fixed modes which outputs exist is known at compile time (no per-group branches)
    const bool do_mean = E::fixed ? E::mean_bf : (p.out_mean_bf != nullptr);
    const bool do_state = (MODE == MODE_AIS_UNITS || MODE == MODE_AIS_UNITS_S) ? true :
                          (E::fixed && MODE != MODE_AIS_UNITS_G) ? E::state_bf : (p.out_state_bf != nullptr);
    const bool do_f32 = E::fixed ? E::f32 : (out_f32 != nullptr);
    const bool f32_vec = n_valid == CW && (p.ld_f32 & 3) == 0;       // rows 16-byte aligned, whole chunk
    // rows 32-byte aligned (row pitch, split stride and base): one full sector per store instruction and thread
    const bool f32_wide = f32_vec && (p.ld_f32 & 7) == 0 && (p.split_stride & 7) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.out_f32) & 31) == 0;
    float f32_prev[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t mean_pk[CW / 2], state_pk[CW / 2];
#pragma unroll
    for (int q = 0; q < CW / 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(c.sbias + ch * CW + q * 4);   // broadcast LDS.128
        const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
        U4 w{0, 0, 0, 0};
        if (smp != SMP_NONE) w = site_block(c.rng, (uint32_t)m, (uint32_t)((n0 >> 2) + q));
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
            x += bq[j];
            float m_ = x;
            float s_;
            if (MODE == MODE_AIS_UNITS || MODE == MODE_AIS_UNITS_S) {
                // x = z (acc_scale = bias_scale = 1).  p = sigmoid(beta_next z) serves the draw AND the weight increment:
                //   softplus(b z) - softplus(a z) = int_a^b z sigmoid(beta z) dbeta = t sigmoid(m z)  (+ t^3 z^-2 ... / 24: < 1e-7 t)
                //   with t = (b - a) z, m = (a + b) / 2, and sigmoid(m z) = sigmoid(beta_next z - d), d = (beta_next - m) z, as its
                //   third-order series in d around p (error terms: AIS_MAX_CT / AIS_MAX_CD above).
                float e_, p_;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e_) : "f"(x * p.ais_next * -1.4426950408889634f));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p_) : "f"(1.0f + e_));
                s_ = ((u32_to_one_two(words[j], c.mant_mask, c.one_bits) - 1.0f) < p_) ? 1.0f : 0.0f;
                if (MODE == MODE_AIS_UNITS) {
                    const float q_ = fmaf(-p_, p_, p_);                                 // sigma1 = p (1 - p)
                    const float d_ = p.ais_cd * x;
                    const float s2 = fmaf(-2.0f, p_, 1.0f);                             // sigma2 / sigma1
                    const float s3 = fmaf(-6.0f, q_, 1.0f);                             // sigma3 / sigma1
                    const float in = fmaf(-d_ * (1.0f / 3.0f), s3, s2);
                    const float sm = fmaf(-d_ * q_, fmaf(-0.5f * d_, in, 1.0f), p_);    // sigmoid(m z)
                    ais_part = fmaf((e < n_valid) ? p.ais_ct * x : 0.f, sm, ais_part);
                }
                mu[j] = 0.f; st[j] = s_;
                continue;
            }
            if (MODE == MODE_AIS_UNITS_G) {
                // any temperatures, optional sampling / output: the closed forms
                if (p.ais_logw && e < n_valid) ais_part += ais_softplus_diff(p.ais_a, p.ais_b, x);
                float e_;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e_) : "f"(x * p.ais_next * -1.4426950408889634f));
                const float d_ = 1.0f + e_;
                if (smp == SMP_BERNOULLI) {
                    const float u12 = u32_to_one_two(words[j], c.mant_mask, c.one_bits);
                    s_ = (fmaf(u12, d_, -d_) < 1.0f) ? 1.0f : 0.0f;
                } else {
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s_) : "f"(d_));
                }
                mu[j] = 0.f; st[j] = s_;
                continue;
            }
            if (MODE == MODE_SIG_BERN_STATE || MODE == MODE_AIS_STATE) {
                // only the draw is needed: u < 1/(1+e)  <=>  u*(1+e) < 1, with u = u12 - 1 folded into one FMA
                // (no reciprocal, no "- 1.0f"); e = inf (p = 0) gives NaN -> 0, e = 0 (p = 1) gives u < 1 -> 1
                float e_;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e_) : "f"(x));
                const float d_ = 1.0f + e_;
                const float u12 = u32_to_one_two(words[j], c.mant_mask, c.one_bits);
                s_ = (fmaf(u12, d_, -d_) < 1.0f) ? 1.0f : 0.0f;
                mu[j] = 0.f; st[j] = s_;
                if (MODE == MODE_AIS_STATE && e < n_valid) ais_part = fmaf(s_, bq[j], ais_part);
                continue;
            }
            if (act == ACT_SIGMOID) m_ = sigmoid_from_neg_log2(x);
            else if (!E::fixed && act == ACT_SOFTPLUS) m_ = fast_softplus(x);
            s_ = m_;
            if (smp == SMP_BERNOULLI) s_ = ((u32_to_one_two(words[j], c.mant_mask, c.one_bits) - 1.0f) < m_) ? 1.0f : 0.0f;
            else if (!E::fixed && smp == SMP_GAUSSIAN)
                s_ = m_ + ((p.noise_sigma && e < n_valid) ? p.noise_sigma[n0 + e] : 1.0f) * g[j];
            mu[j] = m_; st[j] = s_;
        }
        if (do_mean) { mean_pk[2 * q] = pack_bf16(mu[0], mu[1]); mean_pk[2 * q + 1] = pack_bf16(mu[2], mu[3]); }
        if (do_state) { state_pk[2 * q] = pack_bf16(st[0], st[1]); state_pk[2 * q + 1] = pack_bf16(st[2], st[3]); }
        if (do_f32 && store_ok) {
            float* dst = out_f32 + (size_t)m * p.ld_f32 + n0 + q * 4;
            if (f32_wide) {                       // 32-byte stores: every other group carries the previous one with it
                if (q & 1) { const float v8[8] = {f32_prev[0], f32_prev[1], f32_prev[2], f32_prev[3], mu[0], mu[1], mu[2], mu[3]}; stg256(dst - 4, v8); }
                else { f32_prev[0] = mu[0]; f32_prev[1] = mu[1]; f32_prev[2] = mu[2]; f32_prev[3] = mu[3]; }
            } else if (f32_vec) *reinterpret_cast<float4*>(dst) = make_float4(mu[0], mu[1], mu[2], mu[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (q * 4 + j < n_valid) dst[j] = mu[j];
            }
        }
    }
    // columns past N and rows past M are written to the staging buffer too: the TMA store clips them
    const uint32_t u0 = (uint32_t)((ch * CW) & 63) >> 3;           // first 16-byte unit of this chunk in the 128-byte row
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) {
        const uint32_t off = c.row_off + (((u0 + (uint32_t)i) ^ c.row_swz) << 4);
        if (do_mean) sts128(smem_mean + off, mean_pk[4 * i], mean_pk[4 * i + 1], mean_pk[4 * i + 2], mean_pk[4 * i + 3]);
        if (do_state) sts128(smem_state + off, state_pk[4 * i], state_pk[4 * i + 1], state_pk[4 * i + 2], state_pk[4 * i + 3]);
    }
    return ais_part;
}

// One warp's share of a tile: TMEM lane quarter (warp % 4) x the CW-column chunks ch = sub (mod EPI_SUBS).
// All warps sweep the tile left to right together, granule (64 columns) by granule: each granule of each
// bf16 output takes the next buffer of the staging FIFO, is written by the 16 warps and announced on the
// buffer's `full` barrier; the publisher warp stores it with TMA and publishes it to the other SMs.
template <int MODE, bool PAIR>
__device__ __forceinline__ void epilogue_tile(const EpiCtx& c) {
    typedef EpiCfg<MODE> E;
    const EpiPhase& p = c.p;
    const int BN = p.BN;
    const int n_chunks = BN / CW;                  // BN is a multiple of 16
    const bool row_ok = c.m < p.M;
    const bool do_mean = E::fixed ? E::mean_bf : (p.out_mean_bf != nullptr);
    const bool do_state = (MODE == MODE_AIS_UNITS || MODE == MODE_AIS_UNITS_S) ? true :
                          (E::fixed && MODE != MODE_AIS_UNITS_G) ? E::state_bf : (p.out_state_bf != nullptr);
    float ais_sum = 0.f;
    int last_ch = -1;
    for (int ch = c.sub; ch < n_chunks; ch += EPI_SUBS) last_ch = ch;
    if (last_ch < 0) {          // this warp has no chunk in the tile: release the accumulator at once
        __syncwarp();
        if (c.lane == 0) {
            if (PAIR && c.remote_arrive) { if constexpr (PAIR) mbar_arrive_remote(c.tempty, 0); } else mbar_arrive(c.tempty);
        }
    }
    constexpr int gc = GRAN_COLS / CW;             // chunks per granule
    const int n_gran = (n_chunks + gc - 1) / gc;
    uint32_t seq = *c.seq;
    for (int g = 0; g < n_gran; ++g) {
        uint32_t smem_mean = 0, smem_state = 0;
        int b_mean = 0, b_state = 0;
        if (do_mean) {
            b_mean = (int)(seq % N_OUT_BUF);
            mbar_wait_relaxed(&c.out_free[b_mean], ((seq / N_OUT_BUF) & 1u) ^ 1u);      // first use of a buffer passes at once
            smem_mean = c.out_base + (uint32_t)b_mean * OUT_BUF_BYTES; ++seq;
        }
        if (do_state) {
            b_state = (int)(seq % N_OUT_BUF);
            mbar_wait_relaxed(&c.out_free[b_state], ((seq / N_OUT_BUF) & 1u) ^ 1u);
            smem_state = c.out_base + (uint32_t)b_state * OUT_BUF_BYTES; ++seq;
        }
        const int ch_end = min(n_chunks, (g + 1) * gc);
        for (int ch = g * gc + c.sub; ch < ch_end; ch += EPI_SUBS) {
            uint32_t v[CW];
            __syncwarp();                            // tcgen05.ld is warp-collective (.sync.aligned)
            tmem_ld16(c.t_row + (uint32_t)(ch * CW), v);
            tmem_ld_wait();
            if (ch == last_ch) {
                // all of this warp's reads of the accumulator are done: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (c.lane == 0) {
                    if (PAIR && c.remote_arrive) { if constexpr (PAIR) mbar_arrive_remote(c.tempty, 0); } else mbar_arrive(c.tempty);
                }
            }
            const int n0 = c.n_blk * BN + ch * CW;
            if (n0 < p.N) ais_sum += chunk_body<MODE>(c, v, ch, n0, min(CW, p.N - n0), smem_mean, smem_state, row_ok);
        }
        if (do_mean || do_state) {
            fence_proxy_async_smem();                // this thread's staging writes -> visible to the TMA unit
            __syncwarp();
            if (c.lane == 0) {
                if (do_mean) mbar_arrive(&c.out_full[b_mean]);
                if (do_state) mbar_arrive(&c.out_full[b_state]);
            }
        }
    }
    *c.seq = seq;
    if constexpr (E::ais) {
        // this thread's share of its row's log-weight: fp64 atomics (four threads per row and column tile)
        if (p.ais_logw && row_ok) {
            const double inc = (double)(MODE == MODE_AIS_STATE ? ais_sum * p.ais_lin : ais_sum);
            asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p.ais_logw + c.m), "d"(inc) : "memory");     // (no return value wanted)
        }
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
struct UnitInfo { int split, m_group, n_blk, c_begin, c_end, total_chunks; };

// Next unit of CTA pair `cid`: ops in program order, inside an op the units cid - pair_begin (mod pair_count).
// (pi, l) = current op and local unit; l < 0: before the first unit of op pi.  All values are warp-uniform.
__device__ __forceinline__ bool walk_next(int& pi, int& l, int cid, int n_ops) {
    while (pi < n_ops) {
        const TcPhaseLite* p = &c_ph[pi];
        const int rel = cid - p->pair_begin;
        if (rel >= 0 && rel < p->pair_count) {
            l = (l < 0) ? rel : l + p->pair_count;
            if (l < p->n_units) return true;
        }
        ++pi; l = -1;
    }
    return false;
}
__device__ __forceinline__ UnitInfo decode_unit(const TcPhaseLite* ph, int local) {
    UnitInfo u;
    u.split = local % ph->splits;
    const int tile = local / ph->splits;
    u.m_group = tile / ph->n_tiles;
    u.n_blk = tile % ph->n_tiles;
    u.total_chunks = ph->chunks[0] + (ph->n_pairs > 1 ? ph->chunks[1] : 0);
    u.c_begin = (int)(((long long)u.total_chunks * u.split) / ph->splits);
    u.c_end = (int)(((long long)u.total_chunks * (u.split + 1)) / ph->splits);
    return u;
}

template <int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_program_kernel(const __grid_constant__ TcLaunch L) {
    constexpr bool pair = (CL == 2);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;                      // [operand ring | output staging FIFO | barriers | bias]
    if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzle atoms need 1024-byte alignment
    uint8_t* const out_stage = smem + RING_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING_BYTES + OUT_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);
    uint64_t* unit_bar = tempty + ACC_STAGES + 1;                 // [ACC_STAGES] all epilogue warps stored the unit
    uint64_t* out_full = unit_bar + ACC_STAGES;                   // [N_OUT_BUF] staged granule written by all epilogue warps
    uint64_t* out_free = out_full + N_OUT_BUF;                    // [N_OUT_BUF] staged granule read by its TMA store

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { DBG_MARK(0); if (L.dbg && blockIdx.x == 0) L.dbg[6] = gtime(); }
    float* const s_bias = reinterpret_cast<float*>(smem + RING_BYTES + OUT_BYTES + SMEM_BARRIER_BYTES);
    const TcPhase* const gph = L.n_phases ? L.phases : &L.inl;      // tensor maps (global / parameter memory)
    if (warp == W_TMA && lane == 0) {
        tma_prefetch_desc(&gph->tmA[0]); tma_prefetch_desc(&gph->tmB[0]);
    }
    if (warp == W_MMA && lane == 0) {
        // pair: the leader's `full` collects its own expect_tx-arrive and the peer's arrive; its `tempty`
        // collects the epilogue warps of both CTAs; `empty`/`tfull` get one multicast commit each
        for (int s = 0; s < L.stages; ++s) { mbar_init(&full[s], (uint32_t)CL); mbar_init(&empty[s], 1); }
        // tempty: the epilogue warps and the two publisher warps of both CTAs
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], (uint32_t)((EPI_WARPS + 2) * CL)); }
        for (int a = 0; a < ACC_STAGES; ++a) mbar_init(&unit_bar[a], EPI_WARPS);
        for (int b = 0; b < N_OUT_BUF; ++b) { mbar_init(&out_full[b], EPI_WARPS); mbar_init(&out_free[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_PUB0) {
        if constexpr (pair) {        // one warp of each CTA of the pair allocates collectively
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    if constexpr (pair) cluster_sync_all(); else __syncthreads();   // peers' barriers must exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) DBG_MARK(1);

    const int n_ops = L.n_phases ? L.n_phases : 1;
    // clusters are (2,1,1): rank = blockIdx.x & 1, cluster id = blockIdx.x >> 1 (kept as expressions of
    // blockIdx / gridDim so that the compiler knows they are warp-uniform)
    const int crank = pair ? (int)(blockIdx.x & 1u) : 0;
    const int cid = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;

    if (warp == W_TMA) {
        // ================================ TMA producer =====================================
        // The whole warp walks the K chunks; for each chunk lane 0 arms the barrier and lanes
        // 0..n_ops-1 issue one bulk-tensor copy each in the same warp instruction.
        int stage = 0; uint32_t phase = 0;
        int pi = 0, ul = -1;
        int ord = -1;
        while (walk_next(pi, ul, cid, n_ops)) {
            ++ord;
            const TcPhaseLite* ph = &c_ph[pi];
            const TcPhase* gp = &gph[pi];
            const UnitInfo u = decode_unit(ph, ul);
            if (lane == 0) DBG_UNIT(0, ord);
            // this op's tensor maps -> descriptor cache while the unit still waits for its inputs
            if ((L.flags & 8) && lane < 2 * ph->n_pairs) tma_prefetch_desc(lane & 1 ? &gp->tmB[lane >> 1] : &gp->tmA[lane >> 1]);
            const int BN = ph->BN;
            const int half_bn = BN >> 1;
            const uint32_t tx_bytes = pair ? 2u * (A_BYTES + (uint32_t)half_bn * BK * 2) : A_BYTES + (uint32_t)BN * BK * 2;
            const int b_cols = pair ? half_bn : BN;          // B columns this CTA fetches
            const int m_blk = u.m_group * CL + crank;
            const int n_col0 = u.n_blk * BN + (pair ? crank * half_bn : 0);
            // ---- dataflow: wait until the row blocks this unit reads have been written ----------
            if (ph->n_deps > 0) {
                if (lane == 0) {
                    uint32_t hn = 0; unsigned long long ht0 = 0;
                    for (int d = 0; d < ph->n_deps; ++d) {
                        const int* ctr = ph->dep_ctr[d];
                        const int need = ph->dep_need[d] * L.epoch;
                        if (ph->dep_groups[d] == 0) {
                            if (ph->dep_chunk_ctr) continue;          // handled chunk by chunk below
                            while (ld_relaxed(ctr + u.m_group) < need) { __nanosleep(64); hang_guard(hn, ht0); }
                        } else {
                            for (int gq = 0; gq < ph->dep_groups[d]; ++gq)
                                while (ld_relaxed(ctr + gq) < need) { __nanosleep(64); hang_guard(hn, ht0); }
                        }
                    }
                    // acquire the producers' (generic-proxy) stores, then order them before this unit's
                    // TMA reads, which go through the async proxy
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                __syncwarp();
            }
            if (lane == 0) DBG_UNIT(1, ord);
            const int* const cdep = ph->dep_chunk_ctr ? ph->dep_chunk_ctr + (size_t)u.m_group * ph->dep_gran_row : nullptr;
            // ---- per-unit setup (warp-uniform: uniform registers): operand-pair configuration -------
            const int chunks0 = ph->chunks[0];
            uint64_t mapA[2], mapB[2];
            int a_c0[2], a_c1[2], b_c0[2], b_c1[2], a_mn[2], b_mn[2];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                const int shift = ph->a_batch[pr] ? L.batch_row : 0;
                a_mn[pr] = ph->a_mn[pr]; b_mn[pr] = ph->b_mn[pr];
                mapA[pr] = (uint64_t)&gp->tmA[pr]; mapB[pr] = (uint64_t)&gp->tmB[pr];
                // box coordinates at K offset 0: K-major {k, row}; MN-major {mn, k}
                a_c0[pr] = a_mn[pr] ? m_blk * BM : 0;
                a_c1[pr] = a_mn[pr] ? ph->a_k0[pr] + shift : ph->a_row0[pr] + shift + m_blk * BM;
                b_c0[pr] = b_mn[pr] ? n_col0 : 0;
                b_c1[pr] = b_mn[pr] ? 0 : n_col0;
            }
            const int nB_mn = b_cols >> 6;            // 64-column boxes of an MN-major B tile
            const uint32_t smem_base = smem_u32(smem);
            const bool ordered = cdep != nullptr;
            int fenced_upto = u.c_begin;         // K positions [c_begin, fenced_upto) are known ready and fenced
            for (int ci = u.c_begin; ci < u.c_end; ++ci) {
                int c = ci;
                if (ordered) {
                    c = (int)ph->k_order[ci];
                    if (ci >= fenced_upto) {
                        // every lane polls one K position (ci + lane): one L2 round trip tells how long the
                        // prefix of ready positions is (a serial scan costs a round trip per position)
                        const int need = ph->dep_chunk_need * L.epoch;
                        const int n = ci + lane;
                        int ga = 0, gb = 0;
                        if (n < u.c_end) { const int j = (int)ph->k_order[n]; ga = (int)ph->k_dep_a[j]; gb = (int)ph->k_dep_b[j]; }
                        int ready_prefix = 0;
                        uint32_t hn = 0; unsigned long long ht0 = 0;
                        for (;;) {
                            bool ok;
                            if (L.flags & 1) ok = n < u.c_end && ld_acquire(cdep + ga) >= need && (gb == ga || ld_acquire(cdep + gb) >= need);
                            else ok = n < u.c_end && ld_relaxed(cdep + ga) >= need && (gb == ga || ld_relaxed(cdep + gb) >= need);
                            const uint32_t m = __ballot_sync(0xffffffffu, ok);
                            ready_prefix = __ffs(~m) - 1;              // ~m != 0: lanes past c_end are never ready... (32 positions max)
                            if (m == 0xffffffffu) ready_prefix = 32;
                            if (ready_prefix > 0) break;
                            __nanosleep((uint32_t)L.poll_ns);
                            hang_guard(hn, ht0);
                        }
                        fenced_upto = ci + ready_prefix;
                        if (L.dbg && blockIdx.x == 0 && ord == 1 && lane == 0 && ci - u.c_begin < 20) { L.dbg[456 + ci - u.c_begin] = (unsigned long long)clock64(); L.dbg[480 + ci - u.c_begin] = (unsigned long long)fenced_upto; }
                        // acquire the producers' stores, then order them before this unit's TMA reads (async proxy)
                        if (!(L.flags & 1)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                        // The producers' data was written by TMA stores and is read here by TMA loads (the same
                        // proxy); only the counter travels through the generic proxy.  flags bit 1 narrows the
                        // cross-proxy fence to the global state space, bit 2 drops it (the TMA issue below is
                        // control-dependent on the acquired counter value) -- a full fence.proxy.async also waits
                        // for this thread's own in-flight bulk copies (~1.2 k cycles per poll, measured).
                        if (L.flags & 4) {}
                        else if (L.flags & 2) asm volatile("fence.proxy.async.global;" ::: "memory");
                        else asm volatile("fence.proxy.async;" ::: "memory");
                        if (L.dbg && blockIdx.x == 0 && ord == 1 && lane == 0 && ci - u.c_begin < 20) L.dbg[432 + ci - u.c_begin] = (unsigned long long)clock64();
                    }
                }
                const int pr = (c >= chunks0) ? 1 : 0;
                const int kc = (c - (pr ? chunks0 : 0)) * BK;
                const uint32_t sA = smem_base + (uint32_t)stage * (uint32_t)L.stage_bytes;
                const uint32_t sB = sA + (uint32_t)A_BYTES;
                const uint64_t mA = pr ? mapA[1] : mapA[0], mB = pr ? mapB[1] : mapB[0];
                const int amn = pr ? a_mn[1] : a_mn[0], bmn = pr ? b_mn[1] : b_mn[0];
                const int ac0 = pr ? a_c0[1] : a_c0[0], ac1 = pr ? a_c1[1] : a_c1[0];
                const int bc0 = pr ? b_c0[1] : b_c0[0], bc1 = pr ? b_c1[1] : b_c1[0];
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {               // one thread issues; every operand above is warp-uniform
                    if (ci - u.c_begin < 24 && (L.n_phases == 0 || ord == 1)) DBG_MARK(8 + (ci - u.c_begin));
                    const uint32_t fbar = smem_u32(&full[stage]);
                    if constexpr (pair) {
                        // both CTAs load their own A rows and their half of the B tile; bytes are counted on the leader
                        const uint32_t lbar = fbar & 0xFEFFFFFFu;
                        if (crank == 0) mbar_expect_tx(&full[stage], tx_bytes); else mbar_arrive_remote(&full[stage], 0);
                        if (!amn) tma_load_2d_2sm(sA, mA, lbar, kc, ac1);
                        else { tma_load_2d_2sm(sA, mA, lbar, ac0, ac1 + kc); tma_load_2d_2sm(sA + 8192u, mA, lbar, ac0 + 64, ac1 + kc); }
                        if (!bmn) tma_load_2d_2sm(sB, mB, lbar, kc, bc1);
                        else for (int j = 0; j < nB_mn; ++j) tma_load_2d_2sm(sB + (uint32_t)j * 8192u, mB, lbar, bc0 + j * 64, kc);
                    } else {
                        mbar_expect_tx(&full[stage], tx_bytes);
                        if (!amn) tma_load_2d(sA, mA, fbar, kc, ac1);
                        else { tma_load_2d(sA, mA, fbar, ac0, ac1 + kc); tma_load_2d(sA + 8192u, mA, fbar, ac0 + 64, ac1 + kc); }
                        if (!bmn) tma_load_2d(sB, mB, fbar, kc, bc1);
                        else for (int j = 0; j < nB_mn; ++j) tma_load_2d(sB + (uint32_t)j * 8192u, mB, fbar, bc0 + j * 64, kc);
                    }
                }
                if (++stage == L.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == W_MMA) {
        // ================================ MMA issuer (the pair's leader CTA only) ============
        // The whole warp runs the loop converged (all values warp-uniform -> uniform registers);
        // lane 0 issues the tcgen05 instructions.
        if (crank == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int pi = 0, ul = -1;
            int ord = -1;
            const uint32_t smem_base = smem_u32(smem);
            while (walk_next(pi, ul, cid, n_ops)) {
                ++ord;
                const TcPhaseLite* ph = &c_ph[pi];
                const UnitInfo u = decode_unit(ph, ul);
                // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=bf16 [7,10), B=bf16 [10,13),
                // a_negate 13, a_major 15, b_major 16, N>>3 [17,23), M>>4 [24,29)
                const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ph->BN >> 3) << 17) |
                                            ((uint32_t)((pair ? 2 * BM : BM) >> 4) << 24);
                // per operand pair: instruction descriptor, K=16 slice strides and the constant upper
                // halves of the shared-memory descriptors -- the chunk loop only adds addresses
                uint32_t idesc_p[2], a_step_p[2], b_step_p[2], a_hi_p[2], b_hi_p[2], a_lo_p[2], b_lo_p[2];
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    const int a_mn = ph->a_mn[pr], b_mn = ph->b_mn[pr];
                    idesc_p[pr] = idesc_base | ((uint32_t)ph->a_neg[pr] << 13) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
                    a_step_p[pr] = a_mn ? (2048u >> 4) : (32u >> 4);    // descriptor units (16 B) per K=16 slice
                    b_step_p[pr] = b_mn ? (2048u >> 4) : (32u >> 4);
                    const uint64_t da = make_smem_desc(0, a_mn), db = make_smem_desc(0, b_mn);
                    a_hi_p[pr] = (uint32_t)(da >> 32); a_lo_p[pr] = (uint32_t)da;
                    b_hi_p[pr] = (uint32_t)(db >> 32); b_lo_p[pr] = (uint32_t)db;
                }
                const int chunks0 = ph->chunks[0];
                const bool ordered = ph->dep_chunk_ctr != nullptr;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                uint32_t accumulate = 0;
                for (int ci = u.c_begin; ci < u.c_end; ++ci) {
                    const int c = ordered ? (int)ph->k_order[ci] : ci;     // same order as the producer warp
                    const int pr = (c >= chunks0) ? 1 : 0;
                    const uint32_t idesc = pr ? idesc_p[1] : idesc_p[0];
                    const uint32_t a_step = pr ? a_step_p[1] : a_step_p[0], b_step = pr ? b_step_p[1] : b_step_p[0];
                    const uint32_t a_hi = pr ? a_hi_p[1] : a_hi_p[0], b_hi = pr ? b_hi_p[1] : b_hi_p[0];
                    const uint32_t aaddr = smem_base + (uint32_t)stage * (uint32_t)L.stage_bytes;
                    // operand addresses are < 256 KiB: the 14-bit start-address field (16-byte units) never carries
                    const uint32_t a_lo = (pr ? a_lo_p[1] : a_lo_p[0]) | ((aaddr & 0x3FFFFu) >> 4);
                    const uint32_t b_lo = (pr ? b_lo_p[1] : b_lo_p[0]) | (((aaddr + A_BYTES) & 0x3FFFFu) >> 4);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        if (ci == u.c_begin) DBG_UNIT(2, ord);
                        if (ci - u.c_begin < 24 && (L.n_phases == 0 || ord == 1)) DBG_MARK(32 + (ci - u.c_begin));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if constexpr (pair) umma_bf16_2sm(d_tmem, a_lo + k * a_step, a_hi, b_lo + k * b_step, b_hi, idesc, accumulate);
                            else umma_bf16(d_tmem, a_lo + k * a_step, a_hi, b_lo + k * b_step, b_hi, idesc, accumulate);
                            accumulate = 1;
                        }
                        // the smem slot is free once these MMAs retire (both CTAs' producers are told)
                        if constexpr (pair) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);
                    }
                    accumulate = 1;
                    if (++stage == L.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) {
                    if constexpr (pair) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);   // accumulator complete -> epilogue
                    DBG_MARK(2); DBG_UNIT(3, ord);
                }
                __syncwarp();
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == W_PUB0 || warp == W_PUB1) {
        // ================================ publishers ========================================
        // Send every staged 64-column granule of the bf16 outputs to global memory with a TMA store
        // and, once that store has completed (bulk async-group wait: no MEMBAR over the SM's whole store
        // stream -- a gpu-scope fence issued while epilogue warps stream stores only returns when they
        // pause), bump the granule's dataflow counter with a release: consumers start on a row block's
        // first columns while its last ones are still in the epilogue.  Two publisher threads take the
        // even / odd granules so that one store's completion latency hides behind the other's.
        // fp32 outputs (dW partials) are stored by the epilogue warps directly and published per unit.
        if (lane == 0) {
            const int pub = warp - W_PUB0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t seq = 0;                               // FIFO position of the unit's first granule
            int pi = 0, ul = -1;
            int ord = -1;
            const uint32_t out_base = smem_u32(out_stage);
            while (walk_next(pi, ul, cid, n_ops)) {
                ++ord;
                const TcPhaseLite* ph = &c_ph[pi];
                const TcPhase* gp = &gph[pi];
                const UnitInfo u = decode_unit(ph, ul);
                const bool has_mean = ph->out_mean_bf != nullptr, has_state = ph->out_state_bf != nullptr;
                const int n_arr = (has_mean ? 1 : 0) + (has_state ? 1 : 0);
                const int gpt = n_arr ? ph->gran_per_tile : 0;
                const int row0 = (u.m_group * CL + crank) * BM;
                int* const ctr = ph->chunk_ctr ? ph->chunk_ctr + ((size_t)u.m_group * ph->n_tiles + u.n_blk) * gpt : nullptr;
                for (int g = pub; g < gpt; g += 2) {
                    const int col0 = u.n_blk * ph->BN + g * GRAN_COLS;
                    uint32_t sq = seq + (uint32_t)(g * n_arr);
                    int bufs[2] = {-1, -1};
                    for (int a = 0; a < 2; ++a) {
                        if (!(a == 0 ? has_mean : has_state)) continue;
                        const int b = (int)(sq % N_OUT_BUF);
                        mbar_wait(&out_full[b], (sq / N_OUT_BUF) & 1u);
                        if (col0 < ph->N) tma_store_2d((uint64_t)&gp->tmOut[a], out_base + (uint32_t)b * OUT_BUF_BYTES, col0, row0);
                        bufs[a] = b; ++sq;
                    }
                    bulk_commit();
                    if (L.dbg && blockIdx.x == 0 && ord < 24 && g < 4) L.dbg[256 + ord * 8 + 4 + g] = (unsigned long long)clock64();
                    bulk_wait<0>();                          // stores complete: data is in global memory
                    for (int a = 0; a < 2; ++a) if (bufs[a] >= 0) mbar_arrive(&out_free[bufs[a]]);
                    if (ctr && col0 < ph->N) {
                        red_release_add(ctr + g, 1);
                        if (L.dbg && L.n_phases && ord == 0 && g == 0) L.dbg[512 + blockIdx.x] = gtime();
                        if (L.dbg && L.n_phases && ord == 0 && g == 3) L.dbg[704 + blockIdx.x] = gtime();
                        if (L.dbg && blockIdx.x == 0 && ord < 24 && g < 4) L.dbg[256 + ord * 8 + g] = (unsigned long long)clock64();
                    }
                }
                seq += (uint32_t)(gpt * n_arr);
                mbar_wait(&unit_bar[acc], acc_phase);      // every epilogue warp has finished the unit (direct fp32 stores included)
                if (ph->done_ctr) red_release_add(ph->done_ctr + u.m_group, 1);
                if (pub == 1) DBG_UNIT(6, ord);
                // the barriers of this accumulator stage may be reused: part of the stage's release
                if (pair && crank == 1) { if constexpr (pair) mbar_arrive_remote(&tempty[acc], 0); } else mbar_arrive(&tempty[acc]);
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue (16 warps) ================================
        // warp w may read TMEM lanes 32*(w%4)..+31; the four warps of a lane quarter interleave over
        // the tile's 16-column chunks.
        const int ew = warp;
        const int quarter = ew & 3;
        EpiCtx c;
        c.sub = ew >> 2; c.lane = lane;
        c.remote_arrive = pair && crank == 1;
        uint32_t out_seq = 0;
        c.seq = &out_seq;
        c.out_base = smem_u32(out_stage); c.out_full = out_full; c.out_free = out_free;
        c.mant_mask = L.mant_mask; c.one_bits = L.one_bits;
        c.row_off = (uint32_t)(quarter * 32 + lane) * 128u; c.row_swz = (uint32_t)(lane & 7);
        int acc = 0; uint32_t acc_phase = 0;
        int pi = 0, ul = -1;
        int ord = -1;
        const int et = threadIdx.x;                // 0..511: the epilogue threads
        while (walk_next(pi, ul, cid, n_ops)) {
            ++ord;
            const TcPhaseLite* ph = &c_ph[pi];
            const UnitInfo u = decode_unit(ph, ul);
            {   // this tile's (pre-scaled) bias slice -> shared memory, while the MMAs are still running
                const float bsc = (ph->act == ACT_SIGMOID) ? ph->bias_scale * -1.4426950408889634f : ph->bias_scale;
                const int n = u.n_blk * ph->BN + et;
                if (et < 256) s_bias[acc * 256 + et] = (ph->bias && et < ph->BN && n < ph->N) ? bsc * __ldg(ph->bias + n) : 0.f;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                c.sbias = s_bias + acc * 256;
            }
            c.p.M = ph->M; c.p.N = ph->N; c.p.BN = ph->BN; c.p.act = ph->act; c.p.sample = ph->sample;
            c.p.acc_scale = ph->acc_scale; c.p.bias_scale = ph->bias_scale;
            c.p.bias = ph->bias; c.p.sigma = ph->sigma; c.p.noise_sigma = ph->noise_sigma;
            c.p.out_mean_bf = ph->out_mean_bf; c.p.ld_mean_bf = ph->ld_mean_bf;
            c.p.out_state_bf = ph->out_state_bf; c.p.ld_state_bf = ph->ld_state_bf;
            c.p.out_f32 = ph->out_f32; c.p.ld_f32 = ph->ld_f32; c.p.split_stride = ph->split_stride;
            c.p.ais_a = ph->ais_a; c.p.ais_b = ph->ais_b; c.p.ais_next = ph->ais_next; c.p.ais_lin = ph->ais_lin; c.p.ais_logw = ph->ais_logw;
            c.p.ais_ct = ph->ais_b - ph->ais_a; c.p.ais_cd = ph->ais_next - 0.5f * (ph->ais_a + ph->ais_b);
            const int ph_mode = ph->mode;
            c.rng.k0 = L.k0; c.rng.k1 = L.k1; c.rng.tick = L.tick + ph->tick_off; c.rng.row0 = L.row0; c.rng.c2 = ph->rng_c2;
            c.m = (u.m_group * CL + crank) * BM + quarter * 32 + lane;
            c.n_blk = u.n_blk; c.split = u.split;
            c.tempty = &tempty[acc];
            mbar_wait_relaxed(&tfull[acc], acc_phase, (uint32_t)L.epi_ns);
            tc_fence_after();
            if (threadIdx.x == 0) { DBG_MARK(3); DBG_UNIT(4, ord); }
            c.t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            switch (ph_mode) {
                case MODE_SIG_BERN_MEAN_STATE: epilogue_tile<MODE_SIG_BERN_MEAN_STATE, pair>(c); break;
                case MODE_SIG_BERN_STATE: epilogue_tile<MODE_SIG_BERN_STATE, pair>(c); break;
                case MODE_SIG_MEAN: epilogue_tile<MODE_SIG_MEAN, pair>(c); break;
                case MODE_RAW_F32: epilogue_tile<MODE_RAW_F32, pair>(c); break;
                case MODE_AIS_UNITS: epilogue_tile<MODE_AIS_UNITS, pair>(c); break;
                case MODE_AIS_UNITS_S: epilogue_tile<MODE_AIS_UNITS_S, pair>(c); break;
                case MODE_AIS_UNITS_G: epilogue_tile<MODE_AIS_UNITS_G, pair>(c); break;
                case MODE_AIS_STATE: epilogue_tile<MODE_AIS_STATE, pair>(c); break;
                default: epilogue_tile<MODE_GENERIC, pair>(c); break;
            }
            if (threadIdx.x == 0) DBG_UNIT(5, ord);
            // this warp's part of the unit is stored: the publisher warp makes it visible to other SMs
            __syncwarp();
            if (lane == 0) mbar_arrive(&unit_bar[acc]);
            if (threadIdx.x == 0) DBG_MARK(4);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    // no CTA may exit while its peer can still signal its barriers or read its shared memory
    if constexpr (pair) cluster_sync_all(); else __syncthreads();
    if (warp == W_PUB0) {
        tc_fence_after();
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    if (threadIdx.x == 0) { DBG_MARK(5); if (L.dbg && blockIdx.x == 0) L.dbg[7] = gtime(); }
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

struct TilePick { int bn, cluster; };

static TilePick pick_tile(int N, bool b_mn, bool staged_out, int m_tiles, int splits, int chunks, int sms, int force_cluster) {
    // Tile width BN and CTA grouping.  A single SM ingests ~50 B/clk from L2 (measured: the ring of a
    // 128x256 tile refills at 52 B/clk with one CTA alone on the chip), while tcgen05 at M=128
    // consumes 8192*(1/128 + 1/BN) B/clk of operands: a lone CTA is ingest-bound.  A CTA pair
    // (cta_group::2, M=256) halves the B bytes each SM needs.
    // Cycle model per CTA: waves * chunks * max(mma, ingest) + epilogue of the last tile.
    TilePick best{b_mn ? 64 : 16, 1};
    double best_cost = 1e30;
    for (int c = 1; c <= 2; ++c) {
        if (force_cluster && c != force_cluster) continue;
        if (!force_cluster && c == 2 && m_tiles < 2) break;
        // every CTA of a pair holds BN/2 columns of B (whole boxes / 8-row atoms); bf16 outputs leave in
        // 64-column TMA-store granules, so their tiles are whole granules wide (the tensor edge clips)
        const int step = b_mn ? 64 * c : (staged_out ? 64 : 16);
        const int slots = sms / c;
        for (int bn = step; bn <= 256; bn += step) {
            const int nt = (N + bn - 1) / bn;
            const int m_groups = (m_tiles + c - 1) / c;
            const long units = (long)m_groups * nt * splits;
            const long waves = (units + slots - 1) / slots;
            const double mma = 4.0 * (bn / 2.0);                                  // cycles per K chunk (4 x K=16)
            const double ingest = (16384.0 + (bn / c) * 128.0) / 45.0;           // bytes per chunk per SM / (B/clk)
            const double cost = (double)waves * chunks * (mma > ingest ? mma : ingest) + 12.0 * bn;
            if (cost < best_cost - 1e-9) { best_cost = cost; best.bn = bn; best.cluster = c; }
        }
    }
    return best;
}

static int epilogue_mode(const TcGemm& g) {
    if (g.ais_kind == 1) {
        const bool emit = g.out_state_bf != nullptr && g.sample == SMP_BERNOULLI;
        const float ct = fabsf(g.ais_b - g.ais_a), cd = fabsf(g.ais_next - 0.5f * (g.ais_a + g.ais_b));
        const bool close = ct <= AIS_MAX_CT && cd <= AIS_MAX_CD;
        if (emit && !g.ais_logw) return MODE_AIS_UNITS_S;
        if (emit && g.ais_logw && close) return MODE_AIS_UNITS;
        return MODE_AIS_UNITS_G;
    }
    if (g.ais_kind == 2) return MODE_AIS_STATE;
    if (g.sigma || g.noise_sigma) return MODE_GENERIC;
    const bool mb = g.out_mean_bf != nullptr, sb = g.out_state_bf != nullptr, f = g.out_f32 != nullptr;
    if (g.act == ACT_SIGMOID && g.sample == SMP_BERNOULLI && mb && sb && !f) return MODE_SIG_BERN_MEAN_STATE;
    if (g.act == ACT_SIGMOID && g.sample == SMP_BERNOULLI && !mb && sb && !f) return MODE_SIG_BERN_STATE;
    if (g.act == ACT_SIGMOID && g.sample == SMP_NONE && mb && !sb && !f) return MODE_SIG_MEAN;
    if (g.act == ACT_LINEAR && g.sample == SMP_NONE && !mb && !sb && f) return MODE_RAW_F32;
    return MODE_GENERIC;
}

// fills everything of a phase descriptor except the dataflow fields and the unit range
static void fill_phase(Ctx* ctx, const TcGemm& g, int cluster, TcPhase& ph) {
    TcPhaseLite& p = ph.l;
    BM_REQUIRE(g.M > 0 && g.N > 0 && g.n_pairs >= 1 && g.n_pairs <= 2, "bad tensor-core GEMM shape");
    memset(&ph, 0, sizeof(ph));
    p.M = g.M; p.N = g.N; p.n_pairs = g.n_pairs;
    const int m_tiles = (g.M + BM - 1) / BM;
    bool need64 = false;
    for (int i = 0; i < g.n_pairs; ++i) need64 = need64 || g.b_t[i];
    int chunks_total = 0;
    for (int i = 0; i < g.n_pairs; ++i) chunks_total += (g.K[i] + BK - 1) / BK;
    const int nsplit = g.splits > 0 ? g.splits : 1;
    const bool staged_out = g.out_mean_bf != nullptr || g.out_state_bf != nullptr;
    TilePick tp = pick_tile(g.N, need64, staged_out, m_tiles, nsplit, (chunks_total + nsplit - 1) / nsplit, ctx->sm_count, cluster);
    if (g.force_bn > 0) tp.bn = g.force_bn;
    BM_REQUIRE(tp.cluster == cluster, "tile picker returned another cluster size");
    BM_REQUIRE(cluster == 1 || (need64 ? tp.bn % 128 == 0 : tp.bn % 16 == 0), "CTA-pair tiles need BN/2 on whole boxes / swizzle atoms");
    BM_REQUIRE(tp.bn % 16 == 0 && tp.bn <= 256 && (!need64 || tp.bn % 64 == 0), "bad tile width");
    BM_REQUIRE(!staged_out || tp.bn % 64 == 0, "ops with bf16 outputs need tiles of whole 64-column granules");
    p.BN = tp.bn;
    p.gran_per_tile = (p.BN + GRAN_COLS - 1) / GRAN_COLS;
    p.m_groups = (m_tiles + cluster - 1) / cluster;
    p.n_tiles = (g.N + p.BN - 1) / p.BN;
    p.splits = nsplit;
    p.split_stride = g.split_stride;
    for (int i = 0; i < 2; ++i) {
        const int j = i < g.n_pairs ? i : 0;
        p.chunks[i] = i < g.n_pairs ? (g.K[j] + BK - 1) / BK : 0;
        p.a_mn[i] = g.a_t[j]; p.b_mn[i] = g.b_t[j]; p.a_neg[i] = g.neg[j];
        p.a_row0[i] = g.a_row0[j]; p.a_k0[i] = g.a_k0[j]; p.a_batch[i] = g.a_batch[j];
        // K-major: box = 64 k x (128 | BN/cluster) rows; MN-major: box = 64 mn x 64 k
        ph.tmA[i] = g.a_t[j] ? make_map(g.A[j], 64, 64) : make_map(g.A[j], 64, BM);
        ph.tmB[i] = g.b_t[j] ? make_map(g.B[j], 64, 64) : make_map(g.B[j], 64, p.BN / cluster);
        BM_REQUIRE(i >= g.n_pairs || g.K[j] > 0, "tensor-core GEMM pair with K == 0");
    }
    BM_REQUIRE(p.splits <= chunks_total, "more K splits than K chunks");
    BM_REQUIRE(p.splits == 1 || (g.out_f32 && !g.out_mean_bf && !g.out_state_bf), "split-K writes fp32 partials only");
    p.acc_scale = g.acc_scale; p.bias_scale = g.bias_scale;
    p.bias = g.bias; p.sigma = g.sigma; p.noise_sigma = g.noise_sigma;
    p.act = g.act; p.sample = g.sample; p.mode = epilogue_mode(g);
    p.rng_c2 = g.rng.c2;
    p.ais_kind = g.ais_kind; p.ais_a = g.ais_a; p.ais_b = g.ais_b; p.ais_next = g.ais_next; p.ais_lin = g.ais_lin;
    p.ais_logw = g.ais_logw; p.tick_off = g.tick_off;
    if (g.ais_kind == 1)
        BM_REQUIRE(!g.sigma && !g.noise_sigma && g.act == ACT_LINEAR && !g.out_mean_bf && !g.out_f32 && g.splits <= 1 &&
                   g.acc_scale == 1.f && g.bias_scale == 1.f && (g.sample == SMP_NONE || g.sample == SMP_BERNOULLI),
                   "AIS unit op: linear accumulator, optional state output only");
    if (g.ais_kind == 2)
        BM_REQUIRE(!g.sigma && !g.noise_sigma && g.act == ACT_SIGMOID && g.sample == SMP_BERNOULLI && g.out_state_bf && !g.out_mean_bf &&
                   !g.out_f32 && g.splits <= 1 && g.bias, "AIS state op: sampled sigmoid states with a bias");
    p.out_mean_bf = g.out_mean_bf; p.ld_mean_bf = g.ld_mean_bf;
    p.out_state_bf = g.out_state_bf; p.ld_state_bf = g.ld_state_bf;
    p.out_f32 = g.out_f32; p.ld_f32 = g.ld_f32;
    // bf16 outputs leave the SM as TMA stores of (64 columns x 128 rows) granules, clipped to [M, N]
    if (g.out_mean_bf) { TcMat o; o.ptr = g.out_mean_bf; o.rows = g.M; o.cols = g.N; o.ld = g.ld_mean_bf; ph.tmOut[0] = make_map(o, GRAN_COLS, BM); }
    if (g.out_state_bf) { TcMat o; o.ptr = g.out_state_bf; o.rows = g.M; o.cols = g.N; o.ld = g.ld_state_bf; ph.tmOut[1] = make_map(o, GRAN_COLS, BM); }
    BM_REQUIRE(!g.out_mean_bf || (g.ld_mean_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
    BM_REQUIRE(!g.out_state_bf || (g.ld_state_bf % 8 == 0), "bf16 output leading dimension must be a multiple of 8");
}

// The ops' descriptors (minus tensor maps) go to constant memory, stream-ordered before the launch;
// skipped when the bank already holds exactly this image (the common case: the same program every step).
static void upload_ops(Ctx* ctx, const TcPhase* ph, int n) {
    struct Bank { std::vector<unsigned char> image; cudaStream_t user = nullptr; };
    static std::map<int, Bank> banks;                             // constant memory is per device
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    Bank& bank = banks[ctx->device];
    // another context of this device (another stream) may still be running a kernel that reads the bank: wait for it
    // before the bank changes hands
    if (bank.user && bank.user != ctx->stream) BM_CUDA(cudaStreamSynchronize(bank.user));
    bank.user = ctx->stream;
    std::vector<unsigned char>& current = bank.image;
    std::vector<unsigned char> img((size_t)n * sizeof(TcPhaseLite));
    for (int i = 0; i < n; ++i) memcpy(img.data() + (size_t)i * sizeof(TcPhaseLite), &ph[i].l, sizeof(TcPhaseLite));
    if (img.size() <= current.size() && memcmp(img.data(), current.data(), img.size()) == 0) return;
    BM_CUDA(cudaMemcpyToSymbolAsync(c_ph, img.data(), img.size(), 0, cudaMemcpyHostToDevice, ctx->stream));
    if (current.size() < img.size()) current.resize(img.size());
    memcpy(current.data(), img.data(), img.size());
}

static void do_launch(Ctx* ctx, TcLaunch& L, int cluster, double flops, int max_bn) {
    static bool attr_set[64] = {false};                         // function attributes are per device
    if (!attr_set[ctx->device & 63]) {
        BM_CUDA(cudaFuncSetAttribute(tc_program_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BM_CUDA(cudaFuncSetAttribute(tc_program_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[ctx->device & 63] = true;
    }
    // ring stages sized for the widest tile of the launch: narrower tiles buy a deeper pipeline
    L.stage_bytes = A_BYTES + (max_bn / cluster) * BK * 2;
    L.stage_bytes = (L.stage_bytes + 1023) & ~1023;           // 128B-swizzle atoms are 1024-byte aligned
    L.stages = RING_BYTES / L.stage_bytes;
    if (L.stages > MAX_STAGES) L.stages = MAX_STAGES;
    { const char* e = getenv("BM_TC_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= L.stages) L.stages = atoi(e); }
    const int max_clusters = ctx->sm_count / cluster;
    // programs address pairs by number (TcPhaseLite::pair_begin/pair_count): launch all of them
    const int n_clusters = L.n_phases ? max_clusters : (L.total_units < max_clusters ? L.total_units : max_clusters);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(n_clusters * cluster); lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = SMEM_BYTES; lc.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (L.n_phases) {
        // a program's CTAs wait for each other through counters in global memory: all of them must be resident at once.  Ask the
        // runtime once per device (a GPU shared through MPS / MIG, or a part with fewer SMs than it reports, would otherwise only
        // show up as the kernel's 4-second trap); BM_TC_SKIP_RESIDENCY_CHECK=1 trusts the device properties instead.
        static int resident[64] = {0};
        int& r = resident[ctx->device & 63];
        if (r == 0) {
            const char* e = getenv("BM_TC_SKIP_RESIDENCY_CHECK");
            int n = 0;
            if (e && atoi(e) == 1) r = -1;
            else if (cudaOccupancyMaxActiveClusters(&n, cluster == 2 ? (const void*)tc_program_kernel<2> : (const void*)tc_program_kernel<1>, &lc) == cudaSuccess) r = n > 0 ? n : -2;
            else { cudaGetLastError(); r = -1; }
            if (getenv("BM_TC_DEBUG_RESIDENCY"))
                fprintf(stderr, "[bm] device %d: cudaOccupancyMaxActiveClusters(tc_program_kernel<%d>) = %d, the program grid has %d clusters\n",
                        ctx->device, cluster, n, n_clusters);
        }
        BM_REQUIRE(r == -1 || r >= n_clusters,
                   "the device cannot hold all CTA clusters of a dataflow program at once (cudaOccupancyMaxActiveClusters); it would deadlock");
    }
    if (ctx->profile_tc) BM_CUDA(cudaEventRecord(profile_event(ctx), ctx->stream));
    // kernels that contain cta_group::2 instructions must be launched as clusters of 2: separate instantiations
    if (cluster == 2) BM_CUDA(cudaLaunchKernelEx(&lc, tc_program_kernel<2>, L));
    else BM_CUDA(cudaLaunchKernelEx(&lc, tc_program_kernel<1>, L));
    if (ctx->profile_tc) {
        BM_CUDA(cudaEventRecord(profile_event(ctx), ctx->stream));
        ctx->prof_flops += flops;       // algorithmic FLOPs (no padding counted)
        ctx->prof_launches++;
    }
    count_launch(ctx);
}

static double gemm_flops(const TcGemm& g) {
    double k_total = 0.0;
    for (int i = 0; i < g.n_pairs; ++i) k_total += g.K[i];
    return 2.0 * g.M * g.N * k_total;
}

void launch_tc_gemm(Ctx* ctx, const TcGemm& g) {
    const int m_tiles = (g.M + BM - 1) / BM;
    int cluster = g.force_cluster > 0 ? g.force_cluster : (m_tiles >= 2 ? 2 : 1);
    BM_REQUIRE(cluster == 1 || cluster == 2, "cluster must be 1 or 2");
    if (cluster == 2) {      // fall back to single CTAs when the B tile cannot be halved on box boundaries
        bool need64 = false;
        for (int i = 0; i < g.n_pairs; ++i) need64 = need64 || g.b_t[i];
        if (need64 && g.N < 128 && !g.force_cluster) cluster = 1;
    }
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    fill_phase(ctx, g, cluster, L.inl);
    L.inl.l.n_units = L.inl.l.m_groups * L.inl.l.n_tiles * L.inl.l.splits;
    L.total_units = L.inl.l.n_units;
    {   // every CTA pair (or CTA) of the grid takes part
        const int max_clusters = ctx->sm_count / cluster;
        L.inl.l.pair_begin = 0;
        L.inl.l.pair_count = L.total_units < max_clusters ? L.total_units : max_clusters;
    }
    L.k0 = g.rng.k0; L.k1 = g.rng.k1; L.tick = g.rng.tick; L.row0 = g.rng.row0;
    L.dbg = g.dbg;
    L.mant_mask = 0x007FFFFFu; L.one_bits = 0x3F800000u;
    L.epoch = 1; L.poll_ns = 64; L.epi_ns = 128;
    upload_ops(ctx, &L.inl, 1);
    do_launch(ctx, L, cluster, gemm_flops(g), L.inl.l.BN);
}

int tc_plan_units(Ctx* ctx, const TcGemm& g) {
    TcPhase ph;
    fill_phase(ctx, g, 2, ph);
    return ph.l.m_groups * ph.l.n_tiles * ph.l.splits;
}

TcProgram::~TcProgram() {
    if (dev_phases) cudaFree(dev_phases);
    if (dev_counters) cudaFree(dev_counters);
}

void launch_tc_program(Ctx* ctx, TcProgram& prog, RngKey rng, int batch_row) {
    const int n = (int)prog.ops.size();
    BM_REQUIRE(n >= 1 && n <= MAX_PHASES, "a program holds 1..96 ops");
    const int cluster = 2;            // programs always run on CTA pairs (all CTAs walk one unit list)
    std::vector<unsigned char> image((size_t)n * sizeof(TcPhase));
    TcPhase* ph = reinterpret_cast<TcPhase*>(image.data());
    // counters: per op one int per row-block group (unit level) + one per (row-block group, tile, granule)
    size_t n_ctr = 0;
    std::vector<size_t> ctr_off(n), cctr_off(n);
    int unit = 0;
    double flops = 0.0;
    for (int i = 0; i < n; ++i) {
        fill_phase(ctx, prog.ops[i], cluster, ph[i]);
        ph[i].l.n_units = ph[i].l.m_groups * ph[i].l.n_tiles * ph[i].l.splits;
        unit += ph[i].l.n_units;
        ctr_off[i] = n_ctr;
        n_ctr += (size_t)ph[i].l.m_groups;
        cctr_off[i] = n_ctr;
        n_ctr += (size_t)ph[i].l.m_groups * ph[i].l.n_tiles * ph[i].l.gran_per_tile;
        flops += gemm_flops(prog.ops[i]);
    }
    if (n_ctr > prog.n_counters) {
        BM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (prog.dev_counters) cudaFree(prog.dev_counters);
        BM_CUDA(cudaMalloc(&prog.dev_counters, n_ctr * sizeof(int)));
        prog.n_counters = n_ctr;
        prog.host_image.clear();
        prog.epoch = 0;
    }
    // pair assignment: chain ops on pairs [0, P), spare-lane ops on [P, total), the rest on all pairs
    {
        const int total = ctx->sm_count / cluster;
        int P = 0;
        for (int i = 0; i < n; ++i) if (prog.ops[i].lane == LANE_CHAIN) P = std::max(P, ph[i].l.n_units);
        if (P > total || P == 0) P = total;
        for (int i = 0; i < n; ++i) {
            int lane = prog.ops[i].lane;
            if (lane == LANE_SPARE && total - P < 1) lane = LANE_ALL;
            ph[i].l.pair_begin = lane == LANE_SPARE ? P : 0;
            ph[i].l.pair_count = lane == LANE_CHAIN ? P : (lane == LANE_SPARE ? total - P : total);
        }
    }
    static int chunk_deps = -1;
    if (chunk_deps < 0) { const char* e = getenv("BM_TC_CHUNK_DEPS"); chunk_deps = e ? atoi(e) : 1; }
    for (int i = 0; i < n; ++i) {
        const TcGemm& g = prog.ops[i];
        ph[i].l.done_ctr = prog.dev_counters + ctr_off[i];
        ph[i].l.n_deps = g.n_deps;
        for (int d = 0; d < g.n_deps; ++d) {
            const int j = g.dep[d];
            BM_REQUIRE(j >= 0 && j < i, "a program op may only depend on earlier ops");
            ph[i].l.dep_ctr[d] = prog.dev_counters + ctr_off[j];
            ph[i].l.dep_need[d] = ph[j].l.n_tiles * ph[j].l.splits * cluster * 2;  // both publishers of every CTA, per unit
            if (g.dep_all[d]) ph[i].l.dep_groups[d] = ph[j].l.m_groups;
            else {
                ph[i].l.dep_groups[d] = 0;
                BM_REQUIRE(ph[i].l.m_groups <= ph[j].l.m_groups, "row-block dependency on an op with fewer row blocks");
                // granule-level dataflow: this op's (K-major) A operand is op j's output (K = its N); only
                // for the first such dependency, single-pair ops without split-K
                const TcPhaseLite& pj = ph[j].l;
                const int kchunks = ph[i].l.chunks[0];
                // (the kernel's wait loop skips the unit-level wait of EVERY same-row-block dependency of an op that follows
                // granules: only an op whose single such dependency is this one may do so)
                int same_row_deps = 0;
                for (int q = 0; q < g.n_deps; ++q) same_row_deps += g.dep_all[q] ? 0 : 1;
                if (chunk_deps && d == 0 && same_row_deps == 1 && g.n_pairs == 1 && !g.a_t[0] && ph[i].l.splits == 1 && pj.splits == 1 &&
                    kchunks <= MAX_KCHUNKS && g.K[0] == pj.N && pj.n_tiles * pj.gran_per_tile <= 255) {
                    ph[j].l.chunk_ctr = prog.dev_counters + cctr_off[j];
                    ph[i].l.dep_chunk_ctr = prog.dev_counters + cctr_off[j];
                    ph[i].l.dep_gran_row = pj.n_tiles * pj.gran_per_tile;
                    ph[i].l.dep_chunk_need = cluster;              // the publisher of each CTA of the producing pair
                    // K chunk c = columns [64c, 64c+64) of the producer's output: the granule(s) holding its
                    // first and last column; consumed in the order the producer's epilogues finish them
                    std::vector<std::pair<int, int>> order;
                    for (int c = 0; c < kchunks; ++c) {
                        const int col_a = 64 * c, col_b = std::min(64 * c + 63, pj.N - 1);
                        const int ta = col_a / pj.BN, tb = col_b / pj.BN;
                        const int ga = (col_a - ta * pj.BN) / GRAN_COLS, gb = (col_b - tb * pj.BN) / GRAN_COLS;
                        ph[i].l.k_dep_a[c] = (unsigned char)(ta * pj.gran_per_tile + ga);
                        ph[i].l.k_dep_b[c] = (unsigned char)(tb * pj.gran_per_tile + gb);
                        order.push_back(std::make_pair(std::max(ga, gb), c));
                    }
                    std::stable_sort(order.begin(), order.end());
                    for (int c = 0; c < kchunks; ++c) ph[i].l.k_order[c] = (unsigned char)order[c].second;
                }
            }
        }
    }
    // descriptors live in device memory; re-uploaded only when they changed (buffers and shapes are
    // stable across steps; the per-step values -- tick, batch cursor -- travel in the kernel parameters)
    if (image.size() != prog.host_image.size() || memcmp(image.data(), prog.host_image.data(), image.size()) != 0) {
        if (image.size() > prog.dev_phases_bytes) {
            BM_CUDA(cudaStreamSynchronize(ctx->stream));   // a running launch may still read the old allocation
            if (prog.dev_phases) cudaFree(prog.dev_phases);
            BM_CUDA(cudaMalloc(&prog.dev_phases, image.size()));
            prog.dev_phases_bytes = image.size();
        }
        // stream-ordered: the copy runs after every earlier launch of this stream has finished reading the descriptors (the
        // pageable source is staged before the call returns) -- no host synchronisation between the launches of a ladder
        BM_CUDA(cudaMemcpyAsync(prog.dev_phases, image.data(), image.size(), cudaMemcpyHostToDevice, ctx->stream));
        prog.host_image = image;
        prog.epoch = 0;                                   // another op list: other arrival counts per launch
    }
    if (prog.epoch == 0 || prog.epoch >= (1 << 20)) {
        BM_CUDA(cudaMemsetAsync(prog.dev_counters, 0, prog.n_counters * sizeof(int), ctx->stream));
        prog.epoch = 0;
    }
    // (the epoch is only advanced once the launch has been accepted: after a refused launch the counters of the
    //  program are re-zeroed before its next launch instead of being waited on one epoch behind)
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.phases = reinterpret_cast<const TcPhase*>(prog.dev_phases);
    L.n_phases = n;
    L.total_units = unit;
    L.k0 = rng.k0; L.k1 = rng.k1; L.tick = rng.tick; L.row0 = rng.row0;
    L.batch_row = batch_row;
    L.epoch = prog.epoch + 1;
    { static int pn = -1, en = -1;
      if (pn < 0) { const char* e = getenv("BM_TC_POLL_NS"); pn = e ? atoi(e) : 64; }
      if (en < 0) { const char* e = getenv("BM_TC_EPI_NS"); en = e ? atoi(e) : 128; }
      L.poll_ns = pn; L.epi_ns = en; }
    { static int fl = -1; if (fl < 0) { const char* e = getenv("BM_TC_FLAGS"); fl = e ? atoi(e) : 3; } L.flags = fl; }
    L.mant_mask = 0x007FFFFFu; L.one_bits = 0x3F800000u;
    static unsigned long long* dbg_buf = nullptr;
    static int dbg_left = -1;
    if (dbg_left < 0) { const char* e = getenv("BM_TC_PROGRAM_TIMELINE"); dbg_left = e ? atoi(e) : 0; }
    const bool dbg = dbg_left > 0;
    if (dbg) {
        if (!dbg_buf) BM_CUDA(cudaMalloc(&dbg_buf, 1024 * sizeof(unsigned long long)));
        BM_CUDA(cudaMemsetAsync(dbg_buf, 0, 1024 * sizeof(unsigned long long), ctx->stream));
        L.dbg = dbg_buf;
    }
    int max_bn = 16;
    for (int i = 0; i < n; ++i) max_bn = std::max(max_bn, ph[i].l.BN);
    try {
        upload_ops(ctx, ph, n);
        do_launch(ctx, L, cluster, flops, max_bn);
        ++prog.epoch;
    } catch (...) {
        prog.epoch = 0;
        throw;
    }
    if (dbg) {
        --dbg_left;
        unsigned long long h[1024];
        BM_CUDA(cudaStreamSynchronize(ctx->stream));
        BM_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "program timeline (SM cycles, CTA 0): setup=%llu exit=%llu wall=%llu ns, %d units total\n",
                h[1] - h[0], h[5] - h[0], h[7] - h[6], unit);
        fprintf(stderr, "  unit 1 tma_issue:");
        for (int i = 0; i < 24 && h[8 + i]; ++i) fprintf(stderr, " %llu", h[8 + i] - h[0]);
        fprintf(stderr, "\n  first-unit granule 0 / granule 3 publish time per CTA (ns after CTA 0 start):");
        for (int i = 0; i < 148; ++i) if (h[512 + i]) fprintf(stderr, " %d:%lld/%lld", i, (long long)(h[512 + i] - h[6]), (long long)(h[704 + i] - h[6]));
        fprintf(stderr, "\n  unit 1 poll ok (time:upto):");
        for (int i = 0; i < 20; ++i) if (h[456 + i]) fprintf(stderr, " [%d] %llu:%llu", i, h[456 + i] - h[0], h[480 + i]);
        fprintf(stderr, "\n  unit 1 fenced:");
        for (int i = 0; i < 20; ++i) if (h[432 + i]) fprintf(stderr, " [%d] %llu", i, h[432 + i] - h[0]);
        fprintf(stderr, "\n  unit 1 mma_ready:");
        for (int i = 0; i < 24 && h[32 + i]; ++i) fprintf(stderr, " %llu", h[32 + i] - h[0]);
        fprintf(stderr, "\n  ord: start dep_ok mma_first mma_done epi_start epi_done published\n");
        for (int o = 0; o < 24 && h[64 + o * 8]; ++o) {
            fprintf(stderr, "  %2d:", o);
            for (int k = 0; k < 7; ++k) fprintf(stderr, " %7llu", h[64 + o * 8 + k] ? h[64 + o * 8 + k] - h[0] : 0ull);
            fprintf(stderr, "   | granule staged:");
            for (int k = 4; k < 8; ++k) fprintf(stderr, " %7llu", h[256 + o * 8 + k] ? h[256 + o * 8 + k] - h[0] : 0ull);
            fprintf(stderr, " published:");
            for (int k = 0; k < 4; ++k) fprintf(stderr, " %7llu", h[256 + o * 8 + k] ? h[256 + o * 8 + k] - h[0] : 0ull);
            fprintf(stderr, "\n");
        }
    }
}

}  // namespace bm
