// Micro-test for the round-2 plan (DESIGN.md 5.1 "cluster-resident chain"): do the primitives the plan combines work
// TOGETHER -- four tcgen05 cta_group::2 CTA pairs (ranks 2j, 2j+1) inside ONE 8-CTA cluster, each pair with its own TMEM
// allocation, its own 256 x 128 x 64 MMA and a commit multicast to exactly its two CTAs (mask 3 << 2j), while tiles are
// pushed between CTAs of DIFFERENT pairs with cp.async.bulk.shared::cluster (completion on the destination's mbarrier)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cluster8_pair_mma tools/ubench/cluster8_pair_mma.cu
//   timeout 60 ./cluster8_pair_mma
//
// Operand tiles are all ones (bf16 1.0): whatever the swizzle / descriptor layout, the matrices are all ones, so every
// accumulator element must be exactly K.  The A tile of the SECOND MMA round is not written locally: it arrives by a
// DSMEM push from the CTA of the same parity in the NEXT pair (rank + 2) -- the hand-over the resident chain needs --
// holding the value 2.0, so round 2 must add 2 * K on top.  Expected accumulator: 64 after round 1, 64 + 128 after round 2.
// Prints PASS / FAIL per check, the number of co-resident 8-CTA clusters at the real kernel's shared-memory footprint, and
// the cycles from "push issued" to "MMA of the pushed tile committed".
//
// Written without a GPU at hand (round 1, budget spent): the PTX forms are the ones bm_tc.cu uses; what is NEW here is their
// combination in a cluster of 8.  A hang is a result too (run under `timeout`).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int THREADS = 160;                       // warps 0-3: TMEM readers (one lane quarter each), warp 4: MMA / push
constexpr int BN = 128;                            // accumulator columns; every CTA of a pair holds BN / 2 rows of B
constexpr int A_BYTES = 128 * 64 * 2;              // 16 KiB: 128 rows x 64 k, bf16
constexpr int B_BYTES = (BN / 2) * 64 * 2;         // 8 KiB
constexpr int SMEM_BYTES = 200 * 1024;             // the real kernel's footprint (co-residency question)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta)); return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void push_tile(uint32_t dst_cluster_addr, uint32_t src_addr, uint32_t bytes, uint32_t dst_bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_addr), "r"(bytes), "r"(dst_bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, 128-byte swizzle: 8-row groups 1024 B apart (as bm_tc.cu::make_smem_desc)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {      // arrives on `bar` in the CTAs of `mask`
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Result { int bad_round1, bad_round2; float sample1, sample2; unsigned long long push_to_commit; };

__global__ void __launch_bounds__(THREADS, 1) pair_mma_kernel(Result* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* A1 = smem;                               // round-1 A tile: written locally (ones)
    uint8_t* A2 = smem + A_BYTES;                     // round-2 A tile: ARRIVES from rank + 2 (twos)
    uint8_t* Bt = smem + 2 * A_BYTES;                 // this CTA's half of the B tile (ones)
    uint8_t* SRC = smem + 2 * A_BYTES + B_BYTES;      // what this CTA pushes to rank - 2 (twos)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * A_BYTES + B_BYTES);
    uint64_t* mma_done = bars;                        // [2] one per round; multicast commit arrives in both CTAs of the pair
    uint64_t* a2_full = bars + 2;                     // leader: own expect_tx arrive + the peer's arrive; tx = its A2 bytes
    uint64_t* a2_peer = bars + 3;                     // non-leader: its own A2 bytes
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = rank >> 1;
    const bool leader = (rank & 1u) == 0;

    const uint32_t one2 = 0x3F803F80u, two2 = 0x40004000u;            // bf16 pairs (1.0, 1.0) / (2.0, 2.0)
    for (int i = threadIdx.x; i < A_BYTES / 4; i += THREADS) { reinterpret_cast<uint32_t*>(A1)[i] = one2; reinterpret_cast<uint32_t*>(SRC)[i] = two2; }
    for (int i = threadIdx.x; i < B_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(Bt)[i] = one2;
    if (threadIdx.x == 0) {
        mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1);
        mbar_init(a2_full, 2);                         // leader's: its own expect_tx arrive + the peer CTA's "my A2 has landed"
        mbar_init(a2_peer, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy tile writes -> visible to the MMA / bulk-copy units
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // D = f32, A = B = bf16, both K-major, N = BN, M = 256 (cute::UMMA::InstrDescriptor, as bm_tc.cu)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint16_t pair_mask = (uint16_t)(3u << (2u * pair));

    unsigned long long t_push = 0, t_commit = 0;
    // ---- arm the receive barriers, then push this CTA's SRC tile into A2 of the CTA two ranks down (same parity) -------
    if (warp == 4 && lane == 0) {
        if (leader) mbar_expect_tx(a2_full, A_BYTES); else mbar_expect_tx(a2_peer, A_BYTES);
    }
    cluster_sync_all();                                // every receiver is armed before any push
    if (warp == 4 && lane == 0) {
        const uint32_t dst = (rank + 8u - 2u) & 7u;
        const bool dst_leader = (dst & 1u) == 0;
        t_push = clock64();
        push_tile(mapa(smem_u32(A2), dst), smem_u32(SRC), A_BYTES, mapa(smem_u32(dst_leader ? a2_full : a2_peer), dst));
    }
    // ---- round 1: local tiles ------------------------------------------------------------------------------------------
    if (leader && warp == 4) {
        if (lane == 0) {
            const uint64_t da = make_smem_desc(smem_u32(A1)), db = make_smem_desc(smem_u32(Bt));
            for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
            umma_commit_2sm(&mma_done[0], pair_mask);
        }
        __syncwarp();
    }
    mbar_wait(&mma_done[0], 0);
    tc_fence_after();
    int bad1 = 0; float s1 = 0.f;
    if (warp < 4) {
        for (int c = 0; c < BN; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) { const float f = __uint_as_float(v[j]); if (f != 64.0f) ++bad1; s1 = f; }
        }
    }
    // ---- round 2: the A tile that arrived through DSMEM (non-leader tells the leader when its half has landed) ----------
    if (!leader && warp == 4 && lane == 0) {
        mbar_wait(a2_peer, 0);
        mbar_arrive_remote(a2_full, rank & ~1u);
    }
    tc_fence_before();
    __syncthreads();                                   // the TMEM reads of round 1 are done before the accumulator is touched again
    if (leader && warp == 4) {
        if (lane == 0) {
            mbar_wait(a2_full, 0);                     // own bytes (tx) + own arrive + the peer's arrive
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(A2)), db = make_smem_desc(smem_u32(Bt));
            for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            umma_commit_2sm(&mma_done[1], pair_mask);
        }
        __syncwarp();
    }
    mbar_wait(&mma_done[1], 0);
    if (warp == 4 && lane == 0) t_commit = clock64();
    tc_fence_after();
    int bad2 = 0; float s2 = 0.f;
    if (warp < 4) {
        for (int c = 0; c < BN; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) { const float f = __uint_as_float(v[j]); if (f != 192.0f) ++bad2; s2 = f; }
        }
    }
    __shared__ int sbad[2];
    if (threadIdx.x == 0) { sbad[0] = 0; sbad[1] = 0; }
    __syncthreads();
    if (bad1) atomicAdd(&sbad[0], bad1);
    if (bad2) atomicAdd(&sbad[1], bad2);
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x].bad_round1 = sbad[0]; out[blockIdx.x].bad_round2 = sbad[1]; }
    if (threadIdx.x == 5) { out[blockIdx.x].sample1 = s1; out[blockIdx.x].sample2 = s2; }
    if (warp == 4 && lane == 0) out[blockIdx.x].push_to_commit = t_commit - t_push;
    tc_fence_before();
    cluster_sync_all();                                // nobody frees TMEM / exits while a peer may still signal or read it
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
    }
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    CK(cudaFuncSetAttribute(pair_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CK(cudaFuncSetAttribute(pair_mma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc{};
    lc.blockDim = dim3(THREADS); lc.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    lc.gridDim = dim3(8);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, pair_mma_kernel, &lc);
    printf("8-CTA clusters co-resident at %d KB of shared memory per CTA: %d (%s); the resident chain needs 16\n",
           SMEM_BYTES / 1024, max_clusters, cudaGetErrorString(e));
    if (e != cudaSuccess || max_clusters == 0) return 1;
    const int n_clusters = max_clusters < 16 ? max_clusters : 16;
    lc.gridDim = dim3(8 * n_clusters);
    Result* d_out;
    CK(cudaMalloc(&d_out, sizeof(Result) * 8 * n_clusters));
    CK(cudaMemset(d_out, 0xFF, sizeof(Result) * 8 * n_clusters));
    CK(cudaLaunchKernelEx(&lc, pair_mma_kernel, d_out));
    CK(cudaDeviceSynchronize());
    std::vector<Result> h(8 * n_clusters);
    CK(cudaMemcpy(h.data(), d_out, sizeof(Result) * h.size(), cudaMemcpyDeviceToHost));
    long bad1 = 0, bad2 = 0;
    unsigned long long lat_min = ~0ull, lat_max = 0;
    for (const Result& r : h) {
        bad1 += r.bad_round1; bad2 += r.bad_round2;
        if (r.push_to_commit < lat_min) lat_min = r.push_to_commit;
        if (r.push_to_commit > lat_max) lat_max = r.push_to_commit;
    }
    printf("round 1 (local tiles, 4 pairs x cta_group::2 in a cluster of 8, commit mask 3 << 2j): %s (wrong elements: %ld, sample %.1f, want 64)\n",
           bad1 == 0 ? "PASS" : "FAIL", bad1, h[0].sample1);
    printf("round 2 (A tile pushed through DSMEM from rank + 2, accumulated on top):              %s (wrong elements: %ld, sample %.1f, want 192)\n",
           bad2 == 0 ? "PASS" : "FAIL", bad2, h[0].sample2);
    printf("push issued -> commit of the MMAs that read the pushed tile: %llu .. %llu cycles (includes round 1 and its TMEM read-back)\n",
           lat_min, lat_max);
    return (bad1 == 0 && bad2 == 0) ? 0 : 2;
}
