// Micro-benchmark for the round-2 plan (DESIGN.md 5.1 "cluster-resident chain"): how fast and with what latency can a
// CTA push 16 KB tiles into the shared memory of other CTAs of an 8-CTA cluster with
// cp.async.bulk.shared::cluster.shared::cta (completion on the DESTINATION's mbarrier), each CTA holding ~200 KB of
// dynamic shared memory and 640 threads like tc_program_kernel?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dsmem_push tools/ubench/dsmem_push.cu && ./dsmem_push
//
// Prints, per cluster size: how many clusters fit on the chip, the one-tile latency (issue -> destination barrier
// phase flips, SM cycles) and the steady push throughput per CTA (bytes / cycle) for 1, 2 and 4 destinations.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int TILE_BYTES = 16 * 1024;
constexpr int SLOTS = 6;                      // receive ring per CTA
constexpr int SMEM_BYTES = 200 * 1024;
constexpr int THREADS = 640;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
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
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// local shared memory -> shared memory of another CTA of the cluster; bytes complete on the destination CTA's barrier
__device__ __forceinline__ void push_tile(uint32_t dst_cluster_addr, uint32_t src_addr, uint32_t bytes, uint32_t dst_bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_addr), "r"(bytes), "r"(dst_bar_cluster_addr) : "memory");
}

struct Result { unsigned long long latency, cycles; unsigned long long bytes; };

// Every CTA is a producer (thread 0 pushes) and a consumer (thread 32 re-arms its ring slots).  Producer r pushes to
// consumers (r + 2 * d) % n for d = 1..n_dst (same parity of rank = "the CTA of the same rank in another pair").
__global__ void __launch_bounds__(THREADS, 1) push_kernel(int n_dst, int rounds, Result* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* src = smem;                                               // one tile to send
    uint8_t* ring = smem + TILE_BYTES;                                 // SLOTS tiles received
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TILE_BYTES * (1 + SLOTS));
    const uint32_t rank = cluster_ctarank(), n = cluster_nctarank();
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TILE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(src)[i] = rank * 1000u + i;
    __syncthreads();
    if (threadIdx.x == 32) for (int s = 0; s < SLOTS; ++s) mbar_expect_tx(&full[s], (uint32_t)TILE_BYTES * (uint32_t)n_dst);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster_sync_all();

    unsigned long long lat = 0, t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        // ---- latency: one tile to one destination, wait for the REMOTE barrier by polling it through DSMEM ----------
        // (the real kernel waits on the consumer side; here the producer measures issue -> remote phase flip)
        t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            const int slot = r % SLOTS;
            for (int d = 1; d <= n_dst; ++d) {
                const uint32_t dst = (rank + 2u * d) % n;
                push_tile(mapa(smem_u32(ring + slot * TILE_BYTES), dst), smem_u32(src), TILE_BYTES, mapa(smem_u32(&full[slot]), dst));
            }
            // flow control stand-in: never run more than SLOTS pushes ahead of the LOCAL consumer (symmetric traffic:
            // what arrives here mirrors what leaves).  Good enough to measure with; the real kernel needs credits from
            // the destination (a remote arrive on the producer's `empty` barrier) before a slot is reused.
            if (r >= SLOTS - 1) {
                const int w = (r - (SLOTS - 1)) % SLOTS;
                const uint32_t par = ((r - (SLOTS - 1)) / SLOTS) & 1u;
                while (!mbar_test(&full[w], par)) { }
                mbar_expect_tx(&full[w], (uint32_t)TILE_BYTES * (uint32_t)n_dst);
            }
        }
        t1 = clock64();
    }
    cluster_sync_all();
    __syncthreads();
    // ---- isolated latency: every CTA pushes ONE tile to (rank + 2) % n and waits for the tile from (rank - 2) % n ----
    cluster_sync_all();
    if (threadIdx.x == 0) {
        // ring slot 0 was re-armed an unknown number of times above: use a dedicated barrier
        uint64_t* lbar = &full[SLOTS];
        mbar_init(lbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(lbar, TILE_BYTES);
    }
    cluster_sync_all();
    if (threadIdx.x == 0) {
        const uint32_t dst = (rank + 2u) % n;
        const unsigned long long a = clock64();
        push_tile(mapa(smem_u32(ring), dst), smem_u32(src), TILE_BYTES, mapa(smem_u32(&full[SLOTS]), dst));
        while (!mbar_test(&full[SLOTS], 0)) { }
        lat = clock64() - a;
        out[blockIdx.x].latency = lat;
        out[blockIdx.x].cycles = t1 - t0;
        out[blockIdx.x].bytes = (unsigned long long)rounds * n_dst * TILE_BYTES;
    }
    cluster_sync_all();
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    CK(cudaFuncSetAttribute(push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    for (int csz : {2, 4, 8}) {
        cudaLaunchConfig_t lc{};
        lc.blockDim = dim3(THREADS); lc.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        lc.gridDim = dim3(csz);
        int max_clusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, push_kernel, &lc);
        printf("cluster size %d: max co-resident clusters = %d (%s) -> %d SMs usable\n", csz, max_clusters, cudaGetErrorString(e), max_clusters * csz);
        if (e != cudaSuccess || max_clusters == 0) continue;
        const int n_clusters = max_clusters;
        lc.gridDim = dim3(n_clusters * csz);
        Result* d_out;
        CK(cudaMalloc(&d_out, sizeof(Result) * n_clusters * csz));
        for (int n_dst : {1, 2, 3}) {
            if (2 * n_dst >= csz && csz > 2) { if (n_dst > 1) continue; }
            if (csz == 2 && n_dst > 1) continue;
            const int rounds = 200;
            CK(cudaMemset(d_out, 0, sizeof(Result) * n_clusters * csz));
            CK(cudaLaunchKernelEx(&lc, push_kernel, n_dst, rounds, d_out));
            CK(cudaDeviceSynchronize());
            std::vector<Result> h(n_clusters * csz);
            CK(cudaMemcpy(h.data(), d_out, sizeof(Result) * h.size(), cudaMemcpyDeviceToHost));
            double lat = 0, bpc = 0;
            for (auto& r : h) { lat += (double)r.latency; bpc += r.cycles ? (double)r.bytes / (double)r.cycles : 0.0; }
            printf("  %d destination(s): one-tile latency %.0f cycles (mean over CTAs), steady push %.1f B/clk per CTA (all %d CTAs pushing)\n",
                   n_dst, lat / h.size(), bpc / h.size(), n_clusters * csz);
        }
        cudaFree(d_out);
    }
    return 0;
}
