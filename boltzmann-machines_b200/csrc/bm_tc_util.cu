// Small CUDA-core kernels that operate on the bf16 activations of the tensor-core path:
// conversions, the three column statistics of a CD step (base_rbm.py:451-457) and the split-K
// reduction of the dW partials.
#include "bm_tc.h"

namespace bm {

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
