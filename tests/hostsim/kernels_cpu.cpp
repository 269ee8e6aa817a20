// TEST INFRASTRUCTURE ONLY -- CPU interpretation of the kernel launches the stand-in runtime (fake_cudart.cpp) records.
//
// The tensor-core program kernel is interpreted from its DESCRIPTORS (bm_tc_desc.h, the same definitions the kernel reads):
// for every op of a launch, C = sum over operand pairs of (+/-) A_p B_p^T from the tensor maps' views (out-of-range elements
// read as zero, like TMA), then the epilogue the descriptor asks for (scales, bias, sigma, activation, Philox draw from the
// shared counter layout of bm_rng.cuh, bf16 / fp32 outputs, split-K slices).  That is a statement of what a launch MEANS, not
// of how the kernel computes it -- it checks the host's wiring of operands, orientations, scales, sites and buffers against the
// numpy oracles, with no GPU.  The small CUDA-core kernels around it are restated one by one from their sources (cited).
// Kernels without a restatement here are skipped and counted (fakecuda_skipped).
#include <cuda_runtime_api.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#undef __host__
#undef __device__
#undef __forceinline__
#include "../../boltzmann-machines_b200/csrc/bm_rng.cuh"      // host build: defines the three qualifiers away
#include "../../boltzmann-machines_b200/csrc/bm_tc_desc.h"
#include "../../include/bm.h"

namespace fakecuda {

// what fake_encode_tiled stores in the opaque CUtensorMap
struct MapView { const void* base; unsigned long long cols, rows, ld_bytes; unsigned box0, box1; unsigned magic; };
static const unsigned MAP_MAGIC = 0xB200C0DEu;

static inline float bf2f(__nv_bfloat16 v) { unsigned short b; memcpy(&b, &v, 2); unsigned u = (unsigned)b << 16; float f; memcpy(&f, &u, 4); return f; }
static inline __nv_bfloat16 f2bf(float f) {           // round to nearest even, as __float2bfloat16_rn
    unsigned u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { u |= 0x00400000u; } else { u += 0x7fffu + ((u >> 16) & 1u); }
    unsigned short b = (unsigned short)(u >> 16); __nv_bfloat16 v; memcpy(&v, &b, 2); return v;
}
template <class T> static inline T arg(void** a, int i) { return *reinterpret_cast<T*>(a[i]); }

struct View {
    const __nv_bfloat16* p; long cols, rows, ld;
    float at(long r, long c) const { return (r >= 0 && r < rows && c >= 0 && c < cols) ? bf2f(p[r * ld + c]) : 0.f; }   // TMA zero fill
};
static View view_of(const CUtensorMap& tm) {
    MapView mv; memcpy(&mv, &tm, sizeof(mv));
    View v{reinterpret_cast<const __nv_bfloat16*>(mv.base), (long)mv.cols, (long)mv.rows, (long)(mv.ld_bytes / 2)};
    if (mv.magic != MAP_MAGIC) v = View{nullptr, 0, 0, 0};
    return v;
}

static const int ACT_LINEAR = 0, ACT_SIGMOID = 1, ACT_SOFTPLUS = 2;
static const int SMP_NONE = 0, SMP_BERNOULLI = 1, SMP_GAUSSIAN = 2;

// ---- the program kernel, from its descriptors ------------------------------------------------------------------------------
static void run_tc_op(const bm::TcPhase& ph, const bm::TcLaunch& L) {
    const bm::TcPhaseLite& p = ph.l;
    const int shift_all = L.batch_row;
    const int total_chunks = p.chunks[0] + (p.n_pairs > 1 ? p.chunks[1] : 0);
    View A[2], B[2];
    for (int pr = 0; pr < p.n_pairs; ++pr) { A[pr] = view_of(ph.tmA[pr]); B[pr] = view_of(ph.tmB[pr]); }
    std::vector<float> acc((size_t)p.M * p.N);
    for (int split = 0; split < p.splits; ++split) {
        const int c_begin = (int)(((long long)total_chunks * split) / p.splits), c_end = (int)(((long long)total_chunks * (split + 1)) / p.splits);
        std::fill(acc.begin(), acc.end(), 0.f);
        for (int c = c_begin; c < c_end; ++c) {
            const int pr = (c >= p.chunks[0]) ? 1 : 0;
            const int k0 = (c - (pr ? p.chunks[0] : 0)) * 64;
            const int shift = p.a_batch[pr] ? shift_all : 0;
            const float sign = p.a_neg[pr] ? -1.f : 1.f;
            // gather the chunk's operand tiles once
            std::vector<float> At((size_t)p.M * 64), Bt((size_t)p.N * 64);
            for (int m = 0; m < p.M; ++m)
                for (int k = 0; k < 64; ++k)
                    At[(size_t)m * 64 + k] = p.a_mn[pr] ? A[pr].at(p.a_k0[pr] + shift + k0 + k, m) : A[pr].at(p.a_row0[pr] + shift + m, k0 + k);
            for (int n = 0; n < p.N; ++n)
                for (int k = 0; k < 64; ++k)
                    Bt[(size_t)n * 64 + k] = p.b_mn[pr] ? B[pr].at(k0 + k, n) : B[pr].at(n, k0 + k);
            for (int m = 0; m < p.M; ++m) {
                const float* a = &At[(size_t)m * 64];
                float* out = &acc[(size_t)m * p.N];
                for (int n = 0; n < p.N; ++n) {
                    const float* b = &Bt[(size_t)n * 64];
                    float s = 0.f;
                    for (int k = 0; k < 64; ++k) s += a[k] * b[k];
                    out[n] += sign * s;
                }
            }
        }
        // epilogue (bm_tc.cu::chunk_body, generic semantics)
        bm::RngKey rng; rng.k0 = L.k0; rng.k1 = L.k1; rng.tick = L.tick; rng.row0 = L.row0; rng.c2 = p.rng_c2;
        for (int m = 0; m < p.M; ++m)
            for (int nb = 0; nb < p.N; nb += 4) {
                bm::U4 w{0, 0, 0, 0};
                if (p.sample != SMP_NONE) w = bm::site_block(rng, (uint32_t)m, (uint32_t)(nb >> 2));
                const uint32_t words[4] = {w.x, w.y, w.z, w.w};
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                if (p.sample == SMP_GAUSSIAN) {
                    const float u1a = fmaxf(bm::u32_to_unit_float(w.x), 1.0e-7f), u1b = fmaxf(bm::u32_to_unit_float(w.z), 1.0e-7f);
                    const float ra = sqrtf(-2.0f * logf(u1a)), rb = sqrtf(-2.0f * logf(u1b));
                    const float va = 6.2831853071795864769f * bm::u32_to_unit_float(w.y), vb = 6.2831853071795864769f * bm::u32_to_unit_float(w.w);
                    g[0] = sinf(va) * ra; g[1] = cosf(va) * ra; g[2] = sinf(vb) * rb; g[3] = cosf(vb) * rb;
                }
                for (int j = 0; j < 4 && nb + j < p.N; ++j) {
                    const int n = nb + j;
                    float x = p.acc_scale * acc[(size_t)m * p.N + n];
                    if (p.sigma) x *= p.sigma[n];
                    if (p.bias) x += p.bias_scale * p.bias[n];
                    float mean = x;
                    if (p.act == ACT_SIGMOID) mean = 1.0f / (1.0f + expf(-x));
                    else if (p.act == ACT_SOFTPLUS) mean = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
                    float state = mean;
                    if (p.sample == SMP_BERNOULLI) state = (bm::u32_to_unit_float(words[j]) < mean) ? 1.0f : 0.0f;
                    else if (p.sample == SMP_GAUSSIAN) state = mean + (p.noise_sigma ? p.noise_sigma[n] : 1.0f) * g[j];
                    if (p.out_mean_bf) p.out_mean_bf[(size_t)m * p.ld_mean_bf + n] = f2bf(mean);
                    if (p.out_state_bf) p.out_state_bf[(size_t)m * p.ld_state_bf + n] = f2bf(state);
                    if (p.out_f32) (p.out_f32 + (size_t)split * p.split_stride)[(size_t)m * p.ld_f32 + n] = mean;
                }
            }
    }
}
static void k_tc_program(void** a) {
    const bm::TcLaunch& L = *reinterpret_cast<const bm::TcLaunch*>(a[0]);
    if (L.n_phases == 0) { run_tc_op(L.inl, L); return; }
    for (int i = 0; i < L.n_phases; ++i) run_tc_op(L.phases[i], L);      // program order is a topological order of the dependencies
}

// ---- bm_tc_util.cu -----------------------------------------------------------------------------------------------------------
static void k_f32_to_bf16(void** a) {
    const float* src = arg<const float*>(a, 0); const int lds = arg<int>(a, 1); __nv_bfloat16* dst = arg<__nv_bfloat16*>(a, 2);
    const int ldd = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; c += 2) {
            dst[(size_t)r * ldd + c] = f2bf(src[(size_t)r * lds + c]);
            if (c + 1 < cols) dst[(size_t)r * ldd + c + 1] = f2bf(src[(size_t)r * lds + c + 1]);
            else if (c + 1 < ldd) dst[(size_t)r * ldd + c + 1] = f2bf(0.f);
        }
}
static void k_bf16_to_f32(void** a) {
    const __nv_bfloat16* src = arg<const __nv_bfloat16*>(a, 0); const int lds = arg<int>(a, 1); float* dst = arg<float*>(a, 2);
    const int ldd = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) dst[(size_t)r * ldd + c] = bf2f(src[(size_t)r * lds + c]);
}
struct ColsumJobs {                          // bm_tc_util.cu
    const __nv_bfloat16* P[3]; int ldp[3];
    const __nv_bfloat16* Q[3]; int ldq[3];
    float s1[3], s2[3];
    float* out[3];
    int cols[3];
    int rows, n;
};
static void k_colsum_bf16_finish(void** a) {              // the two-stage reduction collapsed: out = s1 sum P + s2 sum Q
    const ColsumJobs& j = *reinterpret_cast<const ColsumJobs*>(a[0]);
    for (int job = 0; job < j.n; ++job)
        for (int c = 0; c < j.cols[job]; ++c) {
            float s = 0.f;
            for (int r = 0; r < j.rows; ++r) {
                s += j.s1[job] * bf2f(j.P[job][(size_t)r * j.ldp[job] + c]);
                if (j.Q[job]) s += j.s2[job] * bf2f(j.Q[job][(size_t)r * j.ldq[job] + c]);
            }
            j.out[job][c] = s;
        }
}

static void k_reduce_partials(void** a) {
    const float* partial = arg<const float*>(a, 0); const size_t stride = arg<size_t>(a, 1); const int splits = arg<int>(a, 2);
    float* G = arg<float*>(a, 3); const size_t n = arg<size_t>(a, 4);
    for (size_t i = 0; i < n; ++i) { float x = partial[i]; for (int s = 1; s < splits; ++s) x += partial[(size_t)s * stride + i]; G[i] = x; }
}
static void k_weight_update_splitk(void** a) {
    const float* partial = arg<const float*>(a, 0); const size_t stride = arg<size_t>(a, 1); const int splits = arg<int>(a, 2); const float g_div = arg<float>(a, 3);
    float* W = arg<float*>(a, 4); float* dW = arg<float*>(a, 5); const int H = arg<int>(a, 6); const size_t n = arg<size_t>(a, 7);
    const float* pen = arg<const float*>(a, 8); const float l2 = arg<float>(a, 9), lr = arg<float>(a, 10), mom = arg<float>(a, 11);
    __nv_bfloat16* Wb = arg<__nv_bfloat16*>(a, 12); const int ldwb = arg<int>(a, 13);
    for (size_t i = 0; i < n; ++i) {
        float g = partial[i];
        for (int s = 1; s < splits; ++s) g += partial[(size_t)s * stride + i];
        const int h = (int)(i % (size_t)H); const size_t v = i / (size_t)H;
        const float d = lr * (mom * dW[i] + (g / g_div - l2 * W[i] - pen[h]));
        const float wn = W[i] + d;
        dW[i] = d; W[i] = wn; Wb[v * (size_t)ldwb + h] = f2bf(wn);
    }
}
static void k_sqdiff_bf16_partial(void** a, dim3 grid) {
    const __nv_bfloat16* P = arg<const __nv_bfloat16*>(a, 0); const int ldp = arg<int>(a, 1); const __nv_bfloat16* Q = arg<const __nv_bfloat16*>(a, 2);
    const int ldq = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5); double* partial = arg<double*>(a, 6);
    double s = 0.0;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) { const float d = bf2f(P[(size_t)r * ldp + c]) - bf2f(Q[(size_t)r * ldq + c]); s += (double)(d * d); }
    for (unsigned i = 0; i < grid.x; ++i) partial[i] = 0.0;
    partial[0] = s;
}
static void k_u8_to_bf16(void** a) {
    const uint8_t* src = arg<const uint8_t*>(a, 0); const int lds = arg<int>(a, 1); __nv_bfloat16* dst = arg<__nv_bfloat16*>(a, 2);
    const int ldd = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) dst[(size_t)r * ldd + c] = f2bf((float)src[(size_t)r * lds + c]);
}

// ---- bm_simt.cu (float instantiations) --------------------------------------------------------------------------------------
template <typename T> struct BiasUpdateT {            // bm_internal.h::BiasUpdate<T>
    int V, H;
    const T* dvb_raw; const T* dhb_raw; const T* qsum;
    T *vb, *hb, *dvb, *dhb, *q_means, *pen;
    T n_div;
    T lr, mom, damp, cost, target;
};
static void k_bias_update_f32(void** a) {
    const BiasUpdateT<float>& u = *reinterpret_cast<const BiasUpdateT<float>*>(a[0]);
    for (int i = 0; i < u.H; ++i) {
        const float q = u.damp * u.q_means[i] + (1.f - u.damp) * u.qsum[i];
        u.q_means[i] = q;
        const float pen = u.cost * (q - u.target);
        u.pen[i] = pen;
        const float g = u.dhb_raw[i] / u.n_div - pen;
        const float d = u.lr * (u.mom * u.dhb[i] + g);
        u.dhb[i] = d; u.hb[i] += d;
    }
    for (int i = 0; i < u.V; ++i) { const float d = u.lr * (u.mom * u.dvb[i] + u.dvb_raw[i] / u.n_div); u.dvb[i] = d; u.vb[i] += d; }
}
static void k_prepare_input_f32(void** a) {
    const float* X = arg<const float*>(a, 0); const int ldx = arg<int>(a, 1); float* Xp = arg<float*>(a, 2); const int ldxp = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); const float* sigma = arg<const float*>(a, 6); const float keep = arg<float>(a, 7);
    const int do_dropout = arg<int>(a, 8); const bm::RngKey rng = arg<bm::RngKey>(a, 9);
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            bm::U4 w{0, 0, 0, 0};
            if (do_dropout) w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            for (int j = 0; j < 4 && cb * 4 + j < cols; ++j) {
                const int c = cb * 4 + j;
                float x = X[(size_t)r * ldx + c];
                if (sigma) x = x / sigma[c];
                if (do_dropout) { const float m = floorf(keep + bm::u32_to_unit_float(words[j])); x = x / keep * m; }
                Xp[(size_t)r * ldxp + c] = x;
            }
        }
}
static void k_colsum_f32(void** a) {
    const float* P = arg<const float*>(a, 0); const int ldp = arg<int>(a, 1); const float* Q = arg<const float*>(a, 2); const int ldq = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); const float s1 = arg<float>(a, 6), s2 = arg<float>(a, 7); float* out = arg<float*>(a, 8);
    for (int c = 0; c < cols; ++c) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) { s += (double)s1 * P[(size_t)r * ldp + c]; if (Q) s += (double)s2 * Q[(size_t)r * ldq + c]; }
        out[c] = (float)s;
    }
}
static double g_sq_sum = 0.0;                                   // sqdiff_partial -> finish_sum travel through the partial buffer
static void k_sqdiff_partial_f32(void** a, dim3 grid) {
    const float* P = arg<const float*>(a, 0); const int ldp = arg<int>(a, 1); const float* Q = arg<const float*>(a, 2); const int ldq = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); double* partial = arg<double*>(a, 6);
    double s = 0.0;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) { double d = P[(size_t)r * ldp + c]; if (Q) d -= Q[(size_t)r * ldq + c]; s += d * d; }
    for (unsigned i = 0; i < grid.x; ++i) partial[i] = 0.0;
    partial[0] = s;
}
static void k_finish_sum(void** a) {
    const double* partial = arg<const double*>(a, 0); const int n = arg<int>(a, 1); const double denom = arg<double>(a, 2); double* out = arg<double*>(a, 3);
    double s = 0.0; for (int i = 0; i < n; ++i) s += partial[i];
    *out = s / denom;
}
static void k_weight_update_f32(void** a) {
    const float* G = arg<const float*>(a, 0); const int ldg = arg<int>(a, 1); const float g_div = arg<float>(a, 2);
    float* W = arg<float*>(a, 3); float* dW = arg<float*>(a, 4); const int V = arg<int>(a, 5), H = arg<int>(a, 6);
    const float* pen = arg<const float*>(a, 7); const float l2 = arg<float>(a, 8), lr = arg<float>(a, 9), mom = arg<float>(a, 10);
    __nv_bfloat16* Wb = arg<__nv_bfloat16*>(a, 11); const int ldwb = arg<int>(a, 12);
    for (int v = 0; v < V; ++v)
        for (int h = 0; h < H; ++h) {
            const size_t i = (size_t)v * H + h;
            const float w = W[i];
            float g = G[(size_t)v * ldg + h] / g_div - l2 * w;
            g = g - pen[h];
            const float d = lr * (mom * dW[i] + g);
            dW[i] = d; W[i] = w + d;
            if (Wb) Wb[(size_t)v * ldwb + h] = f2bf(w + d);
        }
}
static void k_fill_f32(void** a) { float* p = arg<float*>(a, 0); const size_t n = arg<size_t>(a, 1); const float v = arg<float>(a, 2); for (size_t i = 0; i < n; ++i) p[i] = v; }

// ---- bm_dbm.cu (float instantiations) ----------------------------------------------------------------------------------------
static void k_particle_init_f32(void** a) {
    float* out = arg<float*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2), kind = arg<int>(a, 3);
    const float* sigma = arg<const float*>(a, 4); const bm::RngKey rng = arg<bm::RngKey>(a, 5);
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            const bm::U4 w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (kind == BM_UNIT_GAUSSIAN) {
                float u1 = fmaxf(bm::u32_to_unit_float(w.x), 1.0e-7f), v1 = 6.2831853071795864769f * bm::u32_to_unit_float(w.y);
                float u2 = sqrtf(-2.0f * logf(u1)); g[0] = sinf(v1) * u2; g[1] = cosf(v1) * u2;
                u1 = fmaxf(bm::u32_to_unit_float(w.z), 1.0e-7f); v1 = 6.2831853071795864769f * bm::u32_to_unit_float(w.w);
                u2 = sqrtf(-2.0f * logf(u1)); g[2] = sinf(v1) * u2; g[3] = cosf(v1) * u2;
            }
            for (int j = 0; j < 4 && cb * 4 + j < cols; ++j) {
                const int c = cb * 4 + j;
                out[(size_t)r * cols + c] = kind == BM_UNIT_GAUSSIAN ? g[j] * sigma[c] : bm::u32_to_unit_float(words[j]);
            }
        }
}
static void k_dbm_vbias_f32(void** a) {
    const int V = arg<int>(a, 0); const float* x_sum = arg<const float*>(a, 1); const float* v_sum = arg<const float*>(a, 2);
    const float n_rows = arg<float>(a, 3), m_div = arg<float>(a, 4); float* vb = arg<float*>(a, 5); float* dvb = arg<float*>(a, 6);
    const float lr = arg<float>(a, 7), mom = arg<float>(a, 8);
    for (int j = 0; j < V; ++j) { const float g = x_sum[j] / n_rows - v_sum[j] / m_div; const float d = lr * (mom * dvb[j] + g); dvb[j] = d; vb[j] += d; }
}
static void k_dbm_sparsity_bias_f32(void** a) {
    const int H = arg<int>(a, 0), layer = arg<int>(a, 1); const float* mu_sum = arg<const float*>(a, 2); const float* h_sum = arg<const float*>(a, 3);
    const float n_div = arg<float>(a, 4), m_div = arg<float>(a, 5); float* q_means = arg<float*>(a, 6); float* mu_means = arg<float*>(a, 7);
    float* pen = arg<float*>(a, 8); float* hb = arg<float*>(a, 9); float* dhb = arg<float*>(a, 10);
    const float damp = arg<float>(a, 11), cost = arg<float>(a, 12), target = arg<float>(a, 13), lr = arg<float>(a, 14), mom = arg<float>(a, 15);
    for (int j = 0; j < H; ++j) {
        const float q = damp * q_means[j] + (1.f - damp) * h_sum[layer];
        const float mm = damp * mu_means[j] + (1.f - damp) * mu_sum[layer];
        q_means[j] = q; mu_means[j] = mm;
        const float pn = cost * (q - target) + cost * (mm - target);
        pen[j] = pn;
        const float g = mu_sum[j] / n_div - h_sum[j] / m_div - pn;
        const float d = lr * (mom * dhb[j] + g);
        dhb[j] = d; hb[j] += d;
    }
}
static void k_colnorm_f32(void** a) {
    const float* W = arg<const float*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2); float* norm = arg<float*>(a, 3);
    for (int c = 0; c < cols; ++c) { double s = 0.0; for (int r = 0; r < rows; ++r) { const double w = W[(size_t)r * cols + c]; s += w * w; } norm[c] = (float)sqrt(s); }
}
static void k_max_norm_scale_f32(void** a) {
    float* W = arg<float*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2); const float* norm = arg<const float*>(a, 3); const float max_norm = arg<float>(a, 4);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const float n = norm[c]; const float num = n < max_norm ? n : max_norm; const float den = n > 1e-8f ? n : 1e-8f;
            W[(size_t)r * cols + c] = W[(size_t)r * cols + c] * num / den;
        }
}
static void k_dbm_bound_rows_f32(void** a) {
    const float* X = arg<const float*>(a, 0); const int V = arg<int>(a, 1); const float* mu0 = arg<const float*>(a, 2); const int H0 = arg<int>(a, 3);
    const float* mu1 = arg<const float*>(a, 4); const int H1 = arg<int>(a, 5); const float* t1 = arg<const float*>(a, 6); const float* t2 = arg<const float*>(a, 7);
    const float* vb = arg<const float*>(a, 8); const float* hb0 = arg<const float*>(a, 9); const float* hb1 = arg<const float*>(a, 10);
    const int rows = arg<int>(a, 11); double* out = arg<double*>(a, 12);
    for (int r = 0; r < rows; ++r) {
        double acc = 0.0;
        for (int j = 0; j < H0; ++j) {
            const double m = mu0[(size_t)r * H0 + j];
            acc += (double)t1[(size_t)r * H0 + j] * m + m * (double)hb0[j];
            const double s = (double)fminf(fmaxf((float)m, 1e-7f), 1.0f - 1e-7f);
            acc += -s * log(s) - (1.0 - s) * log(1.0 - s);
        }
        for (int j = 0; j < H1; ++j) {
            const double m = mu1[(size_t)r * H1 + j];
            acc += (double)t2[(size_t)r * H1 + j] * m + m * (double)hb1[j];
            const double s = (double)fminf(fmaxf((float)m, 1e-7f), 1.0f - 1e-7f);
            acc += -s * log(s) - (1.0 - s) * log(1.0 - s);
        }
        for (int j = 0; j < V; ++j) acc += (double)X[(size_t)r * V + j] * (double)vb[j];
        out[r] = acc;
    }
}

// ---- bm_dbm_tc.cuh ----------------------------------------------------------------------------------------------------------
static void k_max_abs_diff_bf16(void** a) {
    const __nv_bfloat16* x = arg<const __nv_bfloat16*>(a, 0); const int lda = arg<int>(a, 1); const __nv_bfloat16* y = arg<const __nv_bfloat16*>(a, 2);
    const int ldb = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5); unsigned* out = arg<unsigned*>(a, 6);
    float m = 0.f;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m = fmaxf(m, fabsf(bf2f(x[(size_t)r * lda + c]) - bf2f(y[(size_t)r * ldb + c])));
    unsigned bits; memcpy(&bits, &m, 4);
    if (bits > *out) *out = bits;
}
static void k_mf_chunk_diffs(void** a, dim3 grid) {
    const __nv_bfloat16* hist = arg<const __nv_bfloat16*>(a, 0); const size_t stride = arg<size_t>(a, 1); const int ld = arg<int>(a, 2);
    const int rows = arg<int>(a, 3), cols = arg<int>(a, 4); unsigned* flags = arg<unsigned*>(a, 5);
    for (unsigned j = 0; j < grid.y; ++j) {
        const __nv_bfloat16* x = hist + (size_t)(j + 1) * stride; const __nv_bfloat16* y = hist + (size_t)j * stride;
        float m = 0.f;
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m = fmaxf(m, fabsf(bf2f(x[(size_t)r * ld + c]) - bf2f(y[(size_t)r * ld + c])));
        unsigned bits; memcpy(&bits, &m, 4);
        if (bits > flags[j]) flags[j] = bits;
    }
}
static void k_dbm_grad_combine(void** a) {
    const float* pos = arg<const float*>(a, 0); const int sp = arg<int>(a, 1); const float* neg = arg<const float*>(a, 2); const int sn = arg<int>(a, 3);
    const size_t stride = arg<size_t>(a, 4); const float inv_n = arg<float>(a, 5), inv_m = arg<float>(a, 6); float* G = arg<float*>(a, 7); const size_t n = arg<size_t>(a, 8);
    for (size_t i = 0; i < n; ++i) {
        float x = 0.f, y = 0.f;
        for (int s = 0; s < sp; ++s) x += pos[(size_t)s * stride + i];
        for (int s = 0; s < sn; ++s) y += neg[(size_t)s * stride + i];
        G[i] = x * inv_n - y * inv_m;
    }
}
static inline float softplus_diff(float a, float b, float z) { const float s = 1.0f / (1.0f + expf(-a * z)); return log1pf(s * expm1f((b - a) * z)); }
static void ais_units(const float* pre, int ldp, float beta, __nv_bfloat16* out, int ldo, int rows, int cols, int sample, const bm::RngKey& rng) {
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            bm::U4 w{0, 0, 0, 0};
            if (sample) w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            for (int j = 0; j < 4; ++j) {
                const int c = cb * 4 + j;
                const float z = (pre && c < cols) ? beta * pre[(size_t)r * ldp + c] : 0.f;
                const float pr = 1.0f / (1.0f + expf(-z));
                const float o = (c < cols) ? (sample ? ((bm::u32_to_unit_float(words[j]) < pr) ? 1.0f : 0.0f) : pr) : 0.f;
                out[(size_t)r * ldo + c] = f2bf(o);
            }
        }
}
static void k_ais_unit_bf16(void** a) {
    ais_units(arg<const float*>(a, 0), arg<int>(a, 1), arg<float>(a, 2), arg<__nv_bfloat16*>(a, 3), arg<int>(a, 4), arg<int>(a, 5), arg<int>(a, 6),
              arg<int>(a, 7), arg<bm::RngKey>(a, 8));
}
static void ais_weights(double* logw, float a, float b, const __nv_bfloat16* x, int ldx, int H0, const float* hb0, const float* pa, int V,
                        const float* pb, int H1, int rows) {
    for (int r = 0; r < rows; ++r) {
        float lin = 0.f;
        for (int j = 0; j < H0; ++j) lin += bf2f(x[(size_t)r * ldx + j]) * hb0[j];
        double acc = ((double)b - (double)a) * (double)lin;
        for (int j = 0; j < V; ++j) acc += (double)softplus_diff(a, b, pa[(size_t)r * V + j]);
        for (int j = 0; j < H1; ++j) acc += (double)softplus_diff(a, b, pb[(size_t)r * H1 + j]);
        logw[r] += acc;
    }
}
static void k_ais_accum2_bf16(void** a) {
    ais_weights(arg<double*>(a, 0), arg<float>(a, 1), arg<float>(a, 2), arg<const __nv_bfloat16*>(a, 3), arg<int>(a, 4), arg<int>(a, 5),
                arg<const float*>(a, 6), arg<const float*>(a, 7), arg<int>(a, 8), arg<const float*>(a, 9), arg<int>(a, 10), arg<int>(a, 11));
}
static void k_ais_fused_step(void** a) {
    double* logw = arg<double*>(a, 0); const float fa = arg<float>(a, 1), fb = arg<float>(a, 2), beta_next = arg<float>(a, 3); const int emit = arg<int>(a, 4);
    const __nv_bfloat16* x = arg<const __nv_bfloat16*>(a, 5); const int ldx = arg<int>(a, 6), H0 = arg<int>(a, 7); const float* hb0 = arg<const float*>(a, 8);
    const float* pa = arg<const float*>(a, 9); const int V = arg<int>(a, 10); __nv_bfloat16* va = arg<__nv_bfloat16*>(a, 11); const int ldv = arg<int>(a, 12);
    const int sample_v = arg<int>(a, 13); const bm::RngKey rng_v = arg<bm::RngKey>(a, 14);
    const float* pb = arg<const float*>(a, 15); const int H1 = arg<int>(a, 16); __nv_bfloat16* hc = arg<__nv_bfloat16*>(a, 17); const int ldh = arg<int>(a, 18);
    const int sample_h2 = arg<int>(a, 19); const bm::RngKey rng_h2 = arg<bm::RngKey>(a, 20); const int rows = arg<int>(a, 21);
    ais_weights(logw, fa, fb, x, ldx, H0, hb0, pa, V, pb, H1, rows);
    if (emit) { ais_units(pa, V, beta_next, va, ldv, rows, V, sample_v, rng_v); ais_units(pb, H1, beta_next, hc, ldh, rows, H1, sample_h2, rng_h2); }
}

// ---- dispatch ---------------------------------------------------------------------------------------------------------------
bool execute(const std::string& name, dim3 grid, dim3, void** args) {
    auto has = [&](const char* s) { return name.find(s) != std::string::npos; };
    if (has("tc_program_kernel")) { k_tc_program(args); return true; }
    if (has("f32_to_bf16_kernel")) { k_f32_to_bf16(args); return true; }
    if (has("bf16_to_f32_kernel")) { k_bf16_to_f32(args); return true; }
    if (has("reduce_partials_kernel")) { k_reduce_partials(args); return true; }
    if (has("weight_update_splitk_kernel")) { k_weight_update_splitk(args); return true; }
    if (has("sqdiff_bf16_partial_kernel")) { k_sqdiff_bf16_partial(args, grid); return true; }
    if (has("sqdiff_bf16_finish_kernel")) { k_finish_sum(args); return true; }
    if (has("u8_to_bf16_kernel")) { k_u8_to_bf16(args); return true; }
    if (has("bias_update_kernelIfE")) { k_bias_update_f32(args); return true; }
    if (has("prepare_input_kernelIfE")) { k_prepare_input_f32(args); return true; }
    if (has("colsum_bf16_partial_kernel")) return true;                    // folded into the finish kernel's restatement
    if (has("colsum_bf16_finish_kernel")) { k_colsum_bf16_finish(args); return true; }
    if (has("colsum_kernelIfE")) { k_colsum_f32(args); return true; }
    if (has("sqdiff_partial_kernelIfE")) { k_sqdiff_partial_f32(args, grid); return true; }
    if (has("finish_sum_kernel")) { k_finish_sum(args); return true; }
    if (has("weight_update_kernelIfE")) { k_weight_update_f32(args); return true; }
    if (has("fill_kernelIfE")) { k_fill_f32(args); return true; }
    if (has("particle_init_kernelIfE")) { k_particle_init_f32(args); return true; }
    if (has("dbm_vbias_kernelIfE")) { k_dbm_vbias_f32(args); return true; }
    if (has("dbm_sparsity_bias_kernelIfE")) { k_dbm_sparsity_bias_f32(args); return true; }
    if (has("colnorm_kernelIfE")) { k_colnorm_f32(args); return true; }
    if (has("max_norm_scale_kernelIfE")) { k_max_norm_scale_f32(args); return true; }
    if (has("dbm_bound_rows_kernelIfE")) { k_dbm_bound_rows_f32(args); return true; }
    if (has("max_abs_diff_bf16_kernel")) { k_max_abs_diff_bf16(args); return true; }
    if (has("mf_chunk_diffs_kernel")) { k_mf_chunk_diffs(args, grid); return true; }
    if (has("dbm_grad_combine_kernel")) { k_dbm_grad_combine(args); return true; }
    if (has("ais_unit_bf16_kernel")) { k_ais_unit_bf16(args); return true; }
    if (has("ais_accum2_bf16_kernel")) { k_ais_accum2_bf16(args); return true; }
    if (has("ais_fused_step_kernel")) { k_ais_fused_step(args); return true; }
    return false;
}

}  // namespace fakecuda
