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
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>
#undef __host__
#undef __device__
#undef __forceinline__
#include "../../boltzmann-machines_b200/csrc/bm_internal.h"   // LayerOp<T>, BiasUpdate<T>, RngKey (bm_rng.cuh defines the qualifiers away on the host)
#include "../../boltzmann-machines_b200/csrc/bm_tc_desc.h"

namespace fakecuda {
void report_violation(const std::string& m);
int hazard_mode();
void count_hazard_launch();
void count_unhonoured(long n);
bool dropped_dependency(int op, int d);

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
    const void* addr(long r, long c) const { return (r >= 0 && r < rows && c >= 0 && c < cols) ? (const void*)(p + r * ld + c) : nullptr; }
};
static View view_of(const CUtensorMap& tm) {
    MapView mv; memcpy(&mv, &tm, sizeof(mv));
    View v{reinterpret_cast<const __nv_bfloat16*>(mv.base), (long)mv.cols, (long)mv.rows, (long)(mv.ld_bytes / 2)};
    if (mv.magic != MAP_MAGIC) v = View{nullptr, 0, 0, 0};
    return v;
}

using bm::ACT_LINEAR; using bm::ACT_SIGMOID; using bm::ACT_SOFTPLUS; using bm::SMP_NONE; using bm::SMP_BERNOULLI; using bm::SMP_GAUSSIAN;


// ---- dataflow hazards of a program launch --------------------------------------------------------------------------------------
// On the GPU the units of a program run concurrently on the CTA pairs, ordered ONLY by the dependencies the host declared
// (TcGemm::dep / dep_all -> dep_ctr / dep_groups); interpreting the ops one after the other in program order hides a missing one.
// While the ops are interpreted, every element an op reads or writes is therefore looked up in a shadow of the memory this launch
// touched: a read of something an earlier op of the launch wrote, a write over something an earlier op read or wrote, must be
// covered by a dependency path -- row block by row block: "same row block" dependencies (dep_groups == 0) order (op j, block g)
// before (op i, block g) only, dep_all ones order all of op j before every unit of op i.  Units of ONE op are not checked against
// each other, and no credit is given for the order in which one CTA pair happens to walk its units.
struct Hazards {
    const bm::TcLaunch& L;
    int n, G;                                         // ops; largest row-block count of any op
    std::vector<std::vector<char>> hb;                // [i * G + g] -> [j * G + gj]: (op j, block gj) completes before unit (i, g) starts
    struct Cell { int writer = -1; std::vector<int> readers; };      // ids: op * (G + 1) + (block + 1); block -1 = every block
    std::unordered_map<uintptr_t, Cell> shadow;
    int reported = 0;
    std::vector<std::string> skipped_waits;           // declared dependencies the kernel's wait loop does not honour (see below)
    static constexpr int ROWS = 256;                  // rows per row-block group: CTA pairs of two 128-row tiles

    explicit Hazards(const bm::TcLaunch& l) : L(l), n(l.n_phases), G(1) {
        for (int i = 0; i < n; ++i) G = std::max(G, L.phases[i].l.m_groups);
        hb.assign((size_t)n * G, std::vector<char>((size_t)n * G, 0));
        for (int i = 0; i < n; ++i) {
            const bm::TcPhaseLite& p = L.phases[i].l;
            for (int d = 0; d < p.n_deps; ++d) {
                if (dropped_dependency(i, d)) continue;
                int j = -1;
                for (int q = 0; q < i; ++q) if (L.phases[q].l.done_ctr == p.dep_ctr[d]) j = q;
                if (j < 0) { report("op " + std::to_string(i) + " waits on a counter no earlier op of the launch publishes"); continue; }
                const bm::TcPhaseLite& pj = L.phases[j].l;
                // The kernel's wait loop (bm_tc.cu, "dataflow: wait until the row blocks this unit reads have been written"): an op
                // that consumes its A operand granule by granule (dep_chunk_ctr) SKIPS the unit-level wait of every same-row-block
                // dependency; the granule waits order it after the producer whose granule counters those are, and after nothing else.
                if (p.dep_groups[d] == 0 && p.dep_chunk_ctr && pj.chunk_ctr != p.dep_chunk_ctr) {
                    skipped_waits.push_back("op " + std::to_string(i) + " declares a same-row-block dependency on op " + std::to_string(j) +
                                            " that the kernel does not wait for (the op follows the granules of another producer)");
                    continue;
                }
                if (p.dep_groups[d] != 0 && p.dep_groups[d] != pj.m_groups)
                    report("op " + std::to_string(i) + " waits for " + std::to_string(p.dep_groups[d]) + " row blocks of op " + std::to_string(j) +
                           ", which has " + std::to_string(pj.m_groups));
                for (int g = 0; g < p.m_groups; ++g) {
                    std::vector<char>& me = hb[(size_t)i * G + g];
                    for (int gj = 0; gj < pj.m_groups; ++gj) {
                        if (p.dep_groups[d] == 0 && gj != g) continue;
                        me[(size_t)j * G + gj] = 1;
                        const std::vector<char>& before = hb[(size_t)j * G + gj];
                        for (size_t x = 0; x < before.size(); ++x) me[x] |= before[x];
                    }
                }
            }
        }
    }
    void report(const std::string& m) {
        if (reported++ >= 4) return;
        std::string extra;
        for (const std::string& w : skipped_waits) extra += " [" + w + "]";
        report_violation("tc program dataflow: " + m + extra);
    }
    int id(int op, int g) const { return op * (G + 1) + g + 1; }
    // does the access `prev` happen before every unit (i, g) that makes the new access?  (g == -1: every block of op i)
    bool ordered(int prev, int i, int g) const {
        const int j = prev / (G + 1), gj = prev % (G + 1) - 1;
        if (j == i) return true;
        const int g0 = g < 0 ? 0 : g, g1 = g < 0 ? L.phases[i].l.m_groups : g + 1;
        const int q0 = gj < 0 ? 0 : gj, q1 = gj < 0 ? L.phases[j].l.m_groups : gj + 1;
        for (int a = g0; a < g1; ++a)
            for (int b = q0; b < q1; ++b)
                if (!hb[(size_t)i * G + a][(size_t)j * G + b]) return false;
        return true;
    }
    std::string who(int x) const {
        const int g = x % (G + 1) - 1;
        return "op " + std::to_string(x / (G + 1)) + (g < 0 ? " (every row block)" : " (row block " + std::to_string(g) + ")");
    }
    void read(const void* a, int i, int g) {
        if (!a) return;
        Cell& c = shadow[(uintptr_t)a >> 1];
        const int me = id(i, g);
        if (c.writer >= 0 && !ordered(c.writer, i, g)) report(who(me) + " reads what " + who(c.writer) + " writes, without a dependency path");
        if (c.readers.empty() || c.readers.back() != me) c.readers.push_back(me);
    }
    void write(const void* a, int i, int g) {
        Cell& c = shadow[(uintptr_t)a >> 1];
        const int me = id(i, g);
        if (c.writer >= 0 && !ordered(c.writer, i, g)) report(who(me) + " overwrites what " + who(c.writer) + " wrote, without a dependency path");
        for (int r : c.readers)
            if (!ordered(r, i, g)) { report(who(me) + " overwrites what " + who(r) + " reads, without a dependency path"); break; }
        c.writer = me;
        c.readers.clear();
    }
};

// ---- the program kernel, from its descriptors ------------------------------------------------------------------------------
static void run_tc_op(const bm::TcPhase& ph, const bm::TcLaunch& L, Hazards* hz = nullptr, int op = 0) {
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
            if (hz) {
                for (int m = 0; m < p.M; ++m)
                    for (int k = 0; k < 64; ++k)
                        hz->read(p.a_mn[pr] ? A[pr].addr(p.a_k0[pr] + shift + k0 + k, m) : A[pr].addr(p.a_row0[pr] + shift + m, k0 + k),
                                 op, m / Hazards::ROWS);
                for (int n = 0; n < p.N; ++n)             // every unit of the op reads its column block of B over the split's K range
                    for (int k = 0; k < 64; ++k)
                        hz->read(p.b_mn[pr] ? B[pr].addr(k0 + k, n) : B[pr].addr(n, k0 + k), op, -1);
            }
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
        bm::RngKey rng; rng.k0 = L.k0; rng.k1 = L.k1; rng.tick = L.tick + p.tick_off; rng.row0 = L.row0; rng.c2 = p.rng_c2;
        if (p.ais_kind == 1) {
            // AIS "units" op: z = C + bias; logw[m] += sum_n softplus(b z) - softplus(a z); state = sample(sigmoid(next z))
            for (int m = 0; m < p.M; ++m) {
                double acc_w = 0.0;
                for (int nb = 0; nb < p.N; nb += 4) {
                    bm::U4 w{0, 0, 0, 0};
                    if (p.sample != SMP_NONE) w = bm::site_block(rng, (uint32_t)m, (uint32_t)(nb >> 2));
                    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
                    for (int j = 0; j < 4 && nb + j < p.N; ++j) {
                        const int n = nb + j;
                        const float z = acc[(size_t)m * p.N + n] + (p.bias ? p.bias[n] : 0.f);
                        const float sa = 1.0f / (1.0f + expf(-p.ais_a * z));
                        acc_w += (double)log1pf(sa * expm1f((p.ais_b - p.ais_a) * z));
                        const float pr = 1.0f / (1.0f + expf(-p.ais_next * z));
                        const float st = p.sample == SMP_BERNOULLI ? ((bm::u32_to_unit_float(words[j]) < pr) ? 1.0f : 0.0f) : pr;
                        if (p.out_state_bf) p.out_state_bf[(size_t)m * p.ld_state_bf + n] = f2bf(st);
                        if (hz && p.out_state_bf) hz->write(p.out_state_bf + (size_t)m * p.ld_state_bf + n, op, m / Hazards::ROWS);
                    }
                }
                if (p.ais_logw) p.ais_logw[m] += acc_w;
            }
            continue;
        }
        for (int m = 0; m < p.M; ++m) {
          double lin = 0.0;
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
                    if (p.ais_kind == 2 && p.bias) lin += (double)(state * p.bias[n]);
                    if (p.out_mean_bf) p.out_mean_bf[(size_t)m * p.ld_mean_bf + n] = f2bf(mean);
                    if (p.out_state_bf) p.out_state_bf[(size_t)m * p.ld_state_bf + n] = f2bf(state);
                    if (p.out_f32) (p.out_f32 + (size_t)split * p.split_stride)[(size_t)m * p.ld_f32 + n] = mean;
                    if (n == p.N - 1 && (p.N & 7) != 0) {
                        // Observed on the B200 (round 2): a TMA store box clipped by the tensor edge INSIDE a 16-byte unit leaves
                        // the rest of that unit overwritten.  Modelled as poison, so that nothing may keep data there.
                        const int end = (p.N + 7) & ~7;
                        for (int c = p.N; c < end; ++c) {
                            if (p.out_mean_bf && c < p.ld_mean_bf) p.out_mean_bf[(size_t)m * p.ld_mean_bf + c] = f2bf(NAN);
                            if (p.out_state_bf && c < p.ld_state_bf) p.out_state_bf[(size_t)m * p.ld_state_bf + c] = f2bf(NAN);
                        }
                    }
                    if (hz) {
                        const int g = m / Hazards::ROWS;
                        if (p.out_mean_bf) hz->write(p.out_mean_bf + (size_t)m * p.ld_mean_bf + n, op, g);
                        if (p.out_state_bf) hz->write(p.out_state_bf + (size_t)m * p.ld_state_bf + n, op, g);
                        if (p.out_f32) hz->write(p.out_f32 + (size_t)split * p.split_stride + (size_t)m * p.ld_f32 + n, op, g);
                    }
                }
            }
          // ais_lin = (b - a) / (bias_scale * -log2 e) multiplies sum_n state * (bias_scale * bias * -log2 e) in the kernel
          if (p.ais_kind == 2 && p.ais_logw) p.ais_logw[m] += (double)p.ais_lin * (double)(p.bias_scale * -1.4426950408889634f) * lin;
        }
    }
}
static void k_tc_program(void** a) {
    const bm::TcLaunch& L = *reinterpret_cast<const bm::TcLaunch*>(a[0]);
    if (L.n_phases == 0) { run_tc_op(L.inl, L); return; }
    // program order is a topological order of the dependencies; whether the declared dependencies cover every hazard between
    // the ops is checked on the side (Hazards) -- always with mode 2, with mode 1 only for launches small enough to shadow
    double touched = 0;
    for (int i = 0; i < L.n_phases; ++i) {
        const bm::TcPhaseLite& p = L.phases[i].l;
        touched += ((double)p.M + p.N) * 64.0 * (p.chunks[0] + (p.n_pairs > 1 ? p.chunks[1] : 0)) + (double)p.M * p.N * p.splits;
    }
    const int mode = hazard_mode();
    if (mode == 2 || (mode == 1 && touched < 2.0e7)) {
        Hazards hz(L);
        for (int i = 0; i < L.n_phases; ++i) run_tc_op(L.phases[i], L, &hz, i);
        count_hazard_launch();
        count_unhonoured((long)hz.skipped_waits.size());
        if (getenv("BM_HOSTSIM_TRACE_FOLLOWERS")) {
            int f = 0;
            for (int i = 0; i < L.n_phases; ++i) f += L.phases[i].l.dep_chunk_ctr ? 1 : 0;
            fprintf(stderr, "[hostsim] program of %d ops, %d follow their producer granule by granule\n", L.n_phases, f);
        }
    } else {
        for (int i = 0; i < L.n_phases; ++i) run_tc_op(L.phases[i], L);
    }
}

// ---- bm_tc_util.cu -----------------------------------------------------------------------------------------------------------
static void k_f32_to_bf16(void** a) {
    const float* src = arg<const float*>(a, 0); const int lds = arg<int>(a, 1); __nv_bfloat16* dst = arg<__nv_bfloat16*>(a, 2);
    const int ldd = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; c += 2) {
            dst[(size_t)r * ldd + c] = f2bf(src[(size_t)r * lds + c]);
            if (c + 1 < cols) dst[(size_t)r * ldd + c + 1] = f2bf(src[(size_t)r * lds + c + 1]);
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
    // (the kernel takes its 16-byte path only when stride % 4 == 0 and both pointers are 16-byte aligned, else scalars)
    for (size_t i = 0; i < n; ++i) { float x = partial[i]; for (int s = 1; s < splits; ++s) x += partial[(size_t)s * stride + i]; G[i] = x; }
}
static void k_weight_update_splitk(void** a) {
    const float* partial = arg<const float*>(a, 0); const size_t stride = arg<size_t>(a, 1); const int splits = arg<int>(a, 2); const float g_div = arg<float>(a, 3);
    float* W = arg<float*>(a, 4); float* dW = arg<float*>(a, 5); const int H = arg<int>(a, 6); const size_t n = arg<size_t>(a, 7);
    const float* pen = arg<const float*>(a, 8); const float l2 = arg<float>(a, 9), lr = arg<float>(a, 10), mom = arg<float>(a, 11);
    __nv_bfloat16* Wb = arg<__nv_bfloat16*>(a, 12); const int ldwb = arg<int>(a, 13);
    // the kernel reads / writes 16 bytes per thread: every slice, W, dW on 16-byte boundaries, four elements in one row
    if ((stride & 3) != 0 || (H & 3) != 0 || (ldwb & 3) != 0 || (n & 3) != 0 ||
        ((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dW)) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(Wb) & 7) != 0)
        report_violation("weight_update_splitk_kernel: misaligned vector access (stride / n_hidden / pointers)");
    for (size_t i = 0; i < n; ++i) {
        float g = partial[i];
        for (int s = 1; s < splits; ++s) g += partial[(size_t)s * stride + i];
        const int h = (int)(i % (size_t)H); const size_t v = i / (size_t)H;
        const float d = lr * (mom * dW[i] + (g / g_div - l2 * W[i] - pen[h]));
        const float wn = W[i] + d;
        dW[i] = d; W[i] = wn; Wb[v * (size_t)ldwb + h] = f2bf(wn);
    }
}
// cd_tail_kernel (bm_tc_util.cu): slices of [(V + 2) x H] and of [V] -> the whole parameter update of a CD step
static void k_cd_tail(void** a) {
    const bm::CdTail& t = *reinterpret_cast<const bm::CdTail*>(a[0]);
    const int V = t.V, H = t.H;
    if ((t.stride & 3) != 0 && H % 4 == 0) report_violation("cd_tail: slices not 16-byte aligned although the vector path is taken");
    auto part = [&](size_t i) { float x = t.part[i]; for (int s = 1; s < t.splits; ++s) x += t.part[(size_t)s * t.stride + i]; return x; };
    std::vector<float> pen(H, 0.f), qn(H);
    for (int h = 0; h < H; ++h) {
        float qs = 0.f;
        for (int s = 0; s < t.splits; ++s) qs += t.part[(size_t)s * t.stride + (size_t)(t.srow + 1) * H + h];
        qn[h] = t.damp * t.q_old[h] + (1.0f - t.damp) * (-qs);
        pen[h] = t.cost * (qn[h] - t.target);
    }
    for (size_t v = 0; v < (size_t)V; ++v)
        for (int h = 0; h < H; ++h) {
            const size_t i = v * H + h;
            const float p = t.cost != 0.f ? pen[h] : 0.f;
            const float d = t.lr * (t.mom * t.dW[i] + (part(i) / t.n_div - t.l2 * t.W[i] - p));
            const float wn = t.W[i] + d;
            t.dW[i] = d; t.W[i] = wn; t.Wb[v * (size_t)t.ldwb + h] = f2bf(wn);
        }
    for (int h = 0; h < H; ++h) {
        float ds = 0.f;
        for (int s = 0; s < t.splits; ++s) ds += t.part[(size_t)s * t.stride + (size_t)t.srow * H + h];
        t.q_new[h] = qn[h]; t.pen[h] = pen[h];
        const float d = t.lr * (t.mom * t.dhb[h] + (ds / t.n_div - pen[h]));
        t.dhb[h] = d; t.hb[h] += d;
    }
    for (int v = 0; v < V; ++v) {
        float vs = 0.f;
        for (int s = 0; s < t.vsplits; ++s) vs += t.vpart[(size_t)s * t.vstride + v];
        const float d = t.lr * (t.mom * t.dvb[v] + vs / t.n_div);
        t.dvb[v] = d; t.vb[v] += d;
    }
}
static void k_set_column_pair(void** a) {
    __nv_bfloat16* buf = arg<__nv_bfloat16*>(a, 0); const int ld = arg<int>(a, 1); const size_t rows = arg<size_t>(a, 2);
    const int col0 = arg<int>(a, 3); const float x = arg<float>(a, 4), y = arg<float>(a, 5);
    for (size_t r = 0; r < rows; ++r) { buf[r * (size_t)ld + col0] = f2bf(x); buf[r * (size_t)ld + col0 + 1] = f2bf(y); }
}
static void k_fill_bf16(void** a) {
    __nv_bfloat16* buf = arg<__nv_bfloat16*>(a, 0); const size_t n = arg<size_t>(a, 1); const float v = arg<float>(a, 2);
    for (size_t i = 0; i < n; ++i) buf[i] = f2bf(v);
}
static void k_sqdiff_bf16(void** a, dim3 grid) {
    const __nv_bfloat16* P = arg<const __nv_bfloat16*>(a, 0); const int ldp = arg<int>(a, 1); const __nv_bfloat16* Q = arg<const __nv_bfloat16*>(a, 2);
    const int ldq = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5); double* partial = arg<double*>(a, 6);
    unsigned int* arrived = arg<unsigned int*>(a, 7); const double denom = arg<double>(a, 8); double* out = arg<double*>(a, 9);
    if (*arrived != 0u) report_violation("sqdiff_bf16_kernel: the arrival counter of the previous launch was not reset");
    if (reinterpret_cast<uintptr_t>(P) % 16 || reinterpret_cast<uintptr_t>(Q) % 16 || ldp % 8 || ldq % 8)
        report_violation("sqdiff_bf16_kernel: operands are read as 16-byte vectors (pointer / leading dimension alignment)");
    double s = 0.0;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) { const float d = bf2f(P[(size_t)r * ldp + c]) - bf2f(Q[(size_t)r * ldq + c]); s += (double)(d * d); }
    for (unsigned i = 0; i < grid.x; ++i) partial[i] = 0.0;
    partial[0] = s;
    *out = s / denom;
    *arrived = 0u;
}
static void k_u8_to_bf16(void** a) {
    const uint8_t* src = arg<const uint8_t*>(a, 0); const int lds = arg<int>(a, 1); __nv_bfloat16* dst = arg<__nv_bfloat16*>(a, 2);
    const int ldd = arg<int>(a, 3), rows = arg<int>(a, 4), cols = arg<int>(a, 5);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) dst[(size_t)r * ldd + c] = f2bf((float)src[(size_t)r * lds + c]);
}

// ---- bm_simt.cu and bm_dbm.cu, templated on the storage type like the kernels ------------------------------------------------
template <typename T> static inline T sigmoid_t(T x) { return T(1) / (T(1) + (T)std::exp(-x)); }
template <> inline float sigmoid_t<float>(float x) { return 1.0f / (1.0f + expf(-x)); }
template <typename T> static inline T softplus_t(T x) { return std::fmax(x, T(0)) + (T)std::log1p(std::exp(-std::fabs(x))); }
template <> inline float softplus_t<float>(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
static inline void box_muller_f(uint32_t x0, uint32_t x1, float& n0, float& n1) {
    const float u1 = fmaxf(bm::u32_to_unit_float(x0), 1.0e-7f), v1 = 6.2831853071795864769f * bm::u32_to_unit_float(x1), u2 = sqrtf(-2.0f * logf(u1));
    n0 = sinf(v1) * u2; n1 = cosf(v1) * u2;
}

// layer_op_kernel<T>: C = s1 op(A1) op(B1) + s2 op(A2) op(B2), then the fused epilogue
template <typename T> static void k_layer_op(void** a) {
    const bm::LayerOp<T>& op = *reinterpret_cast<const bm::LayerOp<T>*>(a[0]);
    for (int m = 0; m < op.M; ++m)
        for (int nb = 0; nb < op.N; nb += 4) {
            bm::U4 w{0, 0, 0, 0};
            if (op.sample != SMP_NONE) w = bm::site_block(op.rng, (uint32_t)m, (uint32_t)(nb >> 2));
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (op.sample == SMP_GAUSSIAN) { box_muller_f(w.x, w.y, g[0], g[1]); box_muller_f(w.z, w.w, g[2], g[3]); }
            for (int j = 0; j < 4 && nb + j < op.N; ++j) {
                const int n = nb + j;
                T acc[2] = {T(0), T(0)};
                for (int pair = 0; pair < 2; ++pair) {
                    const T* A = pair ? op.A2 : op.A1; const T* B = pair ? op.B2 : op.B1;
                    const int K = pair ? op.K2 : op.K1, lda = pair ? op.lda2 : op.lda1, ldb = pair ? op.ldb2 : op.ldb1, bt = pair ? op.b2_trans : op.b1_trans;
                    if (K <= 0 || !A) continue;
                    T s = T(0);
                    for (int k = 0; k < K; ++k)
                        s = (T)std::fma(op.a_trans ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k], bt ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n], s);
                    acc[pair] = s;
                }
                T pre = op.acc_scale * (op.s1 * acc[0] + op.s2 * acc[1]);
                if (op.sigma) pre = pre * op.sigma[n];
                if (op.bias) pre = pre + op.bias_scale * op.bias[n];
                T mean = pre;
                if (op.act == ACT_SIGMOID) mean = sigmoid_t<T>(pre);
                else if (op.act == ACT_SOFTPLUS) mean = softplus_t<T>(pre);
                if (op.means) op.means[(size_t)m * op.ldm + n] = mean;
                if (op.states) {
                    T st = mean;
                    if (op.sample == SMP_BERNOULLI) st = (T(bm::u32_to_unit_float(words[j])) < mean) ? T(1) : T(0);
                    else if (op.sample == SMP_GAUSSIAN) st = mean + (op.noise_sigma ? op.noise_sigma[n] : T(1)) * T(g[j]);
                    op.states[(size_t)m * op.lds + n] = st;
                }
            }
        }
}
template <typename T> static void k_colsum(void** a) {
    const T* P = arg<const T*>(a, 0); const int ldp = arg<int>(a, 1); const T* Q = arg<const T*>(a, 2); const int ldq = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); const T s1 = arg<T>(a, 6), s2 = arg<T>(a, 7); T* out = arg<T*>(a, 8);
    for (int c = 0; c < cols; ++c) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) { double v = (double)s1 * (double)P[(size_t)r * ldp + c]; if (Q) v += (double)s2 * (double)Q[(size_t)r * ldq + c]; s += v; }
        out[c] = (T)s;
    }
}
template <typename T> static void k_rowdot(void** a) {
    const T* P = arg<const T*>(a, 0); const int ldp = arg<int>(a, 1); const T* w = arg<const T*>(a, 2); const int rows = arg<int>(a, 3), cols = arg<int>(a, 4); T* out = arg<T*>(a, 5);
    for (int r = 0; r < rows; ++r) { double s = 0.0; for (int c = 0; c < cols; ++c) { double v = (double)P[(size_t)r * ldp + c]; if (w) v *= (double)w[c]; s += v; } out[r] = (T)s; }
}
template <typename T> static void k_fe_visible(void** a) {
    const T* X = arg<const T*>(a, 0); const int ldx = arg<int>(a, 1); const T* vb = arg<const T*>(a, 2); const T* sigma = arg<const T*>(a, 3);
    const int kind = arg<int>(a, 4), rows = arg<int>(a, 5), cols = arg<int>(a, 6); T* out = arg<T*>(a, 7);
    for (int r = 0; r < rows; ++r) {
        double s = 0.0;
        for (int c = 0; c < cols; ++c) {
            const T x = X[(size_t)r * ldx + c];
            if (kind == BM_UNIT_GAUSSIAN) { const T d = x - vb[c] / sigma[c]; s += 0.5 * (double)(d * d); } else s -= (double)(x * vb[c]);
        }
        out[r] = (T)s;
    }
}
template <typename T> static void k_mean_combine(void** a) {
    const T* x = arg<const T*>(a, 0); const T* y = arg<const T*>(a, 1); const double b_sign = arg<double>(a, 2); const int n = arg<int>(a, 3); double* out = arg<double*>(a, 4);
    double s = 0.0;
    for (int i = 0; i < n; ++i) { double v = (double)x[i]; if (y) v += b_sign * (double)y[i]; s += v; }
    *out = s / (double)n;
}
template <typename T> static void k_sqdiff_partial(void** a, dim3 grid) {
    const T* P = arg<const T*>(a, 0); const int ldp = arg<int>(a, 1); const T* Q = arg<const T*>(a, 2); const int ldq = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); double* partial = arg<double*>(a, 6);
    double s = 0.0;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) { double d = (double)P[(size_t)r * ldp + c]; if (Q) d -= (double)Q[(size_t)r * ldq + c]; s += d * d; }
    for (unsigned i = 0; i < grid.x; ++i) partial[i] = 0.0;
    partial[0] = s;
}
static void k_finish_sum(void** a) {
    const double* partial = arg<const double*>(a, 0); const int n = arg<int>(a, 1); const double denom = arg<double>(a, 2); double* out = arg<double*>(a, 3);
    double s = 0.0; for (int i = 0; i < n; ++i) s += partial[i];
    *out = s / denom;
}
template <typename T> static void k_prepare_input(void** a) {
    const T* X = arg<const T*>(a, 0); const int ldx = arg<int>(a, 1); T* Xp = arg<T*>(a, 2); const int ldxp = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); const T* sigma = arg<const T*>(a, 6); const T keep = arg<T>(a, 7);
    const int do_dropout = arg<int>(a, 8); const bm::RngKey rng = arg<bm::RngKey>(a, 9);
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            bm::U4 w{0, 0, 0, 0};
            if (do_dropout) w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            for (int j = 0; j < 4 && cb * 4 + j < cols; ++j) {
                const int c = cb * 4 + j;
                T x = X[(size_t)r * ldx + c];
                if (sigma) x = x / sigma[c];
                if (do_dropout) { const float m = floorf((float)keep + bm::u32_to_unit_float(words[j])); x = x / keep * T(m); }
                Xp[(size_t)r * ldxp + c] = x;
            }
        }
}
template <typename T> static void k_pll_corrupt(void** a) {
    const T* X = arg<const T*>(a, 0); const int ldx = arg<int>(a, 1); T* Xc = arg<T*>(a, 2); const int ldxc = arg<int>(a, 3);
    const int rows = arg<int>(a, 4), cols = arg<int>(a, 5); const bm::RngKey rng = arg<bm::RngKey>(a, 6);
    for (int r = 0; r < rows; ++r) {
        const uint32_t idx = bm::site_block(rng, (uint32_t)r, 0).x % (uint32_t)cols;
        for (int c = 0; c < cols; ++c) { const T x = X[(size_t)r * ldx + c]; Xc[(size_t)r * ldxc + c] = ((uint32_t)c == idx) ? T(1) - x : x; }
    }
}
template <typename T> static void k_bias_update(void** a) {
    const bm::BiasUpdate<T>& u = *reinterpret_cast<const bm::BiasUpdate<T>*>(a[0]);
    for (int i = 0; i < u.H; ++i) {
        const T q = u.damp * u.q_means[i] + (T(1) - u.damp) * u.qsum[i];
        u.q_means[i] = q;
        const T pen = u.cost * (q - u.target);
        u.pen[i] = pen;
        const T g = u.dhb_raw[i] / u.n_div - pen;
        const T d = u.lr * (u.mom * u.dhb[i] + g);
        u.dhb[i] = d; u.hb[i] += d;
    }
    for (int i = 0; i < u.V; ++i) { const T d = u.lr * (u.mom * u.dvb[i] + u.dvb_raw[i] / u.n_div); u.dvb[i] = d; u.vb[i] += d; }
}
template <typename T> static void k_weight_update(void** a) {
    const T* G = arg<const T*>(a, 0); const int ldg = arg<int>(a, 1); const T g_div = arg<T>(a, 2);
    T* W = arg<T*>(a, 3); T* dW = arg<T*>(a, 4); const int V = arg<int>(a, 5), H = arg<int>(a, 6);
    const T* pen = arg<const T*>(a, 7); const T l2 = arg<T>(a, 8), lr = arg<T>(a, 9), mom = arg<T>(a, 10);
    __nv_bfloat16* Wb = arg<__nv_bfloat16*>(a, 11); const int ldwb = arg<int>(a, 12);
    for (int v = 0; v < V; ++v)
        for (int h = 0; h < H; ++h) {
            const size_t i = (size_t)v * H + h;
            const T w = W[i];
            T g = G[(size_t)v * ldg + h] / g_div - l2 * w;
            g = g - pen[h];
            const T d = lr * (mom * dW[i] + g);
            dW[i] = d; W[i] = w + d;
            if (Wb) Wb[(size_t)v * ldwb + h] = f2bf((float)(w + d));
        }
}
template <typename T> static void k_softmax_rows(void** a) {
    T* X = arg<T*>(a, 0); const int ldx = arg<int>(a, 1), rows = arg<int>(a, 2), cols = arg<int>(a, 3); const T scale = arg<T>(a, 4);
    for (int r = 0; r < rows; ++r) {
        T* x = X + (size_t)r * ldx;
        double mx = -1e300; for (int c = 0; c < cols; ++c) mx = std::fmax(mx, (double)x[c]);
        const T m = (T)mx;
        double s = 0.0;
        for (int c = 0; c < cols; ++c) { T e = sizeof(T) == 4 ? (T)expf((float)(x[c] - m)) : (T)std::exp((double)(x[c] - m)); x[c] = e; s += (double)e; }
        const T tot = (T)s;
        for (int c = 0; c < cols; ++c) x[c] = scale * x[c] / tot;
    }
}
template <typename T> static void k_multinomial_rows(void** a) {
    const T* means = arg<const T*>(a, 0); const int ldm = arg<int>(a, 1), rows = arg<int>(a, 2), cols = arg<int>(a, 3), n_draws = arg<int>(a, 4);
    T* counts = arg<T*>(a, 5); const int ldc = arg<int>(a, 6); const bm::RngKey rng = arg<bm::RngKey>(a, 7);
    std::vector<double> cdf(cols); std::vector<int> cnt(cols);
    for (int r = 0; r < rows; ++r) {
        const T* mrow = means + (size_t)r * ldm;
        double tot = 0.0; for (int c = 0; c < cols; ++c) tot += (double)mrow[c];
        double run = 0.0;
        for (int c = 0; c < cols; ++c) { const float pr = (float)((double)mrow[c] / tot); run += (double)pr; cdf[c] = run; }
        const double last = cdf[cols - 1];
        for (int c = 0; c < cols; ++c) { cdf[c] = cdf[c] / last; cnt[c] = 0; }
        for (int d = 0; d < n_draws; ++d) {
            const bm::U4 w = bm::site_block(rng, (uint32_t)r, (uint32_t)(d >> 2));
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            const double u = (double)bm::u32_to_unit_float(words[d & 3]);
            int lo = 0, hi = cols;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] > u) hi = mid; else lo = mid + 1; }
            if (lo > cols - 1) lo = cols - 1;
            cnt[lo]++;
        }
        for (int c = 0; c < cols; ++c) counts[(size_t)r * ldc + c] = (T)cnt[c];
    }
}
static void k_tf_normal_fill_f32(void** a) {
    float* W = arg<float*>(a, 0); const size_t n = arg<size_t>(a, 1); const float stddev = arg<float>(a, 2);
    const uint32_t k0 = arg<uint32_t>(a, 3), k1 = arg<uint32_t>(a, 4), s2lo = arg<uint32_t>(a, 5), s2hi = arg<uint32_t>(a, 6);
    for (size_t blk = 0; blk * 4 < n; ++blk) {
        const bm::U4 w = bm::philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), s2lo, s2hi, k0, k1);
        float g[4]; box_muller_f(w.x, w.y, g[0], g[1]); box_muller_f(w.z, w.w, g[2], g[3]);
        for (int j = 0; j < 4 && blk * 4 + j < n; ++j) W[blk * 4 + j] = g[j] * stddev;
    }
}
static void k_tf_normal_fill_f64(void** a) {
    double* W = arg<double*>(a, 0); const size_t n = arg<size_t>(a, 1); const double stddev = arg<double>(a, 2);
    const uint32_t k0 = arg<uint32_t>(a, 3), k1 = arg<uint32_t>(a, 4), s2lo = arg<uint32_t>(a, 5), s2hi = arg<uint32_t>(a, 6);
    for (size_t blk = 0; blk * 2 < n; ++blk) {
        const bm::U4 w = bm::philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), s2lo, s2hi, k0, k1);
        const double u1 = std::fmax(bm::u64_to_unit_double(w.x, w.y), 1.0e-7), v1 = 6.283185307179586476925286766559 * bm::u64_to_unit_double(w.z, w.w);
        const double u2 = std::sqrt(-2.0 * std::log(u1));
        W[blk * 2] = std::sin(v1) * u2 * stddev;
        if (blk * 2 + 1 < n) W[blk * 2 + 1] = std::cos(v1) * u2 * stddev;
    }
}
template <typename T> static void k_fill(void** a) { T* p = arg<T*>(a, 0); const size_t n = arg<size_t>(a, 1); const T v = arg<T>(a, 2); for (size_t i = 0; i < n; ++i) p[i] = v; }
template <typename T> static void k_u8_to_real(void** a) { const uint8_t* src = arg<const uint8_t*>(a, 0); T* dst = arg<T*>(a, 1); const size_t n = arg<size_t>(a, 2); for (size_t i = 0; i < n; ++i) dst[i] = (T)src[i]; }

// ---- bm_dbm.cu ------------------------------------------------------------------------------------------------------------------
template <typename T> static void k_max_abs_diff(void** a) {
    const T* x = arg<const T*>(a, 0); const T* y = arg<const T*>(a, 1); const size_t n = arg<size_t>(a, 2); unsigned* out = arg<unsigned*>(a, 3);
    float m = 0.f;
    for (size_t i = 0; i < n; ++i) m = fmaxf(m, (float)std::fabs((double)x[i] - (double)y[i]));
    unsigned bits; memcpy(&bits, &m, 4);
    if (bits > *out) *out = bits;
}
template <typename T> static void k_particle_init(void** a) {
    T* out = arg<T*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2), kind = arg<int>(a, 3);
    const T* sigma = arg<const T*>(a, 4); const bm::RngKey rng = arg<bm::RngKey>(a, 5);
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            const bm::U4 w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (kind == BM_UNIT_GAUSSIAN) { box_muller_f(w.x, w.y, g[0], g[1]); box_muller_f(w.z, w.w, g[2], g[3]); }
            for (int j = 0; j < 4 && cb * 4 + j < cols; ++j) {
                const int c = cb * 4 + j;
                out[(size_t)r * cols + c] = kind == BM_UNIT_GAUSSIAN ? (T)g[j] * sigma[c] : (T)bm::u32_to_unit_float(words[j]);
            }
        }
}
template <typename T> static void k_scale_all(void** a) { T* p = arg<T*>(a, 0); const size_t n = arg<size_t>(a, 1); const double* total = arg<const double*>(a, 2); for (size_t i = 0; i < n; ++i) p[i] = (T)((double)p[i] / *total); }
template <typename T> static void k_dbm_vbias(void** a) {
    const int V = arg<int>(a, 0); const T* x_sum = arg<const T*>(a, 1); const T* v_sum = arg<const T*>(a, 2);
    const T n_rows = arg<T>(a, 3), m_div = arg<T>(a, 4); T* vb = arg<T*>(a, 5); T* dvb = arg<T*>(a, 6); const T lr = arg<T>(a, 7), mom = arg<T>(a, 8);
    for (int j = 0; j < V; ++j) { const T g = x_sum[j] / n_rows - v_sum[j] / m_div; const T d = lr * (mom * dvb[j] + g); dvb[j] = d; vb[j] += d; }
}
template <typename T> static void k_dbm_sparsity_bias(void** a) {
    const int H = arg<int>(a, 0), layer = arg<int>(a, 1); const T* mu_sum = arg<const T*>(a, 2); const T* h_sum = arg<const T*>(a, 3);
    const T n_div = arg<T>(a, 4), m_div = arg<T>(a, 5); T* q_means = arg<T*>(a, 6); T* mu_means = arg<T*>(a, 7);
    T* pen = arg<T*>(a, 8); T* hb = arg<T*>(a, 9); T* dhb = arg<T*>(a, 10);
    const T damp = arg<T>(a, 11), cost = arg<T>(a, 12), target = arg<T>(a, 13), lr = arg<T>(a, 14), mom = arg<T>(a, 15);
    const T hs = h_sum[layer], ms = mu_sum[layer];                       // element `layer` of the unit vectors (sic, dbm.py:581-586)
    for (int j = 0; j < H; ++j) {
        const T q = damp * q_means[j] + (T(1) - damp) * hs;
        const T mm = damp * mu_means[j] + (T(1) - damp) * ms;
        q_means[j] = q; mu_means[j] = mm;
        const T pn = cost * (q - target) + cost * (mm - target);
        pen[j] = pn;
        const T g = mu_sum[j] / n_div - h_sum[j] / m_div - pn;
        const T d = lr * (mom * dhb[j] + g);
        dhb[j] = d; hb[j] += d;
    }
}
template <typename T> static void k_colnorm(void** a) {
    const T* W = arg<const T*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2); T* norm = arg<T*>(a, 3);
    for (int c = 0; c < cols; ++c) { double s = 0.0; for (int r = 0; r < rows; ++r) { const double w = (double)W[(size_t)r * cols + c]; s += w * w; } norm[c] = (T)std::sqrt(s); }
}
template <typename T> static void k_max_norm_scale(void** a) {
    T* W = arg<T*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2); const T* norm = arg<const T*>(a, 3); const T max_norm = arg<T>(a, 4);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const T n = norm[c]; const T num = n < max_norm ? n : max_norm; const T den = n > T(1e-8) ? n : T(1e-8);
            W[(size_t)r * cols + c] = W[(size_t)r * cols + c] * num / den;
        }
}
template <typename T> static void k_dbm_bound_rows(void** a) {
    const T* X = arg<const T*>(a, 0); const int V = arg<int>(a, 1); const T* mu0 = arg<const T*>(a, 2); const int H0 = arg<int>(a, 3);
    const T* mu1 = arg<const T*>(a, 4); const int H1 = arg<int>(a, 5); const T* t1 = arg<const T*>(a, 6); const T* t2 = arg<const T*>(a, 7);
    const T* vb = arg<const T*>(a, 8); const T* hb0 = arg<const T*>(a, 9); const T* hb1 = arg<const T*>(a, 10);
    const int rows = arg<int>(a, 11); double* out = arg<double*>(a, 12);
    auto clip = [](double m) { return sizeof(T) == 4 ? (double)fminf(fmaxf((float)m, 1e-7f), 1.0f - 1e-7f) : std::fmin(std::fmax(m, 1e-7), 1.0 - 1e-7); };
    for (int r = 0; r < rows; ++r) {
        double acc = 0.0;
        for (int j = 0; j < H0; ++j) { const double m = (double)mu0[(size_t)r * H0 + j]; acc += (double)t1[(size_t)r * H0 + j] * m + m * (double)hb0[j]; const double s = clip(m); acc += -s * std::log(s) - (1.0 - s) * std::log(1.0 - s); }
        for (int j = 0; j < H1; ++j) { const double m = (double)mu1[(size_t)r * H1 + j]; acc += (double)t2[(size_t)r * H1 + j] * m + m * (double)hb1[j]; const double s = clip(m); acc += -s * std::log(s) - (1.0 - s) * std::log(1.0 - s); }
        for (int j = 0; j < V; ++j) acc += (double)X[(size_t)r * V + j] * (double)vb[j];
        out[r] = acc;
    }
}
template <typename T> static void k_ais_accum(void** a) {
    double* logw = arg<double*>(a, 0); const double sign = arg<double>(a, 1), beta = arg<double>(a, 2); const T* x = arg<const T*>(a, 3);
    const int H0 = arg<int>(a, 4); const T* hb0 = arg<const T*>(a, 5); const T* pa = arg<const T*>(a, 6); const int V = arg<int>(a, 7);
    const T* pb = arg<const T*>(a, 8); const int H1 = arg<int>(a, 9), rows = arg<int>(a, 10);
    for (int r = 0; r < rows; ++r) {
        double acc = 0.0;
        for (int j = 0; j < H0; ++j) acc += beta * (double)x[(size_t)r * H0 + j] * (double)hb0[j];
        for (int j = 0; j < V; ++j) { const double z = beta * (double)pa[(size_t)r * V + j]; acc += std::fmax(z, 0.0) + std::log1p(std::exp(-std::fabs(z))); }
        for (int j = 0; j < H1; ++j) { const double z = beta * (double)pb[(size_t)r * H1 + j]; acc += std::fmax(z, 0.0) + std::log1p(std::exp(-std::fabs(z))); }
        logw[r] += sign * acc;
    }
}
template <typename T> static void k_ais_unit(void** a) {
    const T* pre = arg<const T*>(a, 0); const T beta = arg<T>(a, 1); T* out = arg<T*>(a, 2);
    const int rows = arg<int>(a, 3), cols = arg<int>(a, 4), sample = arg<int>(a, 5); const bm::RngKey rng = arg<bm::RngKey>(a, 6);
    for (int r = 0; r < rows; ++r)
        for (int cb = 0; cb * 4 < cols; ++cb) {
            bm::U4 w{0, 0, 0, 0};
            if (sample) w = bm::site_block(rng, (uint32_t)r, (uint32_t)cb);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            for (int j = 0; j < 4 && cb * 4 + j < cols; ++j) {
                const int c = cb * 4 + j;
                const T z = beta * pre[(size_t)r * cols + c];
                const T pr = T(1) / (T(1) + (sizeof(T) == 4 ? (T)expf(-(float)z) : (T)std::exp(-(double)z)));
                out[(size_t)r * cols + c] = sample ? ((T(bm::u32_to_unit_float(words[j])) < pr) ? T(1) : T(0)) : pr;
            }
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
static void k_transpose_f32_to_bf16(void** a) {
    const float* src = arg<const float*>(a, 0); const int rows = arg<int>(a, 1), cols = arg<int>(a, 2); __nv_bfloat16* dst = arg<__nv_bfloat16*>(a, 3); const int ldd = arg<int>(a, 4);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) dst[(size_t)c * ldd + r] = f2bf(src[(size_t)r * cols + c]);
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
    // templated kernels: "<name>I f|d E" in the mangled name selects the instantiation
#define BOTH(fn, kname, ...) do { if (has(kname "IfE")) { fn<float>(__VA_ARGS__); return true; } if (has(kname "IdE")) { fn<double>(__VA_ARGS__); return true; } } while (0)
    if (has("tc_program_kernel")) { k_tc_program(args); return true; }
    if (has("transpose_f32_to_bf16_kernel")) { k_transpose_f32_to_bf16(args); return true; }     // before its substring below
    if (has("f32_to_bf16_kernel")) { k_f32_to_bf16(args); return true; }
    if (has("bf16_to_f32_kernel")) { k_bf16_to_f32(args); return true; }
    if (has("reduce_partials_kernel")) { k_reduce_partials(args); return true; }
    if (has("weight_update_splitk_kernel")) { k_weight_update_splitk(args); return true; }
    if (has("cd_tail_kernel")) { k_cd_tail(args); return true; }
    if (has("set_column_pair_kernel")) { k_set_column_pair(args); return true; }
    if (has("fill_bf16_kernel")) { k_fill_bf16(args); return true; }
    if (has("sqdiff_bf16_kernel")) { k_sqdiff_bf16(args, grid); return true; }
    if (has("u8_to_bf16_kernel")) { k_u8_to_bf16(args); return true; }
    if (has("colsum_bf16_partial_kernel")) return true;                    // folded into the finish kernel's restatement
    if (has("colsum_bf16_finish_kernel")) { k_colsum_bf16_finish(args); return true; }
    BOTH(k_layer_op, "layer_op_kernel", args);
    BOTH(k_colsum, "colsum_kernel", args);
    BOTH(k_rowdot, "rowdot_kernel", args);
    BOTH(k_fe_visible, "fe_visible_kernel", args);
    BOTH(k_mean_combine, "mean_combine_kernel", args);
    BOTH(k_sqdiff_partial, "sqdiff_partial_kernel", args, grid);
    if (has("finish_sum_kernel")) { k_finish_sum(args); return true; }
    BOTH(k_prepare_input, "prepare_input_kernel", args);
    BOTH(k_pll_corrupt, "pll_corrupt_kernel", args);
    BOTH(k_bias_update, "bias_update_kernel", args);
    BOTH(k_weight_update, "weight_update_kernel", args);
    BOTH(k_softmax_rows, "softmax_rows_kernel", args);
    BOTH(k_multinomial_rows, "multinomial_rows_kernel", args);
    if (has("tf_normal_fill_f32")) { k_tf_normal_fill_f32(args); return true; }
    if (has("tf_normal_fill_f64")) { k_tf_normal_fill_f64(args); return true; }
    BOTH(k_fill, "fill_kernel", args);
    BOTH(k_u8_to_real, "u8_to_real_kernel", args);
    BOTH(k_max_abs_diff, "max_abs_diff_kernel", args);
    BOTH(k_particle_init, "particle_init_kernel", args);
    BOTH(k_scale_all, "scale_all_kernel", args);
    BOTH(k_dbm_vbias, "dbm_vbias_kernel", args);
    BOTH(k_dbm_sparsity_bias, "dbm_sparsity_bias_kernel", args);
    BOTH(k_colnorm, "colnorm_kernel", args);
    BOTH(k_max_norm_scale, "max_norm_scale_kernel", args);
    BOTH(k_dbm_bound_rows, "dbm_bound_rows_kernel", args);
    BOTH(k_ais_accum, "ais_accum_kernel", args);
    BOTH(k_ais_unit, "ais_unit_kernel", args);
    if (has("max_abs_diff_bf16_kernel")) { k_max_abs_diff_bf16(args); return true; }
    if (has("mf_chunk_diffs_kernel")) { k_mf_chunk_diffs(args, grid); return true; }
    if (has("dbm_grad_combine_kernel")) { k_dbm_grad_combine(args); return true; }
    if (has("ais_unit_bf16_kernel")) { k_ais_unit_bf16(args); return true; }
    if (has("ais_accum2_bf16_kernel")) { k_ais_accum2_bf16(args); return true; }
    if (has("ais_fused_step_kernel")) { k_ais_fused_step(args); return true; }
#undef BOTH
    return false;
}

}  // namespace fakecuda
