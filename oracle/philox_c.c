/* C accelerator for oracle/philox.py::site_words -- TEST INFRASTRUCTURE ONLY.
 *
 * Same arithmetic as the numpy version (Philox-4x32-10 as in TF-1.3's
 * philox_random.h; pinned by /root/reference/boltzmann_machines/rbm/tests/
 * test_rbm.py:65,67 through tests/test_oracle_philox.py, which also checks
 * this file against the numpy path word for word).
 * Built by oracle/Makefile into oracle/_build/liboracle.so.
 */
#include <stdint.h>

static inline void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* words[r*cols + c] = lane (c%4) of block(c/4, row0+r, c2, tick; key=seed) */
void oracle_site_words(uint32_t* words, int64_t rows, int64_t cols, int64_t row0,
                       uint32_t c2, uint32_t tick, uint64_t seed) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        uint32_t* row = words + r * cols;
        for (int64_t b = 0; b * 4 < cols; ++b) {
            uint32_t o[4];
            philox_block((uint32_t)b, (uint32_t)(row0 + r), c2, tick, k0, k1, o);
            for (int l = 0; l < 4 && b * 4 + l < cols; ++l) row[b * 4 + l] = o[l];
        }
    }
}

/* Fused Bernoulli half-step tail used by the timed CPU baseline so the
 * element-wise work is one threaded pass like a real CPU kernel would be:
 * x <- pre-activation (in), mean <- sigmoid(x), state <- (u < mean).     */
#include <math.h>
void oracle_sigmoid_sample(const float* pre, float* mean, float* state, int64_t rows, int64_t cols,
                           int64_t row0, uint32_t c2, uint32_t tick, uint64_t seed, int do_sample) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        const float* x = pre + r * cols;
        float* m = mean + r * cols;
        float* s = state ? state + r * cols : 0;
        for (int64_t b = 0; b * 4 < cols; ++b) {
            uint32_t o[4] = {0, 0, 0, 0};
            if (do_sample) philox_block((uint32_t)b, (uint32_t)(row0 + r), c2, tick, k0, k1, o);
            for (int l = 0; l < 4 && b * 4 + l < cols; ++l) {
                int64_t c = b * 4 + l;
                float p = 1.0f / (1.0f + expf(-x[c]));
                m[c] = p;
                if (s) {
                    if (do_sample) {
                        union { uint32_t u; float f; } cv;
                        cv.u = (o[l] & 0x7FFFFFu) | 0x3F800000u;
                        s[c] = (cv.f - 1.0f) < p ? 1.0f : 0.0f;
                    } else s[c] = p;
                }
            }
        }
    }
}
