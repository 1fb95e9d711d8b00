/* CPU oracle for the VQ codebook scan  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the arg-max the reference performs in
 *   fourm/vq/quantizers/quantize_lucid.py:388-403 (CosineSimCodebook.forward, eval) and
 *   fourm/vq/quantizers/quantize_lucid.py:263-280 (EuclideanCodebook.forward, eval):
 * fp32 scores, ties resolved to the LOWEST index (torch.argmax semantics).
 * Pinned against the reference's own outputs in tests/golden/vq_golden.pt["scan"].
 * Inputs are expected already l2-normalised for the cosine variant (the normalisation itself is
 * restated in oracle/vq_oracle.py; the product does it on the GPU).
 *
 * Build (see oracle/Makefile):  gcc -O2 -fPIC -shared -o oracle/_build/libvq_argmax_oracle.so oracle/vq_argmax.c
 */
#include <stdint.h>
#include <stddef.h>

/* scores[k] = sum_j z[j] * e[k][j], accumulated left to right in fp32 (no FMA contraction assumed;
 * the tests accept index differences only where the fp64 scores differ by <= 4 fp32 ulps). */
void vq_cosine_argmax_oracle(const float* z, const float* embed, int64_t n, int64_t K, int64_t d,
                             int64_t* idx_out, float* best_out) {
    for (int64_t i = 0; i < n; ++i) {
        const float* zi = z + i * d;
        float best = 0.0f; int64_t arg = -1;
        for (int64_t k = 0; k < K; ++k) {
            const float* ek = embed + k * d;
            float s = 0.0f;
            for (int64_t j = 0; j < d; ++j) s += zi[j] * ek[j];
            if (arg < 0 || s > best) { best = s; arg = k; }   /* strict > keeps the lowest index on ties */
        }
        idx_out[i] = arg;
        if (best_out) best_out[i] = best;
    }
}

/* dist[k] = -(|z|^2 - 2 z.e_k + |e_k|^2), same expression order as the reference. */
void vq_euclid_argmax_oracle(const float* z, const float* embed, int64_t n, int64_t K, int64_t d,
                             int64_t* idx_out, float* best_out) {
    for (int64_t i = 0; i < n; ++i) {
        const float* zi = z + i * d;
        float zz = 0.0f;
        for (int64_t j = 0; j < d; ++j) zz += zi[j] * zi[j];
        float best = 0.0f; int64_t arg = -1;
        for (int64_t k = 0; k < K; ++k) {
            const float* ek = embed + k * d;
            float ze = 0.0f, ee = 0.0f;
            for (int64_t j = 0; j < d; ++j) { ze += zi[j] * ek[j]; ee += ek[j] * ek[j]; }
            float s = -((zz - 2.0f * ze) + ee);
            if (arg < 0 || s > best) { best = s; arg = k; }
        }
        idx_out[i] = arg;
        if (best_out) best_out[i] = best;
    }
}
