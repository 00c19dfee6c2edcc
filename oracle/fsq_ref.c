/* ORACLE (TEST INFRASTRUCTURE): plain-C restatement of the integer FSQ index unpack and the
 * 8-term project_out of vector_quantize_pytorch.ResidualFSQ.get_output_from_indices
 * (library 1.17.8, restated; reference call site tts/core/codec/decoder.py:77, construction
 * tts/core/codec/decoder_modules.py:418-420). Evaluation order per SURVEY.md 3.3-1:
 *   acc = 0; for d = 0..7: acc = acc + code_d * W[c][d]; out = acc + b[c]       (all fp32)
 * Built by oracle/Makefile into oracle/_build/libfsq_ref.so with -ffp-contract=off.
 */
#include <stdint.h>

void fsq_codes_ref(const int64_t* ids, int64_t n, float* codes /* [n][8] */) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t basis = 1;
        for (int d = 0; d < 8; ++d) {
            const int64_t digit = (ids[i] / basis) % 4; /* (idx // 4^d) % 4 */
            codes[i * 8 + d] = (float)(digit - 2) / 2.0f; /* (digit - half_width) / half_width */
            basis *= 4;
        }
    }
}

void fsq_lookup_ref(const int64_t* ids, int64_t n, const float* w_out /* [C][8] */,
                    const float* b_out /* [C] */, int64_t C, float* out /* [n][C] */) {
    for (int64_t i = 0; i < n; ++i) {
        float code[8];
        fsq_codes_ref(ids + i, 1, code);
        for (int64_t c = 0; c < C; ++c) {
            volatile float acc = 0.0f;
            for (int d = 0; d < 8; ++d) {
                const float prod = code[d] * w_out[c * 8 + d];
                acc = acc + prod;
            }
            out[i * C + c] = acc + b_out[c];
        }
    }
}
