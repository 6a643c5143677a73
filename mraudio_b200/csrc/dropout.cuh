// Counter-based dropout for the training forward / backward (model.train() of utils/trainer.py:110: hidden_dropout_prob and
// attention_probs_dropout_prob = 0.1 in the Q-Former's BertConfig; HF port modeling_instructblip.py:530,551,608,781).
//
// No mask is ever stored: every kernel that needs one regenerates it from (seed, stream, element index) with Philox4x32-10,
// so the forward and the backward see the same mask by construction, and oracle/qformer_oracle.py reproduces it bit for bit.
// One Philox call yields 16 bytes; an element is KEPT iff its byte >= thr8 = round(p * 256), i.e. the effective drop
// probability is thr8 / 256 (0.1 -> 26 / 256 = 0.1016) and kept values are scaled by 256 / (256 - thr8) so that the
// expectation is exact.
//
//   hidden sites (embeddings / self-output / cross-output / FFN-output dense): element (m, n) of the [tokens, H] matrix in the
//     split layout (queries of all rows first, then text): call index m * (H / 8) + n / 8, byte n % 8 (of the first 8);
//   attention probabilities: element (row r, head h, query i, key j): call index
//     (((r * heads + h) * Sq + i) * ceil(Sk / 64) + j / 64) * 4 + (j % 8) / 2,  byte ((j % 64) / 8) * 2 + (j % 2)
//     -- the 16 values one thread of an mma.m16n8k16 accumulator fragment holds for one query row and one 64-key chunk.
//   stream = site * 256 + layer  (sites: 1 embeddings, 2 self-attention probs, 3 cross-attention probs, 4 self-output,
//   5 cross-output, 6 FFN output).
#pragma once
#include <stdint.h>

namespace mra {

struct DropoutParams {
    uint32_t seed_lo = 0, seed_hi = 0;
    uint32_t thr8 = 0;      // 0 = dropout off
    float scale = 1.f;      // 256 / (256 - thr8)
    uint32_t stream = 0;    // site * 256 + layer
};

enum DropoutSite { DROP_EMB = 1, DROP_SELF_PROBS = 2, DROP_CROSS_PROBS = 3, DROP_SELF_OUT = 4, DROP_CROSS_OUT = 5, DROP_FFN_OUT = 6 };

__host__ __device__ inline DropoutParams dropout_site(const DropoutParams& base, int site, int layer) {
    DropoutParams d = base;
    d.stream = static_cast<uint32_t>(site * 256 + layer);
    return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// the 16 random bytes of call `idx` of this site's stream
__device__ __forceinline__ uint4 dropout_bytes(const DropoutParams& d, uint64_t idx) {
    return philox4x32_10(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), d.stream, 0x6d72u, d.seed_lo, d.seed_hi);
}

__device__ __forceinline__ uint32_t dropout_byte(const uint4& b, int k) {   // k in [0, 16), compile-time constant in the callers
    const uint32_t w = k < 4 ? b.x : (k < 8 ? b.y : (k < 12 ? b.z : b.w));
    return (w >> ((k & 3) * 8)) & 0xFFu;
}

// multiplier of element `k` of the call: scale if kept, 0 if dropped
__device__ __forceinline__ float dropout_mult(const DropoutParams& d, const uint4& b, int k) {
    return dropout_byte(b, k) >= d.thr8 ? d.scale : 0.f;
}

}  // namespace mra
