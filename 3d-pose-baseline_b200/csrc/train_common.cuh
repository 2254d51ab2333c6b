// Pieces of the training step shared by train.cu and the fused small-batch epilogues of tc_gemm.cu.
#pragma once
#include <cstdint>

namespace p3d {
namespace train {

// Everything that changes from step to step lives in device memory (written by set_scalars_kernel, whose
// arguments travel by value), so that the rest of the step is a replayable CUDA graph.
struct StepScalars {
  float alpha;        // TF-Adam step size lr_t * sqrt(1-b2^t) / (1-b1^t)
  float lr_t;         // decayed learning rate
  float keep, inv_keep;
  unsigned step;      // global_step: part of the Philox counter
  unsigned pad;
  unsigned long long seed;
};

// ------------------------------------------------------------------ Philox4x32-10 dropout mask
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}
// counter = (global_row, col/4, layer, step), key = (seed_lo, seed_hi); word col%4 -> u = w * 2^-32;
// keep iff floor(keep_prob + u) >= 1   (tf.nn.dropout's  floor(keep_prob + random_uniform))
__device__ __forceinline__ uint4 dropout_words(uint64_t seed, uint32_t step, uint32_t layer, uint32_t grow, uint32_t c4) {
  return philox4x32_10(make_uint4(grow, c4, layer, step), make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
}
__device__ __forceinline__ uint8_t keep_bit(uint32_t w, float keep) {
  const float u = static_cast<float>(w) * 2.3283064365386963e-10f;
  return floorf(keep + u) >= 1.f ? 1 : 0;
}

}  // namespace train
}  // namespace p3d
