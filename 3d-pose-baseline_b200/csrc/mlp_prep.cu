// Weight preparation for inference (K0) and input packing.
//
// Folds, per layer, what the reference graph applies at run time:
//   tf.clip_by_norm(w, 1)                   src/linear_model.py:108,123,178,189  (whole-matrix Frobenius clip)
//   tf.layers.batch_normalization(training=False)   :112,181,193 (moving statistics, eps 1e-3)
// into  W' = W * clip_scale * s,  b' = (b - moving_mean) * s + beta,  s = gamma / sqrt(moving_var + eps)
// and writes them (a) fp32 [K,N] for the FFMA path and (b) bf16 transposed [N,Kpad] (K contiguous) for
// the tcgen05 path, all layers stacked along rows.
#include "common.cuh"

namespace p3d {
namespace prep {

__global__ void sumsq_kernel(const float* __restrict__ w, size_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double v = w[i];
    acc += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = lane < (blockDim.x >> 5) ? part[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}

struct FoldArgs {
  const float* W;      // [K,N]
  const float* b;      // [N]
  const float* gamma;  // [N] or null
  const float* beta;
  const float* mm;
  const float* mv;
  const double* norm2; // ||W||_F^2 (device) or null when max_norm is off
  float* wfold;        // [K,N]
  __nv_bfloat16* wt;   // [N rows starting at row_off][kpad]
  float* bias_fold;    // [N] starting at row_off
  int K, N, kpad;
};

__global__ void fold_pack_kernel(const FoldArgs a) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  float clip = 1.f;
  if (a.norm2) {
    const float nrm = static_cast<float>(sqrt(*a.norm2));
    clip = 1.f / fmaxf(nrm, 1.f);
  }
  // read [k][n] coalesced along n
  for (int kk = threadIdx.y; kk < 32; kk += blockDim.y) {
    const int k = k0 + kk, n = n0 + threadIdx.x;
    float v = 0.f;
    if (k < a.K && n < a.N) {
      float s = 1.f;
      if (a.gamma) s = a.gamma[n] / sqrtf(a.mv[n] + kBnEps);
      v = a.W[static_cast<size_t>(k) * a.N + n] * clip * s;
      a.wfold[static_cast<size_t>(k) * a.N + n] = v;
    }
    tile[kk][threadIdx.x] = v;
  }
  __syncthreads();
  // write [n][k] coalesced along k
  for (int nn = threadIdx.y; nn < 32; nn += blockDim.y) {
    const int n = n0 + nn, k = k0 + threadIdx.x;
    if (n < a.N && k < a.K) a.wt[static_cast<size_t>(n) * a.kpad + k] = __float2bfloat16_rn(tile[threadIdx.x][nn]);
  }
  if (blockIdx.y == 0 && threadIdx.y == 0) {
    const int n = n0 + threadIdx.x;
    if (n < a.N) {
      float bb = a.b[n];
      if (a.gamma) {
        const float s = a.gamma[n] / sqrtf(a.mv[n] + kBnEps);
        bb = (bb - a.mm[n]) * s + a.beta[n];
      }
      a.bias_fold[n] = bb;
    }
  }
}

// x fp32 [B,32] -> bf16 [B,64], columns 32..63 zero (layer 0 runs as one K=64 swizzle atom)
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, long long B) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;   // one 8-element group
  if (idx >= B * 8) return;
  const long long row = idx >> 3;
  const int g = static_cast<int>(idx & 7);
  uint4 o = make_uint4(0, 0, 0, 0);
  if (g < 4) {
    const float4* xp = reinterpret_cast<const float4*>(x + row * kIn + g * 8);
    const float4 a = __ldg(xp), b = __ldg(xp + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
  }
  reinterpret_cast<uint4*>(xb)[idx] = o;
}

int prepare(p3d_model* m, cudaStream_t st) {
  const int nl = static_cast<int>(m->layers.size());
  P3D_CUDA(cudaMemsetAsync(m->wt_bf16, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(m->rows_total) * m->kpad, st));
  if (m->cfg.max_norm) {
    P3D_CUDA(cudaMemsetAsync(m->norm2, 0, sizeof(double) * nl, st));
    for (int l = 0; l < nl; ++l) {
      const Layer& ly = m->layers[l];
      const size_t n = static_cast<size_t>(ly.K) * ly.N;
      int blocks = static_cast<int>((n + 256 * 8 - 1) / (256 * 8));
      if (blocks > 1024) blocks = 1024;
      sumsq_kernel<<<blocks, 256, 0, st>>>(m->theta + ly.off_w, n, m->norm2 + l);
      P3D_LAUNCH_CHECK();
    }
  }
  for (int l = 0; l < nl; ++l) {
    const Layer& ly = m->layers[l];
    FoldArgs a;
    a.W = m->theta + ly.off_w; a.b = m->theta + ly.off_b;
    if (ly.has_bn) {
      a.gamma = m->theta + ly.off_gamma; a.beta = m->theta + ly.off_beta;
      a.mm = m->moving + ly.off_mm; a.mv = m->moving + ly.off_mv;
    } else {
      a.gamma = a.beta = a.mm = a.mv = nullptr;
    }
    a.norm2 = m->cfg.max_norm ? m->norm2 + l : nullptr;
    a.wfold = m->wfold + ly.off_wfold;
    a.wt = m->wt_bf16 + static_cast<size_t>(ly.row_off) * m->kpad;
    a.bias_fold = m->bias_fold + ly.row_off;
    a.K = ly.K; a.N = ly.N; a.kpad = m->kpad;
    dim3 grid((ly.N + 31) / 32, (ly.K + 31) / 32), block(32, 8);
    fold_pack_kernel<<<grid, block, 0, st>>>(a);
    P3D_LAUNCH_CHECK();
  }
  m->pack_valid = true;
  return P3D_OK;
}

int pack_input(const float* x, __nv_bfloat16* xb, int64_t B, cudaStream_t st) {
  const long long groups = B * 8;
  pack_input_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, st>>>(x, xb, B);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

}  // namespace prep
}  // namespace p3d
