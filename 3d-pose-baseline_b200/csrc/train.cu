// Training step: model.step(..., isTraining=True)  (src/linear_model.py:129-145,203-235).
//   forward with BATCH statistics (tf.layers.batch_normalization(training=True): momentum .99, eps 1e-3,
//   biased variance, moving averages updated with the step) and TF1 dropout  y/keep * floor(keep+U),
//   loss = mean((y-t)^2) over B*out, backward through clip_by_norm / BN / ReLU / dropout / residual,
//   TF-flavoured Adam with lr = lr0 * 0.96^(global_step/100000).
// P3D_MODE_BF16: every MatMul of the step (forward, tf.gradients wrt activations, weight gradients) runs on the
// tcgen05 tensor cores (tc_gemm.cu) with bf16 operands, fp32 accumulation and fp32 master weights; H, dZ and W
// are consumed in their natural row-major layouts (MN-major UMMA descriptors where the reduction dimension is
// the row index), the forward GEMM's epilogue also produces the BatchNorm column sums.  P3D_MODE_FP32: FFMA GEMMs.
// Data parallel (one process per GPU): rows are sharded, BN statistics and their backward sums are
// all-reduced (SyncBN - required for parity with the single-device reference), and the flat gradient
// buffer is all-reduced once with NCCL before the (replicated) Adam update.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>

#include "common.cuh"

namespace p3d {

namespace train {

// ------------------------------------------------------------------ NCCL through dlopen
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) { set_error("libnccl.so.2 not found: %s", dlerror()); return nullptr; }
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    set_error("libnccl is missing required symbols");
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}
#define P3D_NCCL(expr)                                                                         \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      set_error("%s failed: %s", #expr, ::p3d::train::nccl()->GetErrorString ? ::p3d::train::nccl()->GetErrorString(r__) : "?"); \
      return P3D_ERR_NCCL;                                                                     \
    }                                                                                          \
  } while (0)

static int allreduce(p3d_model* m, void* buf, size_t n, ncclDataType_t dt, cudaStream_t st) {
  if (m->world <= 1) return P3D_OK;
  P3D_NCCL(nccl()->AllReduce(buf, buf, n, dt, ncclSum, static_cast<ncclComm_t>(m->nccl_comm), st));
  return P3D_OK;
}

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

// ------------------------------------------------------------------ kernels
constexpr int RCH = 256;   // rows per block in the column-reduction kernels (block = 32 cols x 8 row lanes)

__global__ void clip_scale_kernel(const double* norm2, float* scale, int n) {
  const int i = threadIdx.x;
  if (i < n) scale[i] = 1.f / fmaxf(static_cast<float>(sqrt(norm2[i])), 1.f);
}

__global__ void colstats_kernel(const float* __restrict__ z, long long B, int L, double* __restrict__ out) {
  __shared__ double s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * RCH;
  const long long r1 = (r0 + RCH < B) ? r0 + RCH : B;
  double a = 0, b = 0;
  if (c < L)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) { const double v = z[r * L + c]; a += v; b += v * v; }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < L) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(out + c, a);
    atomicAdd(out + L + c, b);
  }
}

// mean / biased variance from (sum, sumsq) over the GLOBAL batch; moving averages (momentum .99)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double invB, int L, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ mm, float* __restrict__ mv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L) return;
  const double mu = stats[c] * invB;
  double var = stats[L + c] * invB - mu * mu;
  if (var < 0) var = 0;
  mean[c] = static_cast<float>(mu);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
  mm[c] = mm[c] * kBnMomentum + static_cast<float>(mu) * (1.f - kBnMomentum);
  mv[c] = mv[c] * kBnMomentum + static_cast<float>(var) * (1.f - kBnMomentum);
}

struct ActArgs {
  const float* z; const float* mean; const float* rstd; const float* gamma; const float* beta;
  const float* res; float* h; __nv_bfloat16* hb; uint8_t* mask; const uint8_t* mask_in;
  float keep; unsigned long long seed; unsigned step, layer; long long row0, B; int L; int has_bn; int dropout;
};

// h = dropout(relu(bn(z))) (+res); one thread per 4 columns; also materialises the keep-mask
__global__ void fwd_act_kernel(const ActArgs a) {
  const int L4 = a.L / 4;
  const long long total = a.B * L4;
  const float inv_keep = 1.f / a.keep;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / L4;
    const int c4 = static_cast<int>(i - r * L4), c = c4 * 4;
    const float4 zz = *reinterpret_cast<const float4*>(a.z + r * a.L + c);
    float v[4] = {zz.x, zz.y, zz.z, zz.w};
    if (a.has_bn) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = a.gamma[c + j] * ((v[j] - a.mean[c + j]) * a.rstd[c + j]) + a.beta[c + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    if (a.dropout) {
      uint8_t kb[4];
      if (a.mask_in) {
        const uchar4 mi = *reinterpret_cast<const uchar4*>(a.mask_in + r * a.L + c);
        kb[0] = mi.x; kb[1] = mi.y; kb[2] = mi.z; kb[3] = mi.w;
      } else {
        const uint4 w = dropout_words(a.seed, a.step, a.layer, static_cast<uint32_t>(a.row0 + r), static_cast<uint32_t>(c4));
        kb[0] = keep_bit(w.x, a.keep); kb[1] = keep_bit(w.y, a.keep); kb[2] = keep_bit(w.z, a.keep); kb[3] = keep_bit(w.w, a.keep);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = kb[j] ? v[j] * inv_keep : 0.f;
      *reinterpret_cast<uchar4*>(a.mask + r * a.L + c) = make_uchar4(kb[0], kb[1], kb[2], kb[3]);
    }
    if (a.res) {
      const float4 rr = *reinterpret_cast<const float4*>(a.res + r * a.L + c);
      v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
    }
    *reinterpret_cast<float4*>(a.h + r * a.L + c) = make_float4(v[0], v[1], v[2], v[3]);
    if (a.hb) {      // operand of the next layer's tcgen05 GEMMs (forward and weight gradient)
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(a.hb + r * a.L + c) = pk;
    }
  }
}

// fp32 -> bf16 with a (possibly wider) destination pitch; pad columns are left untouched (zero)
__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int cols, int ld_dst) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    dst[r * ld_dst + c] = __float2bfloat16_rn(src[i]);
  }
}

// dy = 2 (y - t) / (Bg*out), loss accumulator += sum (y-t)^2
__global__ void loss_dy_kernel(const float* __restrict__ y, const float* __restrict__ t, size_t n, float scale,
                               float* __restrict__ dy, double* __restrict__ acc, __nv_bfloat16* __restrict__ dyb, int out, int ld_b) {
  double s = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float d = y[i] - t[i];
    dy[i] = scale * d;
    if (dyb) { const size_t r = i / out; dyb[r * ld_b + (i - r * out)] = __float2bfloat16_rn(scale * d); }
    s += static_cast<double>(d) * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
    atomicAdd(acc, tot);
  }
}
__global__ void finish_step_scalars_kernel(const double* acc, double denom, float lr, float* loss, float* lr_out) {
  if (loss) *loss = static_cast<float>(*acc / denom);
  if (lr_out) *lr_out = lr;
}

struct BwdArgs {
  const float* dh; const float* z; const float* mean; const float* rstd; const float* gamma; const float* beta;
  const uint8_t* mask; float* dz; __nv_bfloat16* dzb; double* sums; float inv_keep; long long B; int L; int has_bn; int dropout;
};

// pass A: da = dh * dropout * relu'(a); column sums of da and da*xhat (for BN backward / dgamma, dbeta)
__global__ void bwd_act_kernel(const BwdArgs a) {
  __shared__ double s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * RCH;
  const long long r1 = (r0 + RCH < a.B) ? r0 + RCH : a.B;
  double p = 0, q = 0;
  if (c < a.L) {
    float mu = 0, rs = 1, g = 1, be = 0;
    if (a.has_bn) { mu = a.mean[c]; rs = a.rstd[c]; g = a.gamma[c]; be = a.beta[c]; }
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float zz = a.z[r * a.L + c];
      const float xh = a.has_bn ? (zz - mu) * rs : 0.f;
      const float act = a.has_bn ? g * xh + be : zz;
      float gr = a.dh[r * a.L + c];
      if (a.dropout) gr = a.mask[r * a.L + c] ? gr * a.inv_keep : 0.f;
      const float da = act > 0.f ? gr : 0.f;
      a.dz[r * a.L + c] = da;
      if (a.dzb) a.dzb[r * a.L + c] = __float2bfloat16_rn(da);     // final dz only when the layer has no BN
      p += da; q += static_cast<double>(da) * xh;
    }
  }
  s1[threadIdx.y][threadIdx.x] = p; s2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < a.L) {
    for (int i = 1; i < 8; ++i) { p += s1[i][threadIdx.x]; q += s2[i][threadIdx.x]; }
    atomicAdd(a.sums + c, p);
    atomicAdd(a.sums + a.L + c, q);
  }
}

// pass B (BN layers): dz = gamma * rstd * (da - mean(da) - xhat * mean(da*xhat)), means over the GLOBAL batch
__global__ void bwd_bn_kernel(float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma,
                              const double* __restrict__ sums, float invB, long long B, int L, __nv_bfloat16* __restrict__ dzb) {
  const long long total = B * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % L);
    const float xh = (z[i] - mean[c]) * rstd[c];
    const float m1 = static_cast<float>(sums[c]) * invB, m2 = static_cast<float>(sums[L + c]) * invB;
    const float v = gamma[c] * rstd[c] * (dz[i] - m1 - xh * m2);
    dz[i] = v;
    if (dzb) dzb[i] = __float2bfloat16_rn(v);
  }
}
__global__ void bn_param_grad_kernel(const double* __restrict__ sums, int L, double scale, float* __restrict__ ggamma, float* __restrict__ gbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < L) { gbeta[c] = static_cast<float>(sums[c] * scale); ggamma[c] = static_cast<float>(sums[L + c] * scale); }
}

__global__ void colsum_kernel(const float* __restrict__ a, long long B, int N, float* __restrict__ out) {
  __shared__ float s1[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * RCH;
  const long long r1 = (r0 + RCH < B) ? r0 + RCH : B;
  float p = 0;
  if (c < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) p += a[r * N + c];
  s1[threadIdx.y][threadIdx.x] = p;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    for (int i = 1; i < 8; ++i) p += s1[i][threadIdx.x];
    atomicAdd(out + c, p);
  }
}

__global__ void dot_kernel(const float* __restrict__ w, const float* __restrict__ g, size_t n, double* __restrict__ out) {
  double s = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    s += static_cast<double>(w[i]) * g[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tot += part[w2];
    atomicAdd(out, tot);
  }
}

// Pull the gradient wrt the clipped weight back through tf.clip_by_norm (src/linear_model.py:108):
//   g = (gc - W <W,gc>/||W||^2) / ||W||  when ||W|| > 1, else g = gc.   In place.
__global__ void clip_grad_kernel(float* __restrict__ grad, const float* __restrict__ w, size_t n, const double* dot,
                                 const double* norm2) {
  const double n2 = *norm2;
  if (n2 <= 1.0) return;
  const float cs = static_cast<float>(1.0 / sqrt(n2)), cd = static_cast<float>(*dot / n2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    grad[i] = (grad[i] - w[i] * cd) * cs;
}

// TF Adam (src/linear_model.py:137): m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= alpha_t m/(sqrt(v)+eps).
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float alpha) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float g = grad[i];
    const float mi = m[i] + (g - m[i]) * 0.1f;
    const float vi = v[i] + (g * g - v[i]) * 0.001f;
    m[i] = mi; v[i] = vi;
    theta[i] -= alpha * mi / (sqrtf(vi) + 1e-8f);
  }
}

// ------------------------------------------------------------------ workspace
void free_workspace(p3d_model* m) {
  TrainWorkspace& w = m->tw;
  cudaFree(w.z); cudaFree(w.h); cudaFree(w.dh); cudaFree(w.dz); cudaFree(w.dres); cudaFree(w.dy);
  cudaFree(w.stats); cudaFree(w.mean); cudaFree(w.rstd); cudaFree(w.scal); cudaFree(w.maskbuf);
  cudaFree(w.xb); cudaFree(w.hb); cudaFree(w.dzb); cudaFree(w.dyb); cudaFree(w.wb);
  w = TrainWorkspace();
  if (m->nccl_comm && nccl()) { nccl()->CommDestroy(static_cast<ncclComm_t>(m->nccl_comm)); m->nccl_comm = nullptr; }
}

constexpr int kOutPad = 48;   // bf16 pitch of dy / W4 rows: 96 B keeps TMA's 16-byte pitch rule for out = 42 too
// bf16 mode runs the step's GEMMs on the tensor cores (tc_gemm.cu); fp32 mode keeps the FFMA GEMMs
static inline bool use_tc(const p3d_model* m) { return m->cfg.mode == P3D_MODE_BF16 && (m->L % 8) == 0; }

static int ensure_workspace(p3d_model* m, int64_t B) {
  TrainWorkspace& w = m->tw;
  const int L = m->L, nh = static_cast<int>(m->layers.size()) - 1, nl = nh + 1;
  if (w.cap_B >= B) return P3D_OK;
  cudaFree(w.z); cudaFree(w.h); cudaFree(w.dh); cudaFree(w.dz); cudaFree(w.dres); cudaFree(w.dy); cudaFree(w.maskbuf);
  cudaFree(w.xb); cudaFree(w.hb); cudaFree(w.dzb); cudaFree(w.dyb);
  w.z = w.h = w.dh = w.dz = w.dres = w.dy = nullptr; w.maskbuf = nullptr; w.cap_B = 0;
  w.xb = w.hb = w.dzb = w.dyb = nullptr;
  const size_t bl = static_cast<size_t>(B) * L;
  P3D_CUDA(cudaMalloc(&w.z, sizeof(float) * bl * nh));
  P3D_CUDA(cudaMalloc(&w.h, sizeof(float) * bl * nh));
  P3D_CUDA(cudaMalloc(&w.dh, sizeof(float) * bl));
  P3D_CUDA(cudaMalloc(&w.dz, sizeof(float) * bl));
  P3D_CUDA(cudaMalloc(&w.dres, sizeof(float) * bl * 2));              // two more rotating gradient buffers
  P3D_CUDA(cudaMalloc(&w.dy, sizeof(float) * static_cast<size_t>(B) * m->out_size));
  P3D_CUDA(cudaMalloc(&w.maskbuf, bl * nh));                          // uint8 keep-masks [nh][B][L]
  if (use_tc(m)) {
    // bf16 operands of the tcgen05 GEMMs, all in their natural row-major layouts
    P3D_CUDA(cudaMalloc(&w.xb, sizeof(__nv_bfloat16) * static_cast<size_t>(B) * kIn));
    P3D_CUDA(cudaMalloc(&w.hb, sizeof(__nv_bfloat16) * bl * nh));
    P3D_CUDA(cudaMalloc(&w.dzb, sizeof(__nv_bfloat16) * bl));
    P3D_CUDA(cudaMalloc(&w.dyb, sizeof(__nv_bfloat16) * static_cast<size_t>(B) * kOutPad));
    P3D_CUDA(cudaMemset(w.dyb, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(B) * kOutPad));   // pad columns stay zero
    if (!w.wb) {
      // [all hidden W, same offsets as theta][W4 with its rows padded to kOutPad columns]
      const Layer& lo = m->layers.back();
      const size_t n = lo.off_w + static_cast<size_t>(lo.K) * kOutPad;
      P3D_CUDA(cudaMalloc(&w.wb, sizeof(__nv_bfloat16) * n));
      P3D_CUDA(cudaMemset(w.wb, 0, sizeof(__nv_bfloat16) * n));
    }
  }
  if (!w.stats) {
    // doubles: stats [nh][2][L] | bwd sums [nh][2][L] | dots [nl] | loss [1]
    P3D_CUDA(cudaMalloc(&w.stats, sizeof(double) * (4ull * nh * L + nl + 1)));
    w.red = w.stats + 2ull * nh * L;
    P3D_CUDA(cudaMalloc(&w.mean, sizeof(float) * static_cast<size_t>(nh) * L));
    P3D_CUDA(cudaMalloc(&w.rstd, sizeof(float) * static_cast<size_t>(nh) * L));
    P3D_CUDA(cudaMalloc(&w.scal, sizeof(float) * (nl + 8)));
  }
  w.cap_B = B;
  return P3D_OK;
}

static inline dim3 colgrid(int cols, int64_t B) { return dim3((cols + 31) / 32, static_cast<unsigned>((B + RCH - 1) / RCH)); }
static inline int egrid(long long n) { long long g = (n + 255) / 256; if (g > 148 * 8) g = 148 * 8; return static_cast<int>(g < 1 ? 1 : g); }

int train_step(p3d_model* m, const float* x, const float* t, int64_t B, float keep, uint64_t seed, const uint8_t* mask_in,
               int64_t Bg, int64_t row0, float* loss, float* lr_used, float* y, cudaStream_t st) {
  using simt::Epilogue;
  using simt::sgemm;
  P3D_TRY(ensure_workspace(m, B));
  TrainWorkspace& w = m->tw;
  const int L = m->L, nlay = static_cast<int>(m->layers.size()), nh = nlay - 1, out = m->out_size;
  const size_t bl = static_cast<size_t>(B) * L;
  const bool dropout = keep < 1.f || mask_in != nullptr;
  const bool clip = m->cfg.max_norm != 0;
  const bool residual = m->cfg.residual != 0;
  const bool tc = use_tc(m);
  using tcg::GemmArgs;
  double* dots = w.red + 2ull * nh * L;
  double* lossacc = dots + nlay;
  uint8_t* maskbuf = w.maskbuf;
  float* scale = w.scal;   // [nlay] clip scales
  const double invBg = 1.0 / static_cast<double>(Bg);

  P3D_CUDA(cudaMemsetAsync(w.stats, 0, sizeof(double) * (4ull * nh * L + nlay + 1), st));
  P3D_CUDA(cudaMemsetAsync(m->grad, 0, sizeof(float) * m->n_train, st));
  if (clip) {
    P3D_CUDA(cudaMemsetAsync(m->norm2, 0, sizeof(double) * nlay, st));
    for (int l = 0; l < nlay; ++l) {
      const Layer& ly = m->layers[l];
      const size_t n = static_cast<size_t>(ly.K) * ly.N;
      dot_kernel<<<egrid(static_cast<long long>(n / 4 + 1)), 256, 0, st>>>(m->theta + ly.off_w, m->theta + ly.off_w, n, m->norm2 + l);
      P3D_LAUNCH_CHECK();
    }
    clip_scale_kernel<<<1, 64, 0, st>>>(m->norm2, scale, nlay);
    P3D_LAUNCH_CHECK();
  }
  if (tc) {
    // bf16 copies of this step's weights (natural [K][N] layout; W4 rows padded to kOutPad) and of x
    const Layer& lo = m->layers[nh];
    to_bf16_kernel<<<egrid(static_cast<long long>(lo.off_w / 4 + 1)), 256, 0, st>>>(m->theta, w.wb, 1, static_cast<int>(lo.off_w), static_cast<int>(lo.off_w));
    P3D_LAUNCH_CHECK();
    to_bf16_kernel<<<egrid(static_cast<long long>(L) * out / 4 + 1), 256, 0, st>>>(m->theta + lo.off_w, w.wb + lo.off_w, L, out, kOutPad);
    P3D_LAUNCH_CHECK();
    to_bf16_kernel<<<egrid(static_cast<long long>(B) * kIn / 4 + 1), 256, 0, st>>>(x, w.xb, B, kIn, kIn);
    P3D_LAUNCH_CHECK();
  }
  // ---------------------------------------------------------------- forward
  for (int li = 0; li < nh; ++li) {
    const Layer& ly = m->layers[li];
    const float* in = li == 0 ? x : w.h + (li - 1) * bl;
    const int lda = li == 0 ? kIn : L;
    float* z = w.z + li * bl;
    double* stats = w.stats + 2ull * li * L;
    if (tc) {
      GemmArgs g;   // z = (x|h) W * clip + b ; BN column sums in the epilogue
      g.M = static_cast<int>(B); g.N = L; g.K = ly.K;
      g.A = li == 0 ? w.xb : w.hb + (li - 1) * bl; g.lda = lda;
      g.B = w.wb + ly.off_w; g.ldb = L; g.b_mn = 1;
      g.C = z; g.ldc = L; g.bias = m->theta + ly.off_b; g.alpha_dev = clip ? scale + li : nullptr;
      g.colsum = ly.has_bn ? stats : nullptr;
      P3D_TRY(tcg::gemm(g, st));
    } else {
      Epilogue e; e.bias = m->theta + ly.off_b; e.alpha_dev = clip ? scale + li : nullptr;
      P3D_TRY(sgemm(false, false, B, L, ly.K, in, lda, m->theta + ly.off_w, L, z, L, e, st));
    }
    float* mean = w.mean + static_cast<size_t>(li) * L;
    float* rstd = w.rstd + static_cast<size_t>(li) * L;
    if (ly.has_bn) {
      if (!tc) {
        colstats_kernel<<<colgrid(L, B), dim3(32, 8), 0, st>>>(z, B, L, stats);
        P3D_LAUNCH_CHECK();
      }
      P3D_TRY(allreduce(m, stats, 2ull * L, ncclDouble, st));
      bn_finalize_kernel<<<(L + 255) / 256, 256, 0, st>>>(stats, invBg, L, mean, rstd, m->moving + ly.off_mm, m->moving + ly.off_mv);
      P3D_LAUNCH_CHECK();
    }
    ActArgs a;
    a.z = z; a.mean = mean; a.rstd = rstd;
    a.gamma = ly.has_bn ? m->theta + ly.off_gamma : nullptr; a.beta = ly.has_bn ? m->theta + ly.off_beta : nullptr;
    a.res = (residual && li >= 2 && (li % 2) == 0) ? w.h + (li - 2) * bl : nullptr;
    a.h = w.h + li * bl; a.hb = tc ? w.hb + li * bl : nullptr; a.mask = maskbuf + li * bl; a.mask_in = mask_in ? mask_in + li * bl : nullptr;
    a.keep = keep; a.seed = seed; a.step = static_cast<unsigned>(m->global_step); a.layer = li;
    a.row0 = row0; a.B = B; a.L = L; a.has_bn = ly.has_bn; a.dropout = dropout;
    fwd_act_kernel<<<egrid(static_cast<long long>(bl / 4)), 256, 0, st>>>(a);
    P3D_LAUNCH_CHECK();
  }
  {
    const Layer& ly = m->layers[nh];
    if (tc) {
      GemmArgs g;
      g.M = static_cast<int>(B); g.N = out; g.K = L;
      g.A = w.hb + (nh - 1) * bl; g.lda = L;
      g.B = w.wb + ly.off_w; g.ldb = kOutPad; g.b_mn = 1;
      g.C = y; g.ldc = out; g.bias = m->theta + ly.off_b; g.alpha_dev = clip ? scale + nh : nullptr;
      P3D_TRY(tcg::gemm(g, st));
    } else {
      Epilogue e; e.bias = m->theta + ly.off_b; e.alpha_dev = clip ? scale + nh : nullptr;
      P3D_TRY(sgemm(false, false, B, out, L, w.h + (nh - 1) * bl, L, m->theta + ly.off_w, out, y, out, e, st));
    }
  }
  const size_t ny = static_cast<size_t>(B) * out;
  loss_dy_kernel<<<egrid(static_cast<long long>(ny)), 256, 0, st>>>(y, t, ny, static_cast<float>(2.0 * invBg / out), w.dy, lossacc,
                                                                          tc ? w.dyb : nullptr, out, kOutPad);
  P3D_LAUNCH_CHECK();
  P3D_TRY(allreduce(m, lossacc, 1, ncclDouble, st));
  // ---------------------------------------------------------------- backward
  float* G[3] = {w.dh, w.dres, w.dres + bl};
  int cur = 0, keepi = -1;
  {
    const Layer& ly = m->layers[nh];
    if (tc) {
      GemmArgs gw;   // dW4 = h^T dy  (reduction over the batch: both operands MN-major, split-K)
      gw.M = L; gw.N = out; gw.K = static_cast<int>(B);
      gw.A = w.hb + (nh - 1) * bl; gw.lda = L; gw.a_mn = 1;
      gw.B = w.dyb; gw.ldb = kOutPad; gw.b_mn = 1;
      gw.C = m->grad + ly.off_w; gw.ldc = out; gw.split_k = 1;
      P3D_TRY(tcg::gemm(gw, st));
    } else {
      Epilogue e0;
      P3D_TRY(sgemm(true, false, L, out, static_cast<int>(B), w.h + (nh - 1) * bl, L, w.dy, out, m->grad + ly.off_w, out, e0, st));
    }
    colsum_kernel<<<colgrid(out, B), dim3(32, 8), 0, st>>>(w.dy, B, out, m->grad + ly.off_b);
    P3D_LAUNCH_CHECK();
    if (tc) {
      GemmArgs gd;   // dh = dy W4^T * clip
      gd.M = static_cast<int>(B); gd.N = L; gd.K = out;
      gd.A = w.dyb; gd.lda = kOutPad;
      gd.B = w.wb + ly.off_w; gd.ldb = kOutPad;
      gd.C = G[cur]; gd.ldc = L; gd.alpha_dev = clip ? scale + nh : nullptr;
      P3D_TRY(tcg::gemm(gd, st));
    } else {
      Epilogue e1; e1.alpha_dev = clip ? scale + nh : nullptr;
      P3D_TRY(sgemm(false, true, B, L, out, w.dy, out, m->theta + ly.off_w, out, G[cur], L, e1, st));
    }
  }
  for (int li = nh - 1; li >= 0; --li) {
    const Layer& ly = m->layers[li];
    if (residual && li >= 2 && (li % 2) == 0) keepi = cur;   // d(h[li]) also flows to h[li-2]
    BwdArgs a;
    a.dh = G[cur]; a.z = w.z + li * bl; a.mean = w.mean + static_cast<size_t>(li) * L; a.rstd = w.rstd + static_cast<size_t>(li) * L;
    a.gamma = ly.has_bn ? m->theta + ly.off_gamma : nullptr; a.beta = ly.has_bn ? m->theta + ly.off_beta : nullptr;
    a.mask = maskbuf + li * bl; a.dz = w.dz; a.dzb = (tc && !ly.has_bn) ? w.dzb : nullptr; a.sums = w.red + 2ull * li * L; a.inv_keep = 1.f / keep;
    a.B = B; a.L = L; a.has_bn = ly.has_bn; a.dropout = dropout;
    bwd_act_kernel<<<colgrid(L, B), dim3(32, 8), 0, st>>>(a);
    P3D_LAUNCH_CHECK();
    if (ly.has_bn) {
      P3D_TRY(allreduce(m, a.sums, 2ull * L, ncclDouble, st));
      bwd_bn_kernel<<<egrid(static_cast<long long>(bl)), 256, 0, st>>>(w.dz, a.z, a.mean, a.rstd, a.gamma, a.sums, static_cast<float>(invBg), B, L,
                                                                         tc ? w.dzb : nullptr);
      P3D_LAUNCH_CHECK();
      // the sums are already global after the all-reduce, so every rank holds the full dgamma/dbeta;
      // pre-divide by world so that the flat gradient all-reduce (a sum) restores them exactly once.
      bn_param_grad_kernel<<<(L + 255) / 256, 256, 0, st>>>(a.sums, L, 1.0 / m->world, m->grad + ly.off_gamma, m->grad + ly.off_beta);
      P3D_LAUNCH_CHECK();
      // the bias feeding a BN layer has an exactly-zero gradient (it is removed by the mean subtraction)
    } else {
      colsum_kernel<<<colgrid(L, B), dim3(32, 8), 0, st>>>(w.dz, B, L, m->grad + ly.off_b);
      P3D_LAUNCH_CHECK();
    }
    const float* in = li == 0 ? x : w.h + (li - 1) * bl;
    const int lda = li == 0 ? kIn : L;
    if (tc) {
      GemmArgs gw;   // dW = in^T dz
      gw.M = ly.K; gw.N = L; gw.K = static_cast<int>(B);
      gw.A = li == 0 ? w.xb : w.hb + (li - 1) * bl; gw.lda = lda; gw.a_mn = 1;
      gw.B = w.dzb; gw.ldb = L; gw.b_mn = 1;
      gw.C = m->grad + ly.off_w; gw.ldc = L; gw.split_k = 1;
      P3D_TRY(tcg::gemm(gw, st));
    } else {
      Epilogue e0;
      P3D_TRY(sgemm(true, false, ly.K, L, static_cast<int>(B), in, lda, w.dz, L, m->grad + ly.off_w, L, e0, st));
    }
    if (li > 0) {
      int nxt = 0;
      while (nxt == cur || nxt == keepi) ++nxt;
      const bool add = residual && (li % 2) == 1 && keepi >= 0;
      if (tc) {
        GemmArgs gd;   // dh_prev = dz W^T * clip (+ the gradient that bypassed the block)
        gd.M = static_cast<int>(B); gd.N = L; gd.K = L;
        gd.A = w.dzb; gd.lda = L;
        gd.B = w.wb + ly.off_w; gd.ldb = L;
        gd.C = G[nxt]; gd.ldc = L; gd.alpha_dev = clip ? scale + li : nullptr;
        if (add) { gd.res = G[keepi]; gd.ldres = L; }
        P3D_TRY(tcg::gemm(gd, st));
      } else {
        Epilogue e1; e1.alpha_dev = clip ? scale + li : nullptr;
        if (add) e1.res = G[keepi];
        P3D_TRY(sgemm(false, true, B, L, L, w.dz, L, m->theta + ly.off_w, L, G[nxt], L, e1, st));
      }
      if (add) keepi = -1;
      cur = nxt;
    }
  }
  // ---------------------------------------------------------------- gradient exchange + update
  P3D_TRY(allreduce(m, m->grad, m->n_train, ncclFloat, st));
  if (clip) {
    for (int l = 0; l < nlay; ++l) {
      const Layer& ly = m->layers[l];
      const size_t n = static_cast<size_t>(ly.K) * ly.N;
      dot_kernel<<<egrid(static_cast<long long>(n / 4 + 1)), 256, 0, st>>>(m->theta + ly.off_w, m->grad + ly.off_w, n, dots + l);
      P3D_LAUNCH_CHECK();
      clip_grad_kernel<<<egrid(static_cast<long long>(n)), 256, 0, st>>>(m->grad + ly.off_w, m->theta + ly.off_w, n, dots + l, m->norm2 + l);
      P3D_LAUNCH_CHECK();
    }
  }
  // learning-rate schedule (src/linear_model.py:86-90) and TF's bias-corrected step size
  const double tstep = static_cast<double>(m->global_step);
  const float lr_t = m->cfg.learning_rate * powf(0.96f, static_cast<float>(tstep / 100000.0));
  const double tt = tstep + 1.0;
  const float alpha = static_cast<float>(static_cast<double>(lr_t) * std::sqrt(1.0 - std::pow(0.999, tt)) / (1.0 - std::pow(0.9, tt)));
  finish_step_scalars_kernel<<<1, 1, 0, st>>>(lossacc, static_cast<double>(Bg) * out, lr_t, loss, lr_used);
  P3D_LAUNCH_CHECK();
  adam_kernel<<<egrid(static_cast<long long>(m->n_train)), 256, 0, st>>>(m->theta, m->grad, m->adam_m, m->adam_v, m->n_train, alpha);
  P3D_LAUNCH_CHECK();
  m->global_step += 1;
  m->pack_valid = false;
  return P3D_OK;
}

}  // namespace train
}  // namespace p3d

using namespace p3d;

extern "C" {

int p3d_model_train_step(p3d_model* m, const float* x, const float* t, int64_t B, float keep_prob, uint64_t seed,
                         const uint8_t* mask_or_null, int64_t global_B, int64_t row0, float* loss, float* lr_used, float* y,
                         void* stream) {
  P3D_REQUIRE(m && x && t && y, "train_step: null argument");
  P3D_REQUIRE(B >= 1, "train_step: empty batch");
  P3D_REQUIRE(keep_prob > 0.f && keep_prob <= 1.f, "train_step: dropout_keep_prob must be in (0,1]");
  if (global_B <= 0) global_B = B;
  P3D_REQUIRE(m->world > 1 || global_B == B, "train_step: global_B != B without a communicator");
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  return train::train_step(m, x, t, B, keep_prob, seed, mask_or_null, global_B, row0, loss, lr_used, y, static_cast<cudaStream_t>(stream));
}

int p3d_nccl_unique_id(uint8_t* id_host) {
  P3D_REQUIRE(id_host, "nccl_unique_id: null argument");
  train::NcclApi* api = train::nccl();
  if (!api) return P3D_ERR_NCCL;
  ncclUniqueId id;
  P3D_NCCL(api->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id_host, &id, 128);
  return P3D_OK;
}

int p3d_model_attach_nccl(p3d_model* m, const uint8_t* id_host, int rank, int world) {
  P3D_REQUIRE(m && id_host && world >= 1 && rank >= 0 && rank < world, "attach_nccl: bad argument");
  train::NcclApi* api = train::nccl();
  if (!api) return P3D_ERR_NCCL;
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  ncclUniqueId id;
  memcpy(&id, id_host, 128);
  ncclComm_t comm;
  P3D_NCCL(api->CommInitRank(&comm, world, id, rank));
  m->nccl_comm = comm; m->rank = rank; m->world = world;
  return P3D_OK;
}

}  // extern "C"
