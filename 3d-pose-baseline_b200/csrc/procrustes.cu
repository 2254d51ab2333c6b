// Evaluation kernels (K5): un-normalise -> per-pose Procrustes alignment -> per-joint error (MPJPE).
//   predict_3dpose.evaluate_batches arithmetic   src/predict_3dpose.py:399-442
//   procrustes.compute_similarity_transform      src/procrustes.py:2-63
//
// Layout: a block stages 128 poses x 2 x 48 fp32 through shared memory with coalesced float4 loads,
// then every lane owns ONE pose (a warp = 32 poses): the serial 3x3 eigen-solve keeps all 32 lanes
// busy instead of one lane per warp.  Per-joint error sums are reduced with warp shuffles, kept in
// registers across the grid-stride loop and flushed with one fp64 atomic per joint per block.
// All alignment arithmetic is fp64 (the reference is NumPy float64; tolerance 1e-3 mm).
#include "common.cuh"
#include "math_hd.h"

namespace p3d {
namespace evalk {

constexpr int PB = 128;        // poses per block iteration
constexpr int MAXJ = 17;

struct EvalArgs {
  double mean[48], stdv[48];   // gathered to the used dims (hip excluded)
  double hip[3];               // un-normalised hip = mean3d[0:3] (ignored dims come back as the mean)
  int width;                   // 48 or 42
  int J;                       // joints in the error: 17 (hip + 16) or 14
  int with_hip;                // 1 unless predict_14
  int use_procrustes;
};

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(PB, 4) procrustes_mpjpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                              const __grid_constant__ EvalArgs a, float* __restrict__ dists,
                                                              double* __restrict__ joint_sum, long long N) {
  extern __shared__ float sm[];
  const int W = a.width, WP = a.width + 1;      // padded row pitch: conflict-free per-lane rows
  float* sp = sm;
  float* sg = sm + PB * WP;
  __shared__ double sacc[PB / 32][MAXJ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double jacc[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) jacc[j] = 0.0;
  const long long ntiles = (N + PB - 1) / PB;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * PB;
    const int np = static_cast<int>((N - p0 < PB) ? (N - p0) : PB);
    const int tot = np * W;
    const float* gp = pred + p0 * W;
    const float* gg = gt + p0 * W;
    if ((W & 3) == 0) {
      for (int i = tid; i < tot / 4; i += PB) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(gp) + i);
        const float4 u = __ldg(reinterpret_cast<const float4*>(gg) + i);
        const int e = i * 4, r = e / W, c = e - r * W;
        float* d = sp + r * WP + c; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        float* f = sg + r * WP + c; f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
      }
    } else {   // width 42: rows are only 8-byte aligned
      for (int i = tid; i < tot / 2; i += PB) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(gp) + i);
        const float2 u = __ldg(reinterpret_cast<const float2*>(gg) + i);
        const int e = i * 2, r = e / W, c = e - r * W;
        sp[r * WP + c] = v.x; sp[r * WP + c + 1] = v.y;
        sg[r * WP + c] = u.x; sg[r * WP + c + 1] = u.y;
      }
    }
    __syncthreads();
    const bool live = tid < np;
    const float* mp = sp + tid * WP;
    const float* mg = sg + tid * WP;
    const int J = a.J, J0 = a.with_hip;          // joints J0..J-1 come from the arrays, joint 0 is the hip
    double dj[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) dj[j] = 0.0;
    if (live) {
      double T[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double b = 1.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
      if (a.use_procrustes) {
        // means (X = ground truth, Y = prediction)
        double mx[3] = {0, 0, 0}, my[3] = {0, 0, 0};
        if (J0) { for (int d = 0; d < 3; ++d) { mx[d] = a.hip[d]; my[d] = a.hip[d]; } }
        for (int k = 0; k < W; k += 3) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            mx[d] += static_cast<double>(mg[k + d]) * a.stdv[k + d] + a.mean[k + d];
            my[d] += static_cast<double>(mp[k + d]) * a.stdv[k + d] + a.mean[k + d];
          }
        }
        const double invJ = 1.0 / J;
        for (int d = 0; d < 3; ++d) { mx[d] *= invJ; my[d] *= invJ; }
        double ssx = 0, ssy = 0, A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < J; ++j) {
          double x[3], y[3];
          if (J0 && j == 0) {
            for (int d = 0; d < 3; ++d) { x[d] = a.hip[d] - mx[d]; y[d] = a.hip[d] - my[d]; }
          } else {
            const int k = (j - J0) * 3;
            for (int d = 0; d < 3; ++d) {
              x[d] = static_cast<double>(mg[k + d]) * a.stdv[k + d] + a.mean[k + d] - mx[d];
              y[d] = static_cast<double>(mp[k + d]) * a.stdv[k + d] + a.mean[k + d] - my[d];
            }
          }
          for (int d = 0; d < 3; ++d) { ssx += x[d] * x[d]; ssy += y[d] * y[d]; }
          for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) A[r * 3 + s] += x[r] * y[s];
        }
        const double normX = sqrt(ssx), normY = sqrt(ssy);
        const double inv = 1.0 / (normX * normY);
        for (int i = 0; i < 9; ++i) A[i] *= inv;          // A = X0^T Y0 of the unit-norm point sets
        double tr;
        kabsch_rotation(A, T, tr);
        b = tr * normX / normY;                            // compute_optimal_scale=True (procrustes.py:52-55)
        c0 = mx[0] - b * (my[0] * T[0] + my[1] * T[3] + my[2] * T[6]);
        c1 = mx[1] - b * (my[0] * T[1] + my[1] * T[4] + my[2] * T[7]);
        c2 = mx[2] - b * (my[0] * T[2] + my[1] * T[5] + my[2] * T[8]);
      }
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        if (j < J) {
          double x[3], y[3];
          if (J0 && j == 0) {
            for (int d = 0; d < 3; ++d) { x[d] = a.hip[d]; y[d] = a.hip[d]; }
          } else {
            const int k = (j - J0) * 3;
            for (int d = 0; d < 3; ++d) {
              x[d] = static_cast<double>(mg[k + d]) * a.stdv[k + d] + a.mean[k + d];
              y[d] = static_cast<double>(mp[k + d]) * a.stdv[k + d] + a.mean[k + d];
            }
          }
          double o0 = y[0], o1 = y[1], o2 = y[2];
          if (a.use_procrustes) {     // out = b * out.dot(T) + c  (predict_3dpose.py:419)
            o0 = b * (y[0] * T[0] + y[1] * T[3] + y[2] * T[6]) + c0;
            o1 = b * (y[0] * T[1] + y[1] * T[4] + y[2] * T[7]) + c1;
            o2 = b * (y[0] * T[2] + y[1] * T[5] + y[2] * T[8]) + c2;
          }
          const double e0 = o0 - x[0], e1 = o1 - x[1], e2 = o2 - x[2];
          // the squared error is fp64; its root in fp32 (rel. 6e-8, i.e. < 1e-5 mm) avoids 17 software fp64 sqrt per pose
          dj[j] = static_cast<double>(sqrtf(static_cast<float>(e0 * e0 + e1 * e1 + e2 * e2)));
        } else {
          dj[j] = 0.0;
        }
      }
      if (dists) {
        float* dp = dists + (p0 + tid) * J;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) if (j < J) dp[j] = static_cast<float>(dj[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const double v = (live && j < J) ? dj[j] : 0.0;
      jacc[j] += v;
    }
    __syncthreads();
  }
  // block reduction: shuffles within the warp, shared memory across warps, one atomic per joint
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const double s = warp_sum(jacc[j]);
    if (lane == 0) sacc[warp][j] = s;
  }
  __syncthreads();
  if (tid < a.J) {
    double s = 0;
    for (int w = 0; w < PB / 32; ++w) s += sacc[w][tid];
    atomicAdd(joint_sum + tid, s);
  }
}

// Batched procrustes.compute_similarity_transform on raw fp64 poses; one thread per pose.
__global__ void similarity_transform_kernel(const double* __restrict__ X, const double* __restrict__ Y, int J, int scale,
                                            long long N, double* d_out, double* Z, double* T_out, double* b_out, double* c_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double* x = X + i * J * 3;
    const double* y = Y + i * J * 3;
    double mx[3] = {0, 0, 0}, my[3] = {0, 0, 0};
    for (int j = 0; j < J; ++j)
      for (int d = 0; d < 3; ++d) { mx[d] += x[j * 3 + d]; my[d] += y[j * 3 + d]; }
    for (int d = 0; d < 3; ++d) { mx[d] /= J; my[d] /= J; }
    double ssx = 0, ssy = 0, A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < J; ++j) {
      double a[3], b[3];
      for (int d = 0; d < 3; ++d) { a[d] = x[j * 3 + d] - mx[d]; b[d] = y[j * 3 + d] - my[d]; ssx += a[d] * a[d]; ssy += b[d] * b[d]; }
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) A[r * 3 + s] += a[r] * b[s];
    }
    const double normX = sqrt(ssx), normY = sqrt(ssy);
    const double inv = 1.0 / (normX * normY);
    for (int k = 0; k < 9; ++k) A[k] *= inv;
    double T[9], tr;
    kabsch_rotation(A, T, tr);
    double b, dd, zs;
    if (scale) { b = tr * normX / normY; dd = 1.0 - tr * tr; zs = normX * tr; }
    else { b = 1.0; dd = 1.0 + ssy / ssx - 2.0 * tr * normY / normX; zs = normY; }
    if (d_out) d_out[i] = dd;
    if (b_out) b_out[i] = b;
    if (T_out) for (int k = 0; k < 9; ++k) T_out[i * 9 + k] = T[k];
    if (c_out)
      for (int s = 0; s < 3; ++s) c_out[i * 3 + s] = mx[s] - b * (my[0] * T[s] + my[1] * T[3 + s] + my[2] * T[6 + s]);
    if (Z)
      for (int j = 0; j < J; ++j) {
        const double y0 = (y[j * 3] - my[0]) / normY, y1 = (y[j * 3 + 1] - my[1]) / normY, y2 = (y[j * 3 + 2] - my[2]) / normY;
        for (int s = 0; s < 3; ++s) Z[(i * J + j) * 3 + s] = zs * (y0 * T[s] + y1 * T[3 + s] + y2 * T[6 + s]) + mx[s];
      }
  }
}

}  // namespace evalk
}  // namespace p3d

using namespace p3d;
using namespace p3d::evalk;

extern "C" {

int p3d_procrustes_mpjpe(const float* pred_n, const float* gt_n, const double* mean3d, const double* std3d, int predict_14,
                         int use_procrustes, int64_t N, float* dists, double* joint_sum, void* stream) {
  P3D_REQUIRE(pred_n && gt_n && mean3d && std3d && joint_sum && N >= 0, "procrustes_mpjpe: null argument");
  if (N == 0) return P3D_OK;
  static const int j16[16] = {1, 2, 3, 6, 7, 8, 12, 13, 14, 15, 17, 18, 19, 25, 26, 27};
  static const int j14[14] = {1, 2, 3, 6, 7, 8, 13, 15, 17, 18, 19, 25, 26, 27};
  EvalArgs a;
  memset(&a, 0, sizeof(a));
  const int nj = predict_14 ? 14 : 16;
  const int* jt = predict_14 ? j14 : j16;
  for (int j = 0; j < nj; ++j)
    for (int d = 0; d < 3; ++d) { a.mean[j * 3 + d] = mean3d[jt[j] * 3 + d]; a.stdv[j * 3 + d] = std3d[jt[j] * 3 + d]; }
  for (int d = 0; d < 3; ++d) a.hip[d] = mean3d[d];
  a.width = nj * 3;
  a.with_hip = predict_14 ? 0 : 1;      // dtu3d = [0,1,2] U dim_to_use_3d unless predict_14 (predict_3dpose.py:405)
  a.J = nj + a.with_hip;
  a.use_procrustes = use_procrustes ? 1 : 0;
  const size_t smem = sizeof(float) * 2 * PB * (a.width + 1);
  static bool attr = false;
  if (!attr) {
    P3D_CUDA(cudaFuncSetAttribute(procrustes_mpjpe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  const long long ntiles = (N + PB - 1) / PB;
  const int grid = ntiles < 148 * 4 ? (int)ntiles : 148 * 4;
  procrustes_mpjpe_kernel<<<grid, PB, smem, (cudaStream_t)stream>>>(pred_n, gt_n, a, dists, joint_sum, N);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_similarity_transform_f64(const double* X, const double* Y, int J, int compute_optimal_scale, int64_t N, double* d,
                                 double* Z, double* T, double* b, double* c, void* stream) {
  P3D_REQUIRE(X && Y && J >= 1 && N >= 0, "similarity_transform: bad argument");
  if (N == 0) return P3D_OK;
  long long g = (N + 127) / 128;
  if (g > 148 * 8) g = 148 * 8;
  similarity_transform_kernel<<<(int)g, 128, 0, (cudaStream_t)stream>>>(X, Y, J, compute_optimal_scale, N, d, Z, T, b, c);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

}  // extern "C"
