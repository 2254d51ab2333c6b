// Host/device math shared by the geometry and evaluation kernels.  Header-only and free of CUDA
// types so that tests/hostcheck can compile the very same code with g++ and compare it with the
// oracle on the CPU (test infrastructure only - the product always runs these on the GPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define P3D_HD __host__ __device__ __forceinline__
#else
#define P3D_HD inline
#endif

namespace p3d {

template <typename T>
struct CamT {
  T R[9], Tr[3], f[2], c[2], k[3], p[2];
};

// cameras.world_to_camera_frame (src/cameras.py:55-72): X = R (P - T)
template <typename T>
P3D_HD void world_to_cam(const CamT<T>& cam, T px, T py, T pz, T& X0, T& X1, T& X2) {
  const T dx = px - cam.Tr[0], dy = py - cam.Tr[1], dz = pz - cam.Tr[2];
  X0 = cam.R[0] * dx + cam.R[1] * dy + cam.R[2] * dz;
  X1 = cam.R[3] * dx + cam.R[4] * dy + cam.R[5] * dz;
  X2 = cam.R[6] * dx + cam.R[7] * dy + cam.R[8] * dz;
}

// cameras.camera_to_world_frame (src/cameras.py:74-90): P = R^T X + T
template <typename T>
P3D_HD void cam_to_world(const CamT<T>& cam, T x, T y, T z, T& P0, T& P1, T& P2) {
  P0 = cam.R[0] * x + cam.R[3] * y + cam.R[6] * z + cam.Tr[0];
  P1 = cam.R[1] * x + cam.R[4] * y + cam.R[7] * z + cam.Tr[1];
  P2 = cam.R[2] * x + cam.R[5] * y + cam.R[8] * z + cam.Tr[2];
}

// cameras.project_point_radial (src/cameras.py:39-51) for one point.
// tan pairs p[0] with y and p[1] with x; the additive term is [p[1]; p[0]] * r2 (:44-46).
template <typename T>
P3D_HD void project_point(const CamT<T>& cam, T px, T py, T pz, T& u, T& v, T& D, T& radial, T& tang, T& r2) {
  T X0, X1, X2;
  world_to_cam(cam, px, py, pz, X0, X1, X2);
  const T x = X0 / X2, y = X1 / X2;
  r2 = x * x + y * y;
  radial = T(1) + cam.k[0] * r2 + cam.k[1] * (r2 * r2) + cam.k[2] * (r2 * r2 * r2);
  tang = cam.p[0] * y + cam.p[1] * x;
  const T s = radial + tang;
  u = cam.f[0] * (x * s + cam.p[1] * r2) + cam.c[0];
  v = cam.f[1] * (y * s + cam.p[0] * r2) + cam.c[1];
  D = X2;
}

// ---------------------------------------------------------------------------------------------
// Optimal rotation of procrustes.compute_similarity_transform (src/procrustes.py:38-50).
// Given A = X0^T Y0 (3x3, row-major, any positive scaling), returns T = V diag(1,1,det) U^T where
// A = U S V^T, i.e. the reflection-corrected rotation the reference builds (:45-48), and
// tr = s0 + s1 + det*s2 = trace(T A) (the reference's traceTA before normalisation by |X0||Y0|).
//
// Method: cyclic Jacobi eigen-decomposition of the symmetric A^T A -> V; U's first two columns from
// A v_i / s_i (Gram-Schmidt), third columns as cross products, which yields exactly
// V diag(1,1,sign det(VU^T)) U^T without ever forming a possibly ill-defined third singular vector.
P3D_HD void jacobi_rot(double app, double aqq, double apq, double& c, double& s) {
  if (apq == 0.0) { c = 1.0; s = 0.0; return; }
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  c = 1.0 / sqrt(t * t + 1.0);
  s = t * c;
}
P3D_HD void jacobi_rot_f(float app, float aqq, float apq, float& c, float& s) {
  if (apq == 0.0f) { c = 1.0f; s = 0.0f; return; }
  const float theta = (aqq - app) / (2.0f * apq);
  const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
  c = 1.0f / sqrtf(t * t + 1.0f);
  s = t * c;
}

#if defined(__CUDA_ARCH__)
#define P3D_UNROLL _Pragma("unroll")
#else
#define P3D_UNROLL
#endif

// one cyclic Jacobi sweep (rotations (0,1), (0,2), (1,2)) on symmetric S, accumulating V <- V J
template <typename T, typename ROT>
P3D_HD void jacobi_sweep(T S[3][3], T V[3][3], ROT rot) {
  P3D_UNROLL
  for (int r = 0; r < 3; ++r) {
    const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;
    T c, s;
    rot(S[p][p], S[q][q], S[p][q], c, s);
    P3D_UNROLL
    for (int k = 0; k < 3; ++k) {
      const T skp = S[k][p], skq = S[k][q];
      S[k][p] = c * skp - s * skq;
      S[k][q] = s * skp + c * skq;
    }
    P3D_UNROLL
    for (int k = 0; k < 3; ++k) {
      const T spk = S[p][k], sqk = S[q][k];
      S[p][k] = c * spk - s * sqk;
      S[q][k] = s * spk + c * sqk;
    }
    P3D_UNROLL
    for (int k = 0; k < 3; ++k) {
      const T vkp = V[k][p], vkq = V[k][q];
      V[k][p] = c * vkp - s * vkq;
      V[k][q] = s * vkp + c * vkq;
    }
  }
}

P3D_HD void kabsch_rotation(const double A[9], double T[9], double& tr) {
  // S = A^T A (symmetric, fp64)
  double S[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i][j] = A[0 * 3 + i] * A[0 * 3 + j] + A[1 * 3 + i] * A[1 * 3 + j] + A[2 * 3 + i] * A[2 * 3 + j];
  // Stage 1 (cheap): approximate eigenvectors by 5 Jacobi sweeps in fp32 (fp64 sqrt/div are ~30-instruction
  // software sequences on the GPU; doing all ~18 rotations in fp64 made the evaluation kernel compute bound)
  float Sf[3][3], Vf[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Sf[i][j] = static_cast<float>(S[i][j]);
  for (int sweep = 0; sweep < 5; ++sweep) {
    const float off = fabsf(Sf[0][1]) + fabsf(Sf[0][2]) + fabsf(Sf[1][2]);
    const float diag = fabsf(Sf[0][0]) + fabsf(Sf[1][1]) + fabsf(Sf[2][2]);
    if (off <= 1e-9f * diag) break;
    jacobi_sweep<float>(Sf, Vf, [](float a, float b, float c, float& cc, float& ss) { jacobi_rot_f(a, b, c, cc, ss); });
  }
  // Stage 2: re-orthonormalise the fp32 basis in fp64 (Gram-Schmidt), rotate S into it, and polish with
  // fp64 Jacobi sweeps (quadratic convergence: one sweep takes the ~1e-7 residual below 1e-13)
  double V[3][3];
  {
    double a0[3] = {Vf[0][0], Vf[1][0], Vf[2][0]}, a1[3] = {Vf[0][1], Vf[1][1], Vf[2][1]};
    const double n0 = 1.0 / sqrt(a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2]);
    for (int i = 0; i < 3; ++i) a0[i] *= n0;
    const double d01 = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2];
    for (int i = 0; i < 3; ++i) a1[i] -= d01 * a0[i];
    const double n1 = 1.0 / sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    for (int i = 0; i < 3; ++i) a1[i] *= n1;
    const double a2[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
    for (int i = 0; i < 3; ++i) { V[i][0] = a0[i]; V[i][1] = a1[i]; V[i][2] = a2[i]; }
  }
  {
    double SV[3][3], S1[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) SV[i][j] = S[i][0] * V[0][j] + S[i][1] * V[1][j] + S[i][2] * V[2][j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S1[i][j] = V[0][i] * SV[0][j] + V[1][i] * SV[1][j] + V[2][i] * SV[2][j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[i][j] = (i <= j) ? S1[i][j] : S1[j][i];   // keep exactly symmetric
  }
  for (int sweep = 0; sweep < 6; ++sweep) {
    const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
    const double diag = fabs(S[0][0]) + fabs(S[1][1]) + fabs(S[2][2]);
    if (off <= 1e-17 * diag) break;
    jacobi_sweep<double>(S, V, [](double a, double b, double c, double& cc, double& ss) { jacobi_rot(a, b, c, cc, ss); });
  }
  // two largest eigenpairs by compare-and-swap of (eigenvalue, column) with constant indices
  double l0 = S[0][0], l1 = S[1][1], l2 = S[2][2];
#define P3D_CSWAP(la, lb, ca, cb)                                        \
  if (la < lb) {                                                         \
    double t_ = la; la = lb; lb = t_;                                    \
    for (int k_ = 0; k_ < 3; ++k_) { t_ = V[k_][ca]; V[k_][ca] = V[k_][cb]; V[k_][cb] = t_; } \
  }
  P3D_CSWAP(l0, l1, 0, 1)
  P3D_CSWAP(l0, l2, 0, 2)
  P3D_CSWAP(l1, l2, 1, 2)
#undef P3D_CSWAP
  (void)l0; (void)l1; (void)l2;
  double v1[3] = {V[0][0], V[1][0], V[2][0]};
  double v2[3] = {V[0][1], V[1][1], V[2][1]};
  double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  // u1 = A v1 / |A v1| ; u2 = A v2 orthogonalised against u1 ; u3 = u1 x u2
  double u1[3], u2[3];
  for (int i = 0; i < 3; ++i) {
    u1[i] = A[i * 3 + 0] * v1[0] + A[i * 3 + 1] * v1[1] + A[i * 3 + 2] * v1[2];
    u2[i] = A[i * 3 + 0] * v2[0] + A[i * 3 + 1] * v2[1] + A[i * 3 + 2] * v2[2];
  }
  const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
  const double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  for (int i = 0; i < 3; ++i) u2[i] /= n2;
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  // T = V' U'^T : T[i][j] = v1[i]u1[j] + v2[i]u2[j] + v3[i]u3[j]
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
  // tr = trace(T A)
  tr = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tr += T[i * 3 + j] * A[j * 3 + i];
}

// ---------------------------------------------------------------------------------------------
// fp32 evaluation path (the HBM-bound MPJPE kernel): everything a pose needs in single precision.
//
// Rotation: ONE-SIDED (Hestenes) Jacobi directly on A = U S V^T - column pairs of B = A V are rotated
// until orthogonal, so B = U S and V are obtained without ever squaring the condition number (A^T A in
// fp32 would cost half the digits of the small singular directions).  The two dominant pairs define
// T = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T = V diag(1,1,det) U^T of procrustes.py:45-48.
P3D_HD float p3d_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  const float r = rsqrtf(x);
  return r * (1.5f - 0.5f * x * r * r);      // one Newton step: full fp32 accuracy
#else
  return 1.0f / sqrtf(x);
#endif
}
P3D_HD float p3d_rsqrt_fast(float x) {          // MUFU.RSQ (~2 ulp), x > 0
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
P3D_HD float p3d_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}

P3D_HD float p3d_rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  return __fdividef(1.0f, x);
#else
  return 1.0f / x;
#endif
}
P3D_HD float p3d_sqrt_fast(float x) {          // sqrt.approx: ~1 ulp, no slow path (x >= 0)
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

P3D_HD void kabsch_rotation_f32(const float A[9], float T[9], float& tr) {
  float B[3][3], V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
  P3D_UNROLL
  for (int i = 0; i < 3; ++i) { B[i][0] = A[i * 3]; B[i][1] = A[i * 3 + 1]; B[i][2] = A[i * 3 + 2]; }
  // Three sweeps reach fp32 accuracy on every test set (random, mirrored, nearly planar, near-identity); a fourth
  // runs only for a pose whose third sweep still rotated by more than ~1e-3 rad (quadratic convergence).
  bool again = true;
  P3D_UNROLL
  for (int sweep = 0; sweep < 4; ++sweep) {
    if (sweep == 3 && !again) break;
    float tmax = 0.f;
    P3D_UNROLL
    for (int r = 0; r < 3; ++r) {
      const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;
      const float al = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
      const float be = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
      const float ga = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
      float c = 1.f, s = 0.f;
      if (ga * ga > 1e-16f * al * be) {          // already orthogonal to fp32 accuracy otherwise
        // Rotation by theta with tan 2 theta = 2 ga / (be - al), |theta| <= pi/4, from the double-angle form:
        //   cos 2theta = |a| / r, sin 2theta = sign(a) b / r  (a = be - al, b = 2 ga, r = hypot(a, b)),
        //   c = sqrt((1 + cos 2theta) / 2), s = sin 2theta / (2 c).
        // Two dependent MUFU.RSQ instead of the four special-function steps of the tangent form (rcp, sqrt, rcp,
        // rsqrt): the kernel is bound by the latency of this serial chain (12 rotations per pose), not by issue slots.
        // (c, s) is a rotation times (1 + O(1e-7)): a uniform scaling of the (p, q) plane, applied to B and V alike,
        // which the normalisations after the sweeps remove.
        const float a_ = be - al, b_ = ga + ga;
        const float rinv = p3d_rsqrt_fast(a_ * a_ + b_ * b_);
        const float c2 = 0.5f * fabsf(a_) * rinv + 0.5f;
        const float rs = p3d_rsqrt_fast(c2);
        c = c2 * rs;
        s = (a_ >= 0.f ? b_ : -b_) * (0.5f * rinv) * rs;
        tmax = fmaxf(tmax, fabsf(s));
      }
      P3D_UNROLL
      for (int k = 0; k < 3; ++k) {
        const float bp = B[k][p], bq = B[k][q];
        B[k][p] = c * bp - s * bq; B[k][q] = s * bp + c * bq;
        const float vp = V[k][p], vq = V[k][q];
        V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
      }
    }
    again = tmax > 1e-3f;
  }
  float l0 = B[0][0] * B[0][0] + B[1][0] * B[1][0] + B[2][0] * B[2][0];
  float l1 = B[0][1] * B[0][1] + B[1][1] * B[1][1] + B[2][1] * B[2][1];
  float l2 = B[0][2] * B[0][2] + B[1][2] * B[1][2] + B[2][2] * B[2][2];
#define P3D_CSWAPF(la, lb, ca, cb)                                        \
  if (la < lb) {                                                          \
    float t_ = la; la = lb; lb = t_;                                      \
    for (int k_ = 0; k_ < 3; ++k_) { t_ = V[k_][ca]; V[k_][ca] = V[k_][cb]; V[k_][cb] = t_; \
                                     t_ = B[k_][ca]; B[k_][ca] = B[k_][cb]; B[k_][cb] = t_; } \
  }
  P3D_CSWAPF(l0, l1, 0, 1)
  P3D_CSWAPF(l0, l2, 0, 2)
  P3D_CSWAPF(l1, l2, 1, 2)
#undef P3D_CSWAPF
  (void)l2;
  float v1[3] = {V[0][0], V[1][0], V[2][0]}, v2[3] = {V[0][1], V[1][1], V[2][1]};
  float u1[3] = {B[0][0], B[1][0], B[2][0]}, u2[3] = {B[0][1], B[1][1], B[2][1]};
  {
    const float n = p3d_rsqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    for (int i = 0; i < 3; ++i) v1[i] *= n;
    const float d = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
    for (int i = 0; i < 3; ++i) v2[i] -= d * v1[i];
    const float n2 = p3d_rsqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    for (int i = 0; i < 3; ++i) v2[i] *= n2;
  }
  {
    const float n = p3d_rsqrt(l0);
    for (int i = 0; i < 3; ++i) u1[i] *= n;
    const float d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
    const float n2 = p3d_rsqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    for (int i = 0; i < 3; ++i) u2[i] *= n2;
  }
  const float v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  const float u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  tr = 0.f;
  P3D_UNROLL
  for (int i = 0; i < 3; ++i)
    P3D_UNROLL
    for (int j = 0; j < 3; ++j) {
      T[i * 3 + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
      tr += T[i * 3 + j] * A[j * 3 + i];
    }
}

// Per-pose errors of predict_3dpose.evaluate_batches (src/predict_3dpose.py:399-442) in fp32.
//   W, J0        : row width (48 or 42) and 1 when the implicit hip joint is part of the error (W = 48)
//   g[k], p[k]   : normalised ground truth / prediction, k in [0, W) - OVERWRITTEN with the centred un-normalised
//                  coordinates (the arrays live in registers in the kernel; pass 3 reuses them)
//   sd[k]        : data_std_3d gathered to the used dims
//   mc[k]        : data_mean_3d gathered, minus the mean over the J joints of the un-normalised MEAN pose
//                  (so that centred coordinates come out of one fma and stay small)
//   hipc[3]      : same for the hip joint (un-normalised hip = data_mean_3d[0:3] for both poses)
// Centred joint: x_c = g*sd + mc - (sum_k g_k sd_k)/J.   Aligned error: b (y_c T) - x_c  (= b y T + c - x).
template <int W, int J0>
P3D_HD void pose_errors_f32(float (&g)[W], float (&p)[W], const float* sd, const float* mc, const float* hipc, int use_procrustes, float* dj) {
  constexpr int J = W / 3 + J0;
  if (!use_procrustes) {
    if (J0) dj[0] = 0.f;
    P3D_UNROLL
    for (int j = J0; j < J; ++j) {
      const int k = (j - J0) * 3;
      const float e0 = (p[k] - g[k]) * sd[k], e1 = (p[k + 1] - g[k + 1]) * sd[k + 1], e2 = (p[k + 2] - g[k + 2]) * sd[k + 2];
      dj[j] = p3d_sqrt_fast(e0 * e0 + e1 * e1 + e2 * e2);
    }
    return;
  }
  float sx[3] = {0.f, 0.f, 0.f}, sy[3] = {0.f, 0.f, 0.f};
  P3D_UNROLL
  for (int k = 0; k < W; k += 3) {
    P3D_UNROLL
    for (int d = 0; d < 3; ++d) { sx[d] += g[k + d] * sd[k + d]; sy[d] += p[k + d] * sd[k + d]; }
  }
  const float invJ = 1.0f / static_cast<float>(J);
  P3D_UNROLL
  for (int d = 0; d < 3; ++d) { sx[d] *= invJ; sy[d] *= invJ; }
  float ssx = 0.f, ssy = 0.f, A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float hx[3] = {0.f, 0.f, 0.f}, hy[3] = {0.f, 0.f, 0.f};
  if (J0) {
    P3D_UNROLL
    for (int d = 0; d < 3; ++d) { hx[d] = hipc[d] - sx[d]; hy[d] = hipc[d] - sy[d]; ssx += hx[d] * hx[d]; ssy += hy[d] * hy[d]; }
    P3D_UNROLL
    for (int r = 0; r < 3; ++r)
      P3D_UNROLL
      for (int s = 0; s < 3; ++s) A[r * 3 + s] += hx[r] * hy[s];
  }
  P3D_UNROLL
  for (int k = 0; k < W; k += 3) {
    P3D_UNROLL
    for (int d = 0; d < 3; ++d) {
      g[k + d] = (g[k + d] * sd[k + d] + mc[k + d]) - sx[d];
      p[k + d] = (p[k + d] * sd[k + d] + mc[k + d]) - sy[d];
      ssx += g[k + d] * g[k + d]; ssy += p[k + d] * p[k + d];
    }
    P3D_UNROLL
    for (int r = 0; r < 3; ++r)
      P3D_UNROLL
      for (int s = 0; s < 3; ++s) A[r * 3 + s] += g[k + r] * p[k + s];
  }
  const float inv = p3d_rsqrt(ssx * ssy);
  P3D_UNROLL
  for (int i = 0; i < 9; ++i) A[i] *= inv;
  float T[9], tr;
  kabsch_rotation_f32(A, T, tr);
  const float b = tr * p3d_sqrt_fast(ssx * p3d_rcp(ssy));          // compute_optimal_scale=True (procrustes.py:52-55)
  P3D_UNROLL
  for (int i = 0; i < 9; ++i) T[i] *= b;
  // e = y_c (bT) - x_c with the subtraction folded into the first multiply-add: three FFMA per coordinate
  if (J0) {
    const float e0 = fmaf(hy[2], T[6], fmaf(hy[1], T[3], fmaf(hy[0], T[0], -hx[0])));
    const float e1 = fmaf(hy[2], T[7], fmaf(hy[1], T[4], fmaf(hy[0], T[1], -hx[1])));
    const float e2 = fmaf(hy[2], T[8], fmaf(hy[1], T[5], fmaf(hy[0], T[2], -hx[2])));
    dj[0] = p3d_sqrt_fast(fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
  }
  P3D_UNROLL
  for (int j = J0; j < J; ++j) {
    const int k = (j - J0) * 3;
    const float e0 = fmaf(p[k + 2], T[6], fmaf(p[k + 1], T[3], fmaf(p[k], T[0], -g[k])));
    const float e1 = fmaf(p[k + 2], T[7], fmaf(p[k + 1], T[4], fmaf(p[k], T[1], -g[k + 1])));
    const float e2 = fmaf(p[k + 2], T[8], fmaf(p[k + 1], T[5], fmaf(p[k], T[2], -g[k + 2])));
    dj[j] = p3d_sqrt_fast(fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
  }
}

// ---------------------------------------------------------------------------------------------
// Realtime front-end (src/openpose_3dpose_sandbox_realtime.py:20,137-163).
// kp36 = the first 18 OpenPose/COCO keypoints as (x, y) pairs.  The reference scatters keypoint i into H3.6M joint
// order[i] (order = [15,12,25,26,27,17,18,19,1,2,3,6,7,8], :20,:144-147), then synthesises
//   Hip       (joint 0)  = (RHip(1) + LHip(6)) / 2                      (:150)
//   Neck/Nose (joint 14) = (Head(15) + Spine(12)) / 2                   (:152)
//   Thorax    (joint 13) = 2 * Spine(12) - Neck/Nose(14)                (:154)
// This returns coordinate d (= joint * 2 + xy) of that 64-vector; joints the scatter never writes are 0 (the
// reference's first-frame value; none of them is in dim_to_use_2d).
P3D_HD int openpose_index_of_h36m_joint(int j) {
  switch (j) {
    case 15: return 0;  case 12: return 1;  case 25: return 2;  case 26: return 3;  case 27: return 4;
    case 17: return 5;  case 18: return 6;  case 19: return 7;  case 1:  return 8;  case 2:  return 9;
    case 3:  return 10; case 6:  return 11; case 7:  return 12; case 8:  return 13;
    default: return -1;
  }
}
P3D_HD double openpose_h36m_coord(const double* kp36, int d) {
  const int j = d >> 1, c = d & 1;
  if (j == 0) return (kp36[8 * 2 + c] + kp36[11 * 2 + c]) / 2;
  if (j == 14) return (kp36[0 * 2 + c] + kp36[1 * 2 + c]) / 2;
  if (j == 13) return 2 * kp36[1 * 2 + c] - (kp36[0 * 2 + c] + kp36[1 * 2 + c]) / 2;
  const int i = openpose_index_of_h36m_joint(j);
  return i >= 0 ? kp36[i * 2 + c] : 0.0;
}

}  // namespace p3d
