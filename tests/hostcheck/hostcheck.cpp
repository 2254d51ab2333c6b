// TEST INFRASTRUCTURE: compiles the host/device math header of the CUDA kernels with g++ so that the
// exact device arithmetic (projection, Kabsch/Jacobi rotation) can be checked against the oracle on a
// machine without a GPU.  Never linked into the product.
#include "../../3d-pose-baseline_b200/csrc/math_hd.h"

extern "C" {

void hc_kabsch(const double* A, double* T, double* tr, int n) {
  for (int i = 0; i < n; ++i) p3d::kabsch_rotation(A + 9 * i, T + 9 * i, tr[i]);
}

// cam = R[9] T[3] f[2] c[2] k[3] p[2] ; out = u v D radial tan r2 per point
void hc_project_f64(const double* P, const double* cam, double* out, int n) {
  p3d::CamT<double> c;
  for (int i = 0; i < 9; ++i) c.R[i] = cam[i];
  for (int i = 0; i < 3; ++i) { c.Tr[i] = cam[9 + i]; c.k[i] = cam[16 + i]; }
  for (int i = 0; i < 2; ++i) { c.f[i] = cam[12 + i]; c.c[i] = cam[14 + i]; c.p[i] = cam[19 + i]; }
  for (int i = 0; i < n; ++i)
    p3d::project_point(c, P[3 * i], P[3 * i + 1], P[3 * i + 2], out[6 * i], out[6 * i + 1], out[6 * i + 2], out[6 * i + 3],
                       out[6 * i + 4], out[6 * i + 5]);
}

void hc_project_f32(const float* P, const double* cam, float* out, int n) {
  p3d::CamT<float> c;
  for (int i = 0; i < 9; ++i) c.R[i] = (float)cam[i];
  for (int i = 0; i < 3; ++i) { c.Tr[i] = (float)cam[9 + i]; c.k[i] = (float)cam[16 + i]; }
  for (int i = 0; i < 2; ++i) { c.f[i] = (float)cam[12 + i]; c.c[i] = (float)cam[14 + i]; c.p[i] = (float)cam[19 + i]; }
  for (int i = 0; i < n; ++i)
    p3d::project_point(c, P[3 * i], P[3 * i + 1], P[3 * i + 2], out[6 * i], out[6 * i + 1], out[6 * i + 2], out[6 * i + 3],
                       out[6 * i + 4], out[6 * i + 5]);
}

void hc_kabsch_f32(const float* A, float* T, float* tr, int n) {
  for (int i = 0; i < n; ++i) p3d::kabsch_rotation_f32(A + 9 * i, T + 9 * i, tr[i]);
}

// fp32 evaluation math: dists[n][J] for normalised gt/pred rows [n][W]  (W = 48 with the hip joint, or 42)
void hc_pose_errors_f32(const float* gt, const float* pred, const float* sd, const float* mc, const float* hipc, int W,
                        int use_procrustes, float* dists, int n) {
  const int J = (W == 48) ? 17 : 14;
  for (int i = 0; i < n; ++i) {
    float g[48], p[48];
    for (int k = 0; k < W; ++k) { g[k] = gt[(long long)i * W + k]; p[k] = pred[(long long)i * W + k]; }
    if (W == 48) p3d::pose_errors_f32<48, 1>(g, p, sd, mc, hipc, use_procrustes, dists + (long long)i * J);
    else {
      float g2[42], p2[42];
      for (int k = 0; k < 42; ++k) { g2[k] = g[k]; p2[k] = p[k]; }
      p3d::pose_errors_f32<42, 0>(g2, p2, sd, mc, hipc, use_procrustes, dists + (long long)i * J);
    }
  }
}

// realtime front-end: enc[n][32] = (h36m(kp[n][36])[use2[i]] - mu2[use2[i]]) / sd2[use2[i]] (fp64)
void hc_openpose_frontend(const double* kp, const int* use2, const double* mu2, const double* sd2, double* enc, int n) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < 32; ++i)
      enc[b * 32 + i] = (p3d::openpose_h36m_coord(kp + 36 * b, use2[i]) - mu2[use2[i]]) / sd2[use2[i]];
}
void hc_openpose_h36m64(const double* kp, double* enc64) {
  for (int d = 0; d < 64; ++d) enc64[d] = p3d::openpose_h36m_coord(kp, d);
}

}
