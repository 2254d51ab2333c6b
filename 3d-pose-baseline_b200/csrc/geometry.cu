// Camera-frame preprocessing kernels (K4): HBM-bound, coalesced, staged through shared memory.
//   cameras.project_point_radial / world_to_camera_frame / camera_to_world_frame  (src/cameras.py:13-90)
//   data_utils.project_to_cameras + transform_world_to_camera + postprocess_3d + normalize_data
//                                                   (src/data_utils.py:233-280,339-364,474-494)
//   data_utils.normalize_data / unNormalizeData     (src/data_utils.py:260-311)
#include <cstdlib>

#include "common.cuh"
#include "math_hd.h"

namespace p3d {
namespace geo {

// H36M joints with a name (data_utils.py:19-39); 2D drops Neck/Nose(14), 3D drops Hip(0)
// (and Spine(12), Neck/Nose(14) when predict_14) - data_utils.py:214-228.
__constant__ int kJoints2D[16] = {0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27};
__constant__ int kJoints3D16[16] = {1, 2, 3, 6, 7, 8, 12, 13, 14, 15, 17, 18, 19, 25, 26, 27};
__constant__ int kJoints3D14[14] = {1, 2, 3, 6, 7, 8, 13, 15, 17, 18, 19, 25, 26, 27};
static const int hJoints2D[16] = {0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27};
static const int hJoints3D16[16] = {1, 2, 3, 6, 7, 8, 12, 13, 14, 15, 17, 18, 19, 25, 26, 27};
static const int hJoints3D14[14] = {1, 2, 3, 6, 7, 8, 13, 15, 17, 18, 19, 25, 26, 27};

template <typename T>
static CamT<T> to_cam(const p3d_camera& c) {
  CamT<T> o;
  for (int i = 0; i < 9; ++i) o.R[i] = static_cast<T>(c.R[i]);
  for (int i = 0; i < 3; ++i) { o.Tr[i] = static_cast<T>(c.T[i]); o.k[i] = static_cast<T>(c.k[i]); }
  for (int i = 0; i < 2; ++i) { o.f[i] = static_cast<T>(c.f[i]); o.c[i] = static_cast<T>(c.c[i]); o.p[i] = static_cast<T>(c.p[i]); }
  return o;
}

// ------------------------------------------------------------------ faithful elementwise forms
template <typename T>
__global__ void project_point_radial_kernel(const T* __restrict__ P, const CamT<T> cam, T* __restrict__ proj,
                                            T* __restrict__ D, T* __restrict__ radial, T* __restrict__ tang,
                                            T* __restrict__ r2o, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    T u, v, d, ra, ta, r2;
    project_point(cam, P[3 * i], P[3 * i + 1], P[3 * i + 2], u, v, d, ra, ta, r2);
    proj[2 * i] = u; proj[2 * i + 1] = v;
    if (D) D[i] = d;
    if (radial) radial[i] = ra;
    if (tang) tang[i] = ta;
    if (r2o) r2o[i] = r2;
  }
}

template <bool INVERSE>
__global__ void rigid_kernel(const double* __restrict__ P, const CamT<double> cam, double* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double a, b, c;
    if (INVERSE) cam_to_world(cam, P[3 * i], P[3 * i + 1], P[3 * i + 2], a, b, c);
    else world_to_cam(cam, P[3 * i], P[3 * i + 1], P[3 * i + 2], a, b, c);
    out[3 * i] = a; out[3 * i + 1] = b; out[3 * i + 2] = c;
  }
}

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------ fused project + normalise
constexpr int MAXCAMS = 8;
// Two cameras side by side (.x = camera 2i, .y = camera 2i + 1): every parameter is one 64-bit constant-bank operand of a
// packed fp32x2 instruction (FFMA2 / FMUL2 / FADD2 on sm_100a).  nT = -T.
struct CamPair { float2 R[9], nT[3], k[3], p[2], f[2], c[2]; };
struct FusedArgs {
  CamPair pair[MAXCAMS / 2];
  CamT<float> cam[MAXCAMS];
  float mean2[32], istd2[32];   // gathered to the used dims
  float mean3[48], istd3[48];
  int ncams;
  int out3;        // 48 or 42
  int nj3;         // 16 or 14
  int predict_14;
};

// One WARP owns a pair of poses (lane = local pose * 16 + joint slot) and never meets a block barrier:
//   world rows (2 x 384 B) : two coalesced float4 loads per lane, prefetched one pair ahead in registers,
//                            staged in a 768 B per-warp smem slab for the 12-byte joint gathers
//   2D output              : one float2 per lane per camera = 256 contiguous bytes per warp
//   3D output              : 12-byte joint rows repacked through a double-buffered 384 B smem slab into
//                            float4 stores (384 contiguous bytes per warp per camera)
// The camera-frame root-centred joint is R(q - hip): T cancels, so the hip is subtracted once in world space.
constexpr int PN_WARPS = 8;
#ifndef PN_BLOCKS
#define PN_BLOCKS 4
#endif
#define PIN(v) asm volatile("" : "+f"(v))
// NC = number of cameras, compile time: the camera loop is fully unrolled so that every camera parameter is an
// immediate constant-bank operand of the FFMA that uses it (with a runtime camera index ptxas emits ~30 LDC per
// camera and the kernel was issue bound: 706 warp instructions per pose pair, SM 85 % busy at 51 % of the HBM
// roofline).  H2/H3: which outputs are requested (compile time: no per-camera uniform branches).
constexpr int PN_POSE_PITCH = 100;   // floats between the two poses of a slab: shifts pose 1 by 4 banks (no 2-way conflicts)
template <int NC, bool H2, bool H3>
__global__ void __launch_bounds__(PN_WARPS * 32, PN_BLOCKS) project_normalize_kernel(const float* __restrict__ world, const __grid_constant__ FusedArgs a,
                                                                            float* __restrict__ x2d, float* __restrict__ y3d, long long N) {
  __shared__ __align__(16) float s_in[PN_WARPS][2 * PN_POSE_PITCH];
  __shared__ __align__(16) float s_out[PN_WARPS][2][2 * 48 + 4];      // +4: dump slot for lanes without a 3D joint
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int lp = lane >> 4, j = lane & 15;
  float* sw = s_in[wib];
  const int j2 = kJoints2D[j];
  const bool has3 = j < a.nj3;
  const int j3 = has3 ? (a.predict_14 ? kJoints3D14[j] : kJoints3D16[j]) : 0;
  // out = v * istd + (-mean * istd): one FFMA per output
  float i2x = a.istd2[2 * j], i2y = a.istd2[2 * j + 1];
  float n2x = -a.mean2[2 * j] * i2x, n2y = -a.mean2[2 * j + 1] * i2y;
  float n3[3], i3[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { i3[d] = has3 ? a.istd3[3 * j + d] : 0.f; n3[d] = has3 ? -a.mean3[3 * j + d] * i3[d] : 0.f; }
  // Pin the per-lane constants in registers: left alone, ptxas rematerialises them inside the loop as lane-indexed
  // constant loads (LDC c[0][R+..]), which serialise 16 ways - measured 2.5x slower.
  PIN(n2x); PIN(n2y); PIN(i2x); PIN(i2y);
#pragma unroll
  for (int d = 0; d < 3; ++d) { PIN(n3[d]); PIN(i3[d]); }
  const int out3 = a.out3;
  const int o_off = has3 ? lp * out3 + 3 * j : 2 * 48;   // smem slot of this lane's 3D joint (or the dump slot)
  const int pair_f4 = 2 * out3 / 4;                       // float4s per camera per pose pair (24 or 21)
  const bool y_vec = H3 && ((reinterpret_cast<uintptr_t>(y3d) & 15) == 0) && (((N * out3) & 3) == 0) && ((2 * out3) % 4 == 0);
  const long long npairs = (N + 1) / 2;
  const long long nwarps = static_cast<long long>(gridDim.x) * PN_WARPS;
  const long long tot_f4 = N * 24;
  const long long x_cam_stride = N * 32, y_cam_stride = N * out3;      // floats between camera planes
  const float4* src = reinterpret_cast<const float4*>(world);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // staging slots: float4 i of the pair (i < 48) -> pose i / 24, padded pose pitch
  const int st0 = (lane < 24 ? lane : lane + 1);           // float4 index of chunk `lane`
  const int st1 = 32 + lane + 1;                           // chunk 32 + lane (lane < 16) is always in pose 1
  long long pair = static_cast<long long>(blockIdx.x) * PN_WARPS + wib;
  float4 c0 = zero4, c1 = zero4;
  if (pair < npairs) {
    const long long b = pair * 48;
    if (b + lane < tot_f4) c0 = __ldcs(src + b + lane);
    if (lane < 16 && b + 32 + lane < tot_f4) c1 = __ldcs(src + b + 32 + lane);
  }
  int buf = 0;
  for (; pair < npairs; pair += nwarps) {
    float4 n0 = zero4, n1 = zero4;                          // prefetch the next pair of this warp
    {
      const long long b = (pair + nwarps) * 48;
      if (b + lane < tot_f4) n0 = __ldcs(src + b + lane);
      if (lane < 16 && b + 32 + lane < tot_f4) n1 = __ldcs(src + b + 32 + lane);
    }
    reinterpret_cast<float4*>(sw)[st0] = c0;
    if (lane < 16) reinterpret_cast<float4*>(sw)[st1] = c1;
    __syncwarp();
    const long long p = pair * 2 + lp;
    const bool live = p < N;
    const bool full = pair * 2 + 1 < N;
    const float* w = sw + lp * PN_POSE_PITCH;
    const float px = w[j2 * 3], py = w[j2 * 3 + 1], pz = w[j2 * 3 + 2];
    const float qx = w[j3 * 3] - w[0], qy = w[j3 * 3 + 1] - w[1], qz = w[j3 * 3 + 2] - w[2];
    __syncwarp();                                    // slab consumed: the next iteration may overwrite it
    float2* xo = reinterpret_cast<float2*>(x2d + p * 32) + j;
    float* yo = y3d + pair * 2 * out3;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (H2) {
        // cameras.project_point_radial (src/cameras.py:39-51) in fp32
        const float dx = px - a.cam[c].Tr[0], dy = py - a.cam[c].Tr[1], dz = pz - a.cam[c].Tr[2];
        const float X0 = a.cam[c].R[0] * dx + a.cam[c].R[1] * dy + a.cam[c].R[2] * dz;
        const float X1 = a.cam[c].R[3] * dx + a.cam[c].R[4] * dy + a.cam[c].R[5] * dz;
        const float X2 = a.cam[c].R[6] * dx + a.cam[c].R[7] * dy + a.cam[c].R[8] * dz;
        const float rz = __fdividef(1.f, X2);          // MUFU.RCP (1 ulp): no slow-path branch
        const float x = X0 * rz, y = X1 * rz;
        const float r2 = x * x + y * y;
        const float radial = 1.f + r2 * (a.cam[c].k[0] + r2 * (a.cam[c].k[1] + r2 * a.cam[c].k[2]));
        const float s = radial + (a.cam[c].p[0] * y + a.cam[c].p[1] * x);
        const float u = a.cam[c].f[0] * (x * s + a.cam[c].p[1] * r2) + a.cam[c].c[0];
        const float v = a.cam[c].f[1] * (y * s + a.cam[c].p[0] * r2) + a.cam[c].c[1];
        if (live) __stcs(xo, make_float2(u * i2x + n2x, v * i2y + n2y));   // 256 contiguous bytes per warp
        xo += x_cam_stride / 2;
      }
      if (H3) {
        float* so = s_out[wib][buf];
        float* o = so + o_off;
        o[0] = (a.cam[c].R[0] * qx + a.cam[c].R[1] * qy + a.cam[c].R[2] * qz) * i3[0] + n3[0];
        o[1] = (a.cam[c].R[3] * qx + a.cam[c].R[4] * qy + a.cam[c].R[5] * qz) * i3[1] + n3[1];
        o[2] = (a.cam[c].R[6] * qx + a.cam[c].R[7] * qy + a.cam[c].R[8] * qz) * i3[2] + n3[2];
        __syncwarp();
        if (y_vec && full) {
          if (lane < pair_f4) __stcs(reinterpret_cast<float4*>(yo) + lane, reinterpret_cast<const float4*>(so)[lane]);
        } else {
          const int tot = (full ? 2 : 1) * out3;
          for (int i = lane; i < tot; i += 32) __stcs(yo + i, so[i]);
        }
        yo += y_cam_stride;
        buf ^= 1;                                    // the other slab is free: its readers passed the __syncwarp above
      }
    }
    c0 = n0; c1 = n1;
  }
}

// ---- packed fp32x2 arithmetic (sm_100a): one instruction works on two cameras
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 pk(const float2& c) { return pk(c.x, c.y); }
__device__ __forceinline__ void unpk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// The same kernel with the cameras taken two at a time through packed fp32x2 instructions.  The scalar kernel above is
// issue bound (~400 warp instructions per pose pair at 4 cameras, 187 of them FFMA/FMUL/FADD); pairing the cameras halves
// the arithmetic instructions and the warp barriers of the 3D repack.  NP = number of camera PAIRS (ncams = 2 NP).
// Rounding: every operation is the same IEEE fp32 operation as in the scalar kernel, on the same operands.
template <int NP, bool H2, bool H3>
__global__ void __launch_bounds__(PN_WARPS * 32, PN_BLOCKS) project_normalize_pair_kernel(const float* __restrict__ world, const __grid_constant__ FusedArgs a,
                                                                                 float* __restrict__ x2d, float* __restrict__ y3d, long long N) {
  __shared__ __align__(16) float s_in[PN_WARPS][2 * PN_POSE_PITCH];
  __shared__ __align__(16) float s_out[PN_WARPS][4][2 * 48 + 4];      // two camera pairs in flight; +4: dump slot
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int lp = lane >> 4, j = lane & 15;
  float* sw = s_in[wib];
  const int j2 = kJoints2D[j];
  const bool has3 = j < a.nj3;
  const int j3 = has3 ? (a.predict_14 ? kJoints3D14[j] : kJoints3D16[j]) : 0;
  float i2x = a.istd2[2 * j], i2y = a.istd2[2 * j + 1];
  float n2x = -a.mean2[2 * j] * i2x, n2y = -a.mean2[2 * j + 1] * i2y;
  float n3[3], i3[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { i3[d] = has3 ? a.istd3[3 * j + d] : 0.f; n3[d] = has3 ? -a.mean3[3 * j + d] * i3[d] : 0.f; }
  PIN(n2x); PIN(n2y); PIN(i2x); PIN(i2y);
#pragma unroll
  for (int d = 0; d < 3; ++d) { PIN(n3[d]); PIN(i3[d]); }
  const f2 I2x = pk(i2x, i2x), I2y = pk(i2y, i2y), N2x = pk(n2x, n2x), N2y = pk(n2y, n2y);
  const f2 one2 = pk(1.f, 1.f);
  const int out3 = a.out3;
  const int o_off = has3 ? lp * out3 + 3 * j : 2 * 48;
  const int pair_f4 = 2 * out3 / 4;
  const bool y_vec = H3 && ((reinterpret_cast<uintptr_t>(y3d) & 15) == 0) && (((N * out3) & 3) == 0) && ((2 * out3) % 4 == 0);
  const long long npairs = (N + 1) / 2;
  const long long nwarps = static_cast<long long>(gridDim.x) * PN_WARPS;
  const long long tot_f4 = N * 24;
  const long long x_cam_stride = N * 32, y_cam_stride = N * out3;
  const float4* src = reinterpret_cast<const float4*>(world);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int st0 = (lane < 24 ? lane : lane + 1);
  const int st1 = 32 + lane + 1;
  long long pair = static_cast<long long>(blockIdx.x) * PN_WARPS + wib;
  float4 c0 = zero4, c1 = zero4;
  if (pair < npairs) {
    const long long b = pair * 48;
    if (b + lane < tot_f4) c0 = __ldcs(src + b + lane);
    if (lane < 16 && b + 32 + lane < tot_f4) c1 = __ldcs(src + b + 32 + lane);
  }
  int buf = 0;                                               // slabs (buf, buf + 1) belong to the current camera pair
  for (; pair < npairs; pair += nwarps) {
    float4 n0 = zero4, n1 = zero4;
    {
      const long long b = (pair + nwarps) * 48;
      if (b + lane < tot_f4) n0 = __ldcs(src + b + lane);
      if (lane < 16 && b + 32 + lane < tot_f4) n1 = __ldcs(src + b + 32 + lane);
    }
    reinterpret_cast<float4*>(sw)[st0] = c0;
    if (lane < 16) reinterpret_cast<float4*>(sw)[st1] = c1;
    __syncwarp();
    const long long p = pair * 2 + lp;
    const bool live = p < N;
    const bool full = pair * 2 + 1 < N;
    const float* w = sw + lp * PN_POSE_PITCH;
    const float px = w[j2 * 3], py = w[j2 * 3 + 1], pz = w[j2 * 3 + 2];
    const float qx = w[j3 * 3] - w[0], qy = w[j3 * 3 + 1] - w[1], qz = w[j3 * 3 + 2] - w[2];
    __syncwarp();
    const f2 PX = pk(px, px), PY = pk(py, py), PZ = pk(pz, pz);
    const f2 QX = pk(qx, qx), QY = pk(qy, qy), QZ = pk(qz, qz);
    float2* xo = reinterpret_cast<float2*>(x2d + p * 32) + j;
    float* yo = y3d + pair * 2 * out3;
#pragma unroll
    for (int cp = 0; cp < NP; ++cp) {
      const CamPair& C = a.pair[cp];
      if (H2) {
        // cameras.project_point_radial (src/cameras.py:39-51) in fp32, cameras 2cp and 2cp+1 in the two halves
        const f2 dx = add2(PX, pk(C.nT[0])), dy = add2(PY, pk(C.nT[1])), dz = add2(PZ, pk(C.nT[2]));
        const f2 X0 = fma2(pk(C.R[2]), dz, fma2(pk(C.R[1]), dy, mul2(pk(C.R[0]), dx)));
        const f2 X1 = fma2(pk(C.R[5]), dz, fma2(pk(C.R[4]), dy, mul2(pk(C.R[3]), dx)));
        const f2 X2 = fma2(pk(C.R[8]), dz, fma2(pk(C.R[7]), dy, mul2(pk(C.R[6]), dx)));
        float za, zb;
        unpk(X2, za, zb);
        const f2 rz = pk(__fdividef(1.f, za), __fdividef(1.f, zb));          // MUFU.RCP per camera
        const f2 x = mul2(X0, rz), y = mul2(X1, rz);
        const f2 r2 = fma2(y, y, mul2(x, x));
        const f2 radial = fma2(r2, fma2(r2, fma2(r2, pk(C.k[2]), pk(C.k[1])), pk(C.k[0])), one2);
        const f2 s = add2(radial, fma2(pk(C.p[1]), x, mul2(pk(C.p[0]), y)));
        const f2 u = fma2(pk(C.f[0]), fma2(pk(C.p[1]), r2, mul2(x, s)), pk(C.c[0]));
        const f2 v = fma2(pk(C.f[1]), fma2(pk(C.p[0]), r2, mul2(y, s)), pk(C.c[1]));
        const f2 un = fma2(u, I2x, N2x), vn = fma2(v, I2y, N2y);
        float ua, ub, va, vb;
        unpk(un, ua, ub); unpk(vn, va, vb);
        if (live) {
          __stcs(xo, make_float2(ua, va));
          __stcs(xo + x_cam_stride / 2, make_float2(ub, vb));
        }
        xo += x_cam_stride;                                  // two camera planes (in float2 units: 2 * stride / 2)
      }
      if (H3) {
        float* soa = s_out[wib][buf];
        float* sob = s_out[wib][buf + 1];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const f2 r = fma2(pk(C.R[3 * d + 2]), QZ, fma2(pk(C.R[3 * d + 1]), QY, mul2(pk(C.R[3 * d]), QX)));
          const f2 o = fma2(r, pk(i3[d], i3[d]), pk(n3[d], n3[d]));
          float oa, ob;
          unpk(o, oa, ob);
          soa[o_off + d] = oa; sob[o_off + d] = ob;
        }
        __syncwarp();
        if (y_vec && full) {
          if (lane < pair_f4) {
            __stcs(reinterpret_cast<float4*>(yo) + lane, reinterpret_cast<const float4*>(soa)[lane]);
            __stcs(reinterpret_cast<float4*>(yo + y_cam_stride) + lane, reinterpret_cast<const float4*>(sob)[lane]);
          }
        } else {
          const int tot = (full ? 2 : 1) * out3;
          for (int i = lane; i < tot; i += 32) { __stcs(yo + i, soa[i]); __stcs(yo + y_cam_stride + i, sob[i]); }
        }
        yo += 2 * y_cam_stride;
        buf ^= 2;              // the other two slabs are free: their readers passed the __syncwarp above
      }
    }
    c0 = n0; c1 = n1;
  }
}

static bool pn_packed_enabled() {   // P3D_PN_PACKED=0: scalar kernel for every camera count (A/B measurements)
  static const bool on = [] { const char* e = getenv("P3D_PN_PACKED"); return !(e && e[0] == '0'); }();
  return on;
}
template <int NC>
static void launch_project_normalize(int grid, cudaStream_t st, const float* world, const FusedArgs& a, float* x2d, float* y3d, long long N) {
  if constexpr ((NC % 2) == 0) {
    if (pn_packed_enabled()) {       // an even number of cameras: two per packed fp32x2 instruction
      if (x2d && y3d) project_normalize_pair_kernel<NC / 2, true, true><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
      else if (x2d) project_normalize_pair_kernel<NC / 2, true, false><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
      else project_normalize_pair_kernel<NC / 2, false, true><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
      return;
    }
  }
  if (x2d && y3d) project_normalize_kernel<NC, true, true><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
  else if (x2d) project_normalize_kernel<NC, true, false><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
  else project_normalize_kernel<NC, false, true><<<grid, PN_WARPS * 32, 0, st>>>(world, a, x2d, y3d, N);
}

// ------------------------------------------------------------------ normalise / un-normalise (f64, one array)
struct NormArgs {
  double mean[96], stdv[96];
  short use[96];     // normalize: use[d] = source column of output d ; unnormalize: use[j] = input column of full dim j or -1
  int din, dout;
};

__global__ void normalize_kernel(const double* __restrict__ in, const __grid_constant__ NormArgs a, double* __restrict__ out, long long N) {
  const long long total = N * a.dout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / a.dout;
    const int d = static_cast<int>(i - r * a.dout);
    const int s = a.use[d];
    out[i] = (in[r * a.din + s] - a.mean[s]) / a.stdv[s];
  }
}

__global__ void unnormalize_kernel(const double* __restrict__ in, const __grid_constant__ NormArgs a, double* __restrict__ out, long long N) {
  const long long total = N * a.dout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / a.dout;
    const int j = static_cast<int>(i - r * a.dout);
    const int s = a.use[j];
    // the reference scatters into a float32 matrix first (data_utils.py:299-303)
    const double v = (s >= 0) ? static_cast<double>(static_cast<float>(in[r * a.din + s])) : 0.0;
    out[i] = v * a.stdv[j] + a.mean[j];
  }
}


// ------------------------------------------------------------------ column statistics / root centring (f64)
// data_utils.normalization_stats (src/data_utils.py:211-212): column mean and POPULATION std.
// pass 0 accumulates sum(x), pass 1 accumulates sum((x-mean)^2) - two passes keep fp64 accuracy.
__global__ void colstat_pass_kernel(const double* __restrict__ data, long long N, int D, const double* __restrict__ mean_or_null,
                                    double invN, double* __restrict__ acc) {
  __shared__ double sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * 1024;
  const long long r1 = (r0 + 1024 < N) ? r0 + 1024 : N;
  double s = 0;
  if (c < D) {
    const double mu = mean_or_null ? mean_or_null[c] * invN : 0.0;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) { const double v = data[r * D + c] - mu; s += mean_or_null ? v * v : v; }
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x];
    atomicAdd(acc + c, s);
  }
}
__global__ void colstat_finish_kernel(const double* sum, const double* ssq, double invN, int D, double* mean, double* stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) { mean[c] = sum[c] * invN; stdv[c] = sqrt(ssq[c] * invN); }
}
// data_utils.postprocess_3d (src/data_utils.py:474-494): subtract the root joint from every joint
__global__ void root_center_kernel(const double* __restrict__ in, double* __restrict__ out, double* __restrict__ roots, long long N, int D) {
  const long long total = N * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / D;
    const int c = static_cast<int>(i - r * D);
    const double root = in[r * D + (c % 3)];
    if (roots && c < 3) roots[r * 3 + c] = root;
    out[i] = in[i] - root;
  }
}

static int use_table(int dim, int predict_14, int* use, int* nuse) {
  int n = 0;
  if (dim == 2) {
    for (int j = 0; j < 16; ++j) { use[n++] = hJoints2D[j] * 2; use[n++] = hJoints2D[j] * 2 + 1; }
  } else if (dim == 3) {
    const int nj = predict_14 ? 14 : 16;
    const int* jt = predict_14 ? hJoints3D14 : hJoints3D16;
    for (int j = 0; j < nj; ++j) { use[n++] = jt[j] * 3; use[n++] = jt[j] * 3 + 1; use[n++] = jt[j] * 3 + 2; }
  } else {
    set_error("dim must be 2 or 3");
    return P3D_ERR_ARG;
  }
  *nuse = n;
  return P3D_OK;
}

}  // namespace geo
}  // namespace p3d

using namespace p3d;
using namespace p3d::geo;

extern "C" {

int p3d_project_point_radial_f64(const double* P, const p3d_camera* cam, double* proj, double* D, double* radial,
                                 double* tan_, double* r2, int64_t npts, void* stream) {
  P3D_REQUIRE(P && cam && proj && npts >= 0, "project_point_radial: null argument");
  if (npts == 0) return P3D_OK;
  project_point_radial_kernel<double><<<grid_for(npts), 256, 0, (cudaStream_t)stream>>>(P, to_cam<double>(*cam), proj, D, radial, tan_, r2, npts);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_project_point_radial_f32(const float* P, const p3d_camera* cam, float* proj, float* D, float* radial,
                                 float* tan_, float* r2, int64_t npts, void* stream) {
  P3D_REQUIRE(P && cam && proj && npts >= 0, "project_point_radial: null argument");
  if (npts == 0) return P3D_OK;
  project_point_radial_kernel<float><<<grid_for(npts), 256, 0, (cudaStream_t)stream>>>(P, to_cam<float>(*cam), proj, D, radial, tan_, r2, npts);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_world_to_camera_f64(const double* P, const p3d_camera* cam, double* out, int64_t npts, void* stream) {
  P3D_REQUIRE(P && cam && out && npts >= 0, "world_to_camera: null argument");
  if (npts == 0) return P3D_OK;
  rigid_kernel<false><<<grid_for(npts), 256, 0, (cudaStream_t)stream>>>(P, to_cam<double>(*cam), out, npts);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_camera_to_world_f64(const double* P, const p3d_camera* cam, double* out, int64_t npts, void* stream) {
  P3D_REQUIRE(P && cam && out && npts >= 0, "camera_to_world: null argument");
  if (npts == 0) return P3D_OK;
  rigid_kernel<true><<<grid_for(npts), 256, 0, (cudaStream_t)stream>>>(P, to_cam<double>(*cam), out, npts);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_project_normalize(const float* world, const p3d_camera* cams, int ncams, const double* mean2d, const double* std2d,
                          const double* mean3d, const double* std3d, int predict_14, float* x2d, float* y3d, int64_t N,
                          void* stream) {
  P3D_REQUIRE(world && cams && ncams >= 1 && ncams <= MAXCAMS, "project_normalize: need 1..%d cameras", MAXCAMS);
  P3D_REQUIRE(x2d || y3d, "project_normalize: no output requested");
  P3D_REQUIRE(!x2d || (mean2d && std2d), "project_normalize: 2D statistics missing");
  P3D_REQUIRE(!y3d || (mean3d && std3d), "project_normalize: 3D statistics missing");
  if (N == 0) return P3D_OK;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.ncams = ncams;
  for (int c = 0; c < ncams; ++c) a.cam[c] = to_cam<float>(cams[c]);
  for (int c = 0; c + 1 < ncams; c += 2) {
    const CamT<float>& A = a.cam[c]; const CamT<float>& B = a.cam[c + 1];
    CamPair& P = a.pair[c / 2];
    for (int i = 0; i < 9; ++i) P.R[i] = make_float2(A.R[i], B.R[i]);
    for (int i = 0; i < 3; ++i) { P.nT[i] = make_float2(-A.Tr[i], -B.Tr[i]); P.k[i] = make_float2(A.k[i], B.k[i]); }
    for (int i = 0; i < 2; ++i) { P.p[i] = make_float2(A.p[i], B.p[i]); P.f[i] = make_float2(A.f[i], B.f[i]); P.c[i] = make_float2(A.c[i], B.c[i]); }
  }
  int use[96], n;
  if (x2d) {
    P3D_TRY(use_table(2, 0, use, &n));
    for (int d = 0; d < n; ++d) { a.mean2[d] = (float)mean2d[use[d]]; a.istd2[d] = (float)(1.0 / std2d[use[d]]); }
  }
  a.predict_14 = predict_14 ? 1 : 0;
  a.nj3 = predict_14 ? 14 : 16;
  a.out3 = a.nj3 * 3;
  if (y3d) {
    P3D_TRY(use_table(3, predict_14, use, &n));
    for (int d = 0; d < n; ++d) { a.mean3[d] = (float)mean3d[use[d]]; a.istd3[d] = (float)(1.0 / std3d[use[d]]); }
  }
  const long long nblocks = ((N + 1) / 2 + PN_WARPS - 1) / PN_WARPS;
  const int grid = nblocks < 148 * PN_BLOCKS ? (int)nblocks : 148 * PN_BLOCKS;
  cudaStream_t st = (cudaStream_t)stream;
  switch (ncams) {
#define P3D_PN_CASE(NC) case NC: launch_project_normalize<NC>(grid, st, world, a, x2d, y3d, N); break;
    P3D_PN_CASE(1) P3D_PN_CASE(2) P3D_PN_CASE(3) P3D_PN_CASE(4) P3D_PN_CASE(5) P3D_PN_CASE(6) P3D_PN_CASE(7) P3D_PN_CASE(8)
#undef P3D_PN_CASE
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_normalize_f64(const double* in, const double* mean, const double* stdv, int dim, int predict_14, double* out,
                      int64_t N, void* stream) {
  P3D_REQUIRE(in && mean && stdv && out, "normalize: null argument");
  NormArgs a;
  int use[96], n;
  P3D_TRY(use_table(dim, predict_14, use, &n));
  a.din = dim == 2 ? 64 : 96; a.dout = n;
  for (int j = 0; j < a.din; ++j) { a.mean[j] = mean[j]; a.stdv[j] = stdv[j]; }
  for (int d = 0; d < n; ++d) a.use[d] = (short)use[d];
  if (N == 0) return P3D_OK;
  normalize_kernel<<<grid_for(N * a.dout), 256, 0, (cudaStream_t)stream>>>(in, a, out, N);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_unnormalize_f64(const double* in, const double* mean, const double* stdv, int dim, int predict_14, double* out,
                        int64_t N, void* stream) {
  P3D_REQUIRE(in && mean && stdv && out, "unnormalize: null argument");
  NormArgs a;
  int use[96], n;
  P3D_TRY(use_table(dim, predict_14, use, &n));
  a.dout = dim == 2 ? 64 : 96; a.din = n;
  for (int j = 0; j < a.dout; ++j) { a.mean[j] = mean[j]; a.stdv[j] = stdv[j]; a.use[j] = -1; }
  for (int d = 0; d < n; ++d) a.use[use[d]] = (short)d;
  if (N == 0) return P3D_OK;
  unnormalize_kernel<<<grid_for(N * a.dout), 256, 0, (cudaStream_t)stream>>>(in, a, out, N);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_column_stats_f64(const double* data, int64_t N, int D, double* mean, double* stdv, double* work /*[2*D]*/, void* stream) {
  P3D_REQUIRE(data && mean && stdv && work && N >= 1 && D >= 1, "column_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  P3D_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 2 * D, st));
  dim3 grid((D + 31) / 32, (unsigned)((N + 1023) / 1024)), block(32, 8);
  colstat_pass_kernel<<<grid, block, 0, st>>>(data, N, D, nullptr, 1.0 / N, work);
  P3D_LAUNCH_CHECK();
  colstat_pass_kernel<<<grid, block, 0, st>>>(data, N, D, work, 1.0 / N, work + D);
  P3D_LAUNCH_CHECK();
  colstat_finish_kernel<<<(D + 127) / 128, 128, 0, st>>>(work, work + D, 1.0 / N, D, mean, stdv);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int p3d_root_center_f64(const double* in, double* out, double* roots_or_null, int64_t N, int D, void* stream) {
  P3D_REQUIRE(in && out && N >= 0 && D >= 3 && D % 3 == 0, "root_center: bad argument");
  if (N == 0) return P3D_OK;
  P3D_REQUIRE(in != out, "root_center: in-place is not supported");
  root_center_kernel<<<grid_for(N * D), 256, 0, (cudaStream_t)stream>>>(in, out, roots_or_null, N, D);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

}  // extern "C"
