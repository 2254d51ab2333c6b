// Generic tcgen05 GEMM for the training step (src/linear_model.py:129-145: the MatMuls of the forward
// graph, of tf.gradients and of the weight gradients), bf16 operands, fp32 accumulation in TMEM.
//
//   C[M,N] (+)= alpha * sum_k A(m,k) B(n,k)  (+ bias[n]) (+ res[m,n])
//
// Each operand is described by how it lies in memory:
//   K-major  : [rows = M or N][K contiguous]   (activations for Z = H W;  dZ and W for dH = dZ W^T)
//   MN-major : [K rows][M or N contiguous]     (W for Z = H W;  H and dZ for dW = H^T dZ)
// so the three GEMMs of a layer read H, dZ and W in their NATURAL row-major layouts - no transposed
// copies exist anywhere.  MN-major tiles are fetched as 64-column TMA boxes (128B swizzle) and consumed
// through the MN-major canonical UMMA layout ((8,n),(8,k)):((1,LBO),(8,SBO)) [16-byte units].
//
// One CTA per 128 x BN output tile (BN = 64/128/256 chosen per problem so the grid covers the SMs),
// optional split-K over gridDim.z (fp32 atomics into a zeroed C - the weight-gradient GEMMs have only
// (L/128)*(L/BN) tiles but K = batch).  Warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> alpha/bias/residual -> global).  Out-of-range rows/cols
// and the K tail are zero-filled by TMA, so M, N, K need no padding (K=32 input layer, N=48 output).
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"
#include "train_common.cuh"

namespace p3d {
namespace tcg {

using namespace ptx;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 16;
constexpr int A_BYTES = BM * BK * 2;        // 16 KB
constexpr int BOX_BYTES = 64 * BK * 2;      // one [64 rows x 64 cols] bf16 box = 8 KB
constexpr int NTHREADS = 320;             // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int EPI_THREADS = 256;
constexpr int RING_BYTES = 4 * (A_BYTES + 256 * BK * 2);   // 192 KB: 4 stages at BN=256, 6 at 128, 8 at 64
constexpr int BAR_BYTES = 512;               // full[16] | empty[16] | accf | tmem slot
constexpr int SMEM_BYTES = 1024 + RING_BYTES + BAR_BYTES + 9 * 1024;   // align slack | operand ring | barriers | bias[256] | column sums [4 quadrants][2][256]
// OCC = 2 instantiations: two CTAs share an SM (96 registers per thread, 106.5 KB of shared memory each, BN <= 128 so
// that two accumulators fit the 512 TMEM columns).  One CTA's prologue / epilogue then runs under the other one's
// mainloop - the overlap a one-tile-per-CTA kernel cannot give itself.
constexpr int RING2_BYTES = 3 * (A_BYTES + 128 * BK * 2);  // 96 KB: 3 stages at BN=128, 4 at 64
constexpr int SMEM2_BYTES = 1024 + RING2_BYTES + BAR_BYTES + 9 * 1024;
// the narrower the tile, the deeper the ring: small problems are latency bound on the L2 -> SM round trip


struct Params {
  int M, N, K;
  int bn;            // 64 / 128 / 256
  int k_per_split;   // multiple of BK
  int a_mn, b_mn;
  float* C; int ldc;
  const float* bias;        // [N] or null (added by split 0 only)
  const float* res; int ldres;   // [M,N] or null (added by split 0 only)
  const float* alpha_dev;   // optional device scalar
  float alpha;
  int atomic;               // accumulate into C with fp32 atomics (split-K)
  double* colsum;           // optional [2][N]: column sums of the stored values and of their squares
  // bf16 output (inference layers): out = bf16(relu?(alpha*acc + bias)); with res_b: out = bf16(out + res_b)
  __nv_bfloat16* out_b; int ldob;
  const __nv_bfloat16* res_b; int ldrb;
  int relu;
  int cn;                   // cluster size along N (1, 2, 4): the CTAs of one M tile share A - each fetches 1/cn of its rows and multicasts
  int occ;                  // 2 = the OCC = 2 instantiation (two CTAs per SM, small ring); 1 otherwise
  int cg;                   // 2 = CTA pair along M (cta_group::2): one 256 x BN tile per pair, each CTA stages its own 128 rows of A
                            // and HALF of the B tile - a third less L2 -> SM operand traffic per FLOP (1 = single-CTA tiles)
  int stages;               // ring depth: what fits into 192 KB at this tile size, 16 at most
  int a_bytes;              // bytes of A actually fetched per stage (K-major A of a short problem: only round_up(M, 8) rows)
  int b_independent;        // B does not depend on preceding kernels of the stream (weights): prefetch it before the PDL wait
  unsigned long long* dbg;  // optional [ctas][16] globaltimer stamps (diagnostics)
  FusedTrain ft;            // OUT = 3 / 4 only
  int ts;                   // 1 = the TS instantiation (the tile leaves through TMA stores, tm_c describes C)
  int fi;                   // 1 = the FI instantiation (lean MMA issue path)
};
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define P3D_STAMP(i) do { if (p.dbg && lane == 0) p.dbg[(static_cast<size_t>(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (i)] = gtime(); } while (0)

__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// K-major operand, 128B swizzle: rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_k(uint32_t smem_addr) {
  constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// MN-major operand, 128B swizzle: 64 MN-elements (128 B) contiguous, k rows 128 B apart, 8-k groups SBO = 1024 B
// apart, the next 64 MN-elements LBO = one box (8 KB) further
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr) {
  constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint32_t>(BOX_BYTES >> 4) << 16);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// shared-space accessors: the tile pointer is derived from the aligned dynamic-smem base by integer arithmetic, which
// makes the compiler fall back to GENERIC ld/st (long-scoreboard latency) - say the address space explicitly
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// OUT: 0 = fp32 store, 1 = fp32 reduction (split-K / accumulate), 2 = bf16 (relu, residual in bf16)
// RES: a residual operand is added;  CS: 1 = column sums of the result and its square are accumulated (BatchNorm
// statistics of a forward layer); 2 = the result is dh of a hidden layer: column sums of da = dh * dropout * relu'
// and of da * xhat (the BatchNorm backward sums, what train.cu's bwd_act_kernel computes in a pass of its own)
// CG: 1 = single-CTA tiles, 2 = CTA pair along M (cta_group::2).  A template parameter, not a runtime flag: a kernel
// that CONTAINS cta_group::2 instructions can only be launched with an even cluster size ("cluster misconfiguration"
// otherwise, even if the instructions are never reached - measured), so the single-CTA kernels must not contain them.
// TS: the fp32 tile leaves through TMA (cp.async.bulk.tensor / cp.reduce.async.bulk.tensor.add for split-K) instead of
// per-lane st.global / red.global: 32 x 32 chunks staged row-per-lane (as tcgen05.ld delivers them) in 128B-swizzled
// smem, edges clipped by the tensor map of C (tm_c; unused by the other instantiations).
// FI: lean MMA issue path (descriptors advanced by addition instead of rebuilt per k-block), see the MMA warp.
template <int OUT, bool RES, int CS, int CG = 1, int OCC = 1, bool TS = false, bool FI = false>
__global__ void __launch_bounds__(NTHREADS, OCC)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // the narrower the tile, the deeper the ring: small problems are latency bound on the L2 -> SM round trip
  const int STAGES = p.stages;
  constexpr int cg = CG;                         // 1, or 2 = cta_group::2 pair along M (cluster 2 x 1 x 1, M tiles along x)
  const int B_BYTES = (p.bn / cg) * BK * 2;      // this CTA's part of the B tile
  const int A_SLOT = p.a_bytes;                  // slots are as large as what is fetched (multiple of 1024 B)
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_SLOT;
  constexpr int RING = (OCC == 2) ? RING2_BYTES : RING_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accf = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);
  float* sbias = reinterpret_cast<float*>(smem + RING + BAR_BYTES);
  float* scol = sbias + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = (p.cn > 1 || cg == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cg == 1) || (crank == 0);   // of a pair: issues the MMAs, owns the "full" barriers
  if (warp == 0) P3D_STAMP(0);
  // pairs: the two CTAs of a cta_group::2 pair must be neighbours in x (a 1 x 2 x 1 cluster is refused at launch as
  // "cluster misconfiguration" - measured), so the pair kernels take the M tile from blockIdx.x
  const int n0 = (CG == 2 ? blockIdx.y : blockIdx.x) * p.bn, m0 = (CG == 2 ? blockIdx.x : blockIdx.y) * BM;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = (kbeg + p.k_per_split < p.K) ? kbeg + p.k_per_split : p.K;
  const int nk = (kend - kbeg + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
    if constexpr (TS) tma_prefetch_desc(&tm_c);
    // a ring slot is free again when the MMAs of EVERY CTA of the cluster have read it (peers multicast into it)
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], p.cn); }
    mbar_init(accf, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cg == 2) { tmem_alloc_2sm(tmem_slot, static_cast<uint32_t>(p.bn)); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, static_cast<uint32_t>(p.bn)); tmem_relinquish(); }
  }
  tc_fence_before();
  if (p.cn > 1 || cg == 2) cluster_sync(); else __syncthreads();     // peers must see initialised barriers before the first multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_launch_dependents();     // the next kernel of a programmatic-dependent chain may start its prologue now
  if (warp == 0) P3D_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Programmatic dependent launch: when B is marked independent of the preceding kernels (weights), its tiles for
    // the first ring pass are requested BEFORE griddepcontrol.wait, i.e. while the previous layer is still running.
    const uint32_t bytes = p.a_bytes + B_BYTES;
    // Every TMA operation of a stage belongs to its own lane and all of them leave in ONE instruction: issuing them
    // one after the other from a single thread costs ~0.12 us each, which is what bounded short-M GEMMs (16 serial
    // k-blocks x 2 operations = 4.3 us of a 7.6 us kernel).  Lanes [0, na) fetch A boxes, lanes [na, na + nb) B boxes.
    const int na = p.a_mn ? 2 : 1, nb = p.b_mn ? (p.bn / cg) / 64 : 1;
    const bool is_a = lane < na, is_b = lane >= na && lane < na + nb;
    const int bi = lane - na;                                         // B box index of this lane
    const CUtensorMap* my_map = is_a ? &tm_a : &tm_b;
    const bool my_mn = is_a ? (p.a_mn != 0) : (p.b_mn != 0);
    const int my_box = is_a ? lane : bi;
    const int my_row0 = is_a ? m0 : n0 + (cg == 2 ? static_cast<int>(crank) * (p.bn / 2) : 0);   // pair: this CTA's half of the N tile
    const uint32_t a_mask = static_cast<uint16_t>((1u << p.cn) - 1u);
    auto issue = [&](int stage, int k0, bool with_a, bool with_b) {
      if ((is_a && with_a) || (is_b && with_b)) {
        uint8_t* dst = (is_a ? sA + stage * A_SLOT : sB + stage * B_BYTES) + my_box * BOX_BYTES;
        const int c0 = my_mn ? my_row0 + 64 * my_box : k0, c1 = my_mn ? k0 : my_row0;
        if (is_a && p.cn > 1) {     // this CTA's share of the rows, delivered to every CTA of the cluster (same smem offset, same barrier)
          const int part = p.a_bytes / p.cn, rows = part / (BK * 2);
          tma_load_2d_mcast(dst + crank * part, my_map, &full[stage], k0, m0 + static_cast<int>(crank) * rows, static_cast<uint16_t>(a_mask));
        } else if constexpr (cg == 2) {       // into this CTA's smem, bytes signalled on the LEADER's barrier
          tma_load_2d_2sm(dst, my_map, &full[stage], c0, c1);
        } else {
          tma_load_2d(dst, my_map, &full[stage], c0, c1);
        }
      }
    };
    // pair: both CTAs' bytes are accounted on the leader's barrier (the peer's complete_tx may precede this arrive -
    // the phase cannot complete before the one pending arrival has happened)
    auto expect = [&](int stage) { if (lane == 0 && leader) mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(cg) * bytes); };
    // Programmatic dependent launch: when B is marked independent of the preceding kernels (weights), its tiles for
    // the first ring pass are requested BEFORE griddepcontrol.wait, i.e. while the previous layer is still running.
    const int npre = p.b_independent ? (nk < STAGES ? nk : STAGES) : 0;
    bool lean = false;
    if constexpr (FI) lean = (p.cn == 1);
    if (lean) {
      // FI: everything that does not change from k-block to k-block - which box this lane fetches, through which map,
      // into which offset of a slot, its fixed coordinate - is settled before the loop; a k-block is barrier wait,
      // expect_tx, one TMA instruction per box-owning lane, and ONE test of p.dbg in front of the diagnostics stamps
      // (the general loop below re-derives them through ~110 SASS instructions and a jump table per k-block).
      const uint32_t dst0 = smem_u32(is_a ? sA : sB) + static_cast<uint32_t>(my_box) * BOX_BYTES;
      const uint32_t slot = is_a ? static_cast<uint32_t>(A_SLOT) : static_cast<uint32_t>(B_BYTES);
      const int fix = my_mn ? my_row0 + 64 * my_box : my_row0;             // MN-major: the row coordinate of this lane's box
      const uint32_t full_s = smem_u32(full);
      auto load = [&](int stg, int k0) {
        tma_load_2d_s(dst0 + static_cast<uint32_t>(stg) * slot, my_map, full_s + static_cast<uint32_t>(stg) * 8u, my_mn ? fix : k0, my_mn ? k0 : fix);
      };
      for (int kb = 0; kb < npre; ++kb) {            // weights of the first ring pass, ahead of the dependency wait
        if (lane == 0) mbar_arrive_expect_tx(&full[kb], bytes);
        __syncwarp();
        if (is_b) load(kb, kbeg + kb * BK);
      }
      __syncwarp();
      grid_dependency_wait();
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1, 1);
        if (kb >= npre && lane == 0) mbar_arrive_expect_tx(&full[stage], bytes);
        __syncwarp();
        if (is_a || (is_b && kb >= npre)) load(stage, kbeg + kb * BK);
        if (p.dbg) { if (kb == 3) P3D_STAMP(8); if (kb == 7) P3D_STAMP(9); if (kb == 11) P3D_STAMP(10); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else {
    for (int kb = 0; kb < npre; ++kb) {              // ring slots are free on the first pass
      expect(kb);
      __syncwarp();
      issue(kb, kbeg + kb * BK, false, true);
    }
    __syncwarp();
    grid_dependency_wait();
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < nk; ++kb) {
      const int k0 = kbeg + kb * BK;
      mbar_wait(&empty[stage], phase ^ 1, 1);
      if (kb >= npre) expect(stage);
      __syncwarp();
      issue(stage, k0, true, kb >= npre);
      if (kb == 3) P3D_STAMP(8);
      if (kb == 7) P3D_STAMP(9);
      if (kb == 11) P3D_STAMP(10);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    }
    P3D_STAMP(2);
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (of a pair: the leader CTA only)
    if (leader) {
    const uint32_t idesc = umma_idesc_bf16_f32(BM * cg, p.bn) | (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;   // bytes per K=16 slice
    int stage = 0; uint32_t phase = 0;
    if constexpr (FI) {
      // Lean issue path.  The general loop below rebuilds eight shared-memory descriptors per k-block from byte addresses
      // (shift, mask, layout bits chosen by runtime flags: ~50 uniform-datapath instructions and 16 R2UR in the SASS) and
      // walks two jump tables for the diagnostics stamps - ~110 instructions between the barrier and the first
      // tcgen05.mma, against 4 x 56 cycles of tensor work at N <= 64 (0.14 us of fixed cost per k-block).
      // A descriptor's start-address field is (address >> 4) in its low 14 bits and every operand address is a multiple
      // of 16 below 256 KB, so the field advances LINEARLY: one descriptor per operand is built before the loop, a
      // k-block adds stage * (slot >> 4), a K = 16 slice adds (step >> 4) - no carry can leave the field.
      static_assert(CG == 1, "the lean issue path is single-CTA");
      const uint64_t a0 = p.a_mn ? desc_mn(a_base) : desc_k(a_base), b0 = p.b_mn ? desc_mn(b_base) : desc_k(b_base);
      const uint32_t a_slot16 = static_cast<uint32_t>(A_SLOT) >> 4, b_slot16 = static_cast<uint32_t>(B_BYTES) >> 4;
      const uint32_t a_k16 = a_step >> 4, b_k16 = b_step >> 4;
      const uint16_t cmask = static_cast<uint16_t>((1u << p.cn) - 1u);
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&full[stage], phase, 2);
        if (p.dbg) { if (kb == 0) P3D_STAMP(3); if (kb == 4) P3D_STAMP(11); if (kb == 8) P3D_STAMP(12); if (kb == 12) P3D_STAMP(13); }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = a0 + static_cast<uint32_t>(stage) * a_slot16, bd = b0 + static_cast<uint32_t>(stage) * b_slot16;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(tmem_base, ad + static_cast<uint32_t>(k) * a_k16, bd + static_cast<uint32_t>(k) * b_k16, idesc, (kb | k) != 0 ? 1u : 0u);
          if (p.cn > 1) umma_commit_mcast(&empty[stage], cmask);
          else umma_commit(&empty[stage]);
          if (kb == nk - 1) umma_commit(accf);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(&full[stage], phase, 2);
      if (kb == 0) P3D_STAMP(3);
      if (kb == 4) P3D_STAMP(11);
      if (kb == 8) P3D_STAMP(12);
      if (kb == 12) P3D_STAMP(13);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aa = a_base + stage * A_SLOT, bb = b_base + stage * B_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = p.a_mn ? desc_mn(aa + k * a_step) : desc_k(aa + k * a_step);
          const uint64_t bd = p.b_mn ? desc_mn(bb + k * b_step) : desc_k(bb + k * b_step);
          if constexpr (cg == 2) umma_bf16_ss_2sm(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          else umma_bf16_ss(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if constexpr (cg == 2) {
          umma_commit_2sm(&empty[stage], 0x3);      // frees the slot in both CTAs of the pair
          if (kb == nk - 1) umma_commit_2sm(accf, 0x3);
        } else {
          if (p.cn > 1) umma_commit_mcast(&empty[stage], static_cast<uint16_t>((1u << p.cn) - 1u));
          else umma_commit(&empty[stage]);
          if (kb == nk - 1) umma_commit(accf);
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    }
    P3D_STAMP(4);
  } else {
    // ------------------------------------------------------------ epilogue: warp w owns TMEM lanes 32*(w%4)..+31
    // tcgen05.ld hands every lane one ROW (32 consecutive columns).  Storing that directly makes each instruction
    // touch 32 different lines, 16 bytes each.  Instead the 32 x 32 chunk is transposed through a padded smem tile
    // (the operand ring is free once the accumulator is complete) so that a lane owns 4 COLUMNS x 8 rows: every
    // global access (C, residual, bf16 output, fp32 reductions) is then 4 full 128-byte lines per instruction, bias
    // is a per-lane constant and the column sums need 2 shuffle steps instead of a 31-shuffle butterfly.
    if constexpr (OUT == 3 || OUT == 4) {
      // ------------------------------------------------------------ small-batch fused training epilogues
      // M <= 128: this CTA holds ALL rows of its columns, so the batch statistics of BatchNorm (forward) and the
      // column sums of its backward pass are CTA-local: two passes over the TMEM accumulator with a named barrier
      // in between replace the GEMM + statistics + finalize + activation kernels (3 launches -> 1, both directions).
      const int ew = warp & 3, half = (warp - 2) >> 2;
      const int hw = p.bn >= 64 ? p.bn / 2 : 32;
      const int cbeg = half * hw, cend = (cbeg + hw < p.bn) ? cbeg + hw : p.bn;
      const train::StepScalars* sc = static_cast<const train::StepScalars*>(p.ft.sc);
      grid_dependency_wait();
      const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f);
      const int et = (warp - 2) * 32 + lane;
      for (int j = et; j < p.bn; j += EPI_THREADS) sbias[j] = (p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      named_bar_sync(1, EPI_THREADS);
      constexpr int TP = 36;
      const uint32_t tile_s = smem_u32(smem) + (warp - 2) * 32 * TP * 4;
      const uint32_t sbias_s = smem_u32(sbias);
      const int rg = lane >> 3, cq = (lane & 7) * 4;
      const int mrow0 = ew * 32 + rg;                          // m0 == 0
      const float keep = sc->keep, inv_keep = sc->inv_keep;
      mbar_wait(accf, 0, 3);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      // accumulator chunk -> this lane's 4 columns x 8 rows (rows mrow0 + 4 i), as alpha * acc + bias
      auto chunk = [&](int c0, float (&o)[8][4]) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) sts128(tile_s + (lane * TP + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const float4 b4 = lds128(sbias_s + (c0 + cq) * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = lds128(tile_s + ((i * 4 + rg) * TP + cq) * 4);
          o[i][0] = alpha * t.x + b4.x; o[i][1] = alpha * t.y + b4.y; o[i][2] = alpha * t.z + b4.z; o[i][3] = alpha * t.w + b4.w;
        }
      };
      auto publish = [&](int c0, float (&s1)[4], float (&s2)[4]) {     // fold the 4 row groups, quadrant partial -> smem
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);  s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
        if (rg == 0) {
          float* q1 = scol + ew * 512 + c0 + cq;
          *reinterpret_cast<float4*>(q1) = make_float4(s1[0], s1[1], s1[2], s1[3]);
          *reinterpret_cast<float4*>(q1 + 256) = make_float4(s2[0], s2[1], s2[2], s2[3]);
        }
      };
      auto totals = [&](int c0, float (&S1)[4], float (&S2)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + cq + j;
          S1[j] = (scol[c] + scol[512 + c]) + (scol[1024 + c] + scol[1536 + c]);
          S2[j] = (scol[256 + c] + scol[768 + c]) + (scol[1280 + c] + scol[1792 + c]);
        }
      };
      const bool has_bn = p.ft.has_bn != 0, dropout = p.ft.dropout != 0;
      const float invB = p.ft.invB;
      if constexpr (OUT == 3) {
        // ---- forward: z = alpha acc + bias; batch statistics; BN, ReLU, dropout, residual
        if (has_bn) {
          for (int c0 = cbeg; c0 < cend && n0 + c0 < p.N; c0 += 32) {
            float o[8][4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            chunk(c0, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool live = mrow0 + 4 * i < p.M;
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float q = live ? o[i][j] : 0.f; s1[j] += q; s2[j] += q * q; }
            }
            publish(c0, s1, s2);
          }
        }
        named_bar_sync(1, EPI_THREADS);
        for (int c0 = cbeg; c0 < cend && n0 + c0 < p.N; c0 += 32) {
          float o[8][4];
          chunk(c0, o);
          const int n = n0 + c0 + cq;
          float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {1.f, 1.f, 1.f, 1.f}, ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
          if (has_bn) {
            float S1[4], S2[4];
            totals(c0, S1, S2);
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ft.gamma + n)), b4 = __ldg(reinterpret_cast<const float4*>(p.ft.beta + n));
            ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w; be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
            float var[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              mu[j] = S1[j] * invB;
              var[j] = fmaxf(S2[j] * invB - mu[j] * mu[j], 0.f);        // biased, as TF's non-fused path
              rs[j] = 1.f / sqrtf(var[j] + kBnEps);
            }
            if (ew == 0 && rg == 0) {                                   // one writer per column
              *reinterpret_cast<float4*>(p.ft.mean + n) = make_float4(mu[0], mu[1], mu[2], mu[3]);
              *reinterpret_cast<float4*>(p.ft.rstd + n) = make_float4(rs[0], rs[1], rs[2], rs[3]);
              float4 mm = *reinterpret_cast<const float4*>(p.ft.mov_mean + n), mv = *reinterpret_cast<const float4*>(p.ft.mov_var + n);
              mm.x = mm.x * kBnMomentum + mu[0] * (1.f - kBnMomentum); mm.y = mm.y * kBnMomentum + mu[1] * (1.f - kBnMomentum);
              mm.z = mm.z * kBnMomentum + mu[2] * (1.f - kBnMomentum); mm.w = mm.w * kBnMomentum + mu[3] * (1.f - kBnMomentum);
              mv.x = mv.x * kBnMomentum + var[0] * (1.f - kBnMomentum); mv.y = mv.y * kBnMomentum + var[1] * (1.f - kBnMomentum);
              mv.z = mv.z * kBnMomentum + var[2] * (1.f - kBnMomentum); mv.w = mv.w * kBnMomentum + var[3] * (1.f - kBnMomentum);
              *reinterpret_cast<float4*>(p.ft.mov_mean + n) = mm;
              *reinterpret_cast<float4*>(p.ft.mov_var + n) = mv;
            }
          }
          float4 hr[8];
          uchar4 mi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            const size_t off = static_cast<size_t>(m < p.M ? m : 0) * p.N + n;
            hr[i] = p.ft.hres ? __ldg(reinterpret_cast<const float4*>(p.ft.hres + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            mi[i] = (dropout && p.ft.mask_in) ? *reinterpret_cast<const uchar4*>(p.ft.mask_in + off) : make_uchar4(1, 1, 1, 1);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            if (m >= p.M) continue;
            const size_t off = static_cast<size_t>(m) * p.N + n;
            *reinterpret_cast<float4*>(p.C + off) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);      // z, for the backward pass
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = has_bn ? ga[j] * ((o[i][j] - mu[j]) * rs[j]) + be[j] : o[i][j];
              r[j] = fmaxf(a, 0.f);
            }
            if (dropout) {
              uint8_t kb[4] = {mi[i].x, mi[i].y, mi[i].z, mi[i].w};
              if (!p.ft.mask_in) {
                const uint4 w4 = train::dropout_words(sc->seed, sc->step, static_cast<uint32_t>(p.ft.layer), static_cast<uint32_t>(m), static_cast<uint32_t>(n >> 2));
                kb[0] = train::keep_bit(w4.x, keep); kb[1] = train::keep_bit(w4.y, keep); kb[2] = train::keep_bit(w4.z, keep); kb[3] = train::keep_bit(w4.w, keep);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) r[j] = kb[j] ? r[j] * inv_keep : 0.f;
              *reinterpret_cast<uchar4*>(p.ft.mask + off) = make_uchar4(kb[0], kb[1], kb[2], kb[3]);
            }
            r[0] += hr[i].x; r[1] += hr[i].y; r[2] += hr[i].z; r[3] += hr[i].w;
            *reinterpret_cast<float4*>(p.ft.h + off) = make_float4(r[0], r[1], r[2], r[3]);
            __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.ft.hb) + off) = pk;
          }
        }
      } else {
        // ---- backward: dh = alpha acc (+ res); da = dh * dropout * relu'(act); BN backward with CTA-local sums
        auto da_of = [&](int c0, float (&o)[8][4], float (&xh)[8][4], const float (&mu)[4], const float (&rs)[4], const float (&ga)[4],
                         const float (&be)[4], bool store_dh) {
          const int n = n0 + c0 + cq;
          // all global operands of the 8 rows first (independent loads in flight), then the arithmetic
          float4 z4[8], r4[8];
          uchar4 mk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            const size_t off = static_cast<size_t>(m < p.M ? m : 0) * p.N + n;
            z4[i] = __ldg(reinterpret_cast<const float4*>(p.ft.z + off));
            mk[i] = dropout ? *reinterpret_cast<const uchar4*>(p.ft.mask + off) : make_uchar4(1, 1, 1, 1);
            r4[i] = p.res ? __ldg(reinterpret_cast<const float4*>(p.res + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            const bool live = m < p.M;
            o[i][0] += r4[i].x; o[i][1] += r4[i].y; o[i][2] += r4[i].z; o[i][3] += r4[i].w;
            if (store_dh && p.ft.dh_out && live)
              *reinterpret_cast<float4*>(p.ft.dh_out + static_cast<size_t>(m) * p.N + n) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
            const float zz[4] = {z4[i].x, z4[i].y, z4[i].z, z4[i].w};
            const unsigned char kk[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              xh[i][j] = has_bn ? (zz[j] - mu[j]) * rs[j] : 0.f;
              const float act = has_bn ? ga[j] * xh[i][j] + be[j] : zz[j];
              float gr = o[i][j];
              if (dropout) gr = kk[j] ? gr * inv_keep : 0.f;
              o[i][j] = (live && act > 0.f) ? gr : 0.f;            // da
            }
          }
        };
        auto bn_consts = [&](int c0, float (&mu)[4], float (&rs)[4], float (&ga)[4], float (&be)[4]) {
          const int n = n0 + c0 + cq;
#pragma unroll
          for (int j = 0; j < 4; ++j) { mu[j] = 0.f; rs[j] = 1.f; ga[j] = 1.f; be[j] = 0.f; }
          if (has_bn) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.ft.mean + n)), r4 = __ldg(reinterpret_cast<const float4*>(p.ft.rstd + n));
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ft.gamma + n)), b4 = __ldg(reinterpret_cast<const float4*>(p.ft.beta + n));
            mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w; rs[0] = r4.x; rs[1] = r4.y; rs[2] = r4.z; rs[3] = r4.w;
            ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w; be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
          }
        };
        for (int c0 = cbeg; c0 < cend && n0 + c0 < p.N; c0 += 32) {
          float o[8][4], xh[8][4], mu[4], rs[4], ga[4], be[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
          chunk(c0, o);
          bn_consts(c0, mu, rs, ga, be);
          da_of(c0, o, xh, mu, rs, ga, be, true);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += o[i][j]; s2[j] += o[i][j] * xh[i][j]; }
          publish(c0, s1, s2);
        }
        named_bar_sync(1, EPI_THREADS);
        for (int c0 = cbeg; c0 < cend && n0 + c0 < p.N; c0 += 32) {
          float o[8][4], xh[8][4], mu[4], rs[4], ga[4], be[4], P[4], Q[4];
          chunk(c0, o);
          bn_consts(c0, mu, rs, ga, be);
          da_of(c0, o, xh, mu, rs, ga, be, false);
          totals(c0, P, Q);
          const int n = n0 + c0 + cq;
          if (ew == 0 && rg == 0) {                                     // one writer per column
            if (has_bn) {
              *reinterpret_cast<float4*>(p.ft.gbeta + n) = make_float4(P[0], P[1], P[2], P[3]);
              *reinterpret_cast<float4*>(p.ft.ggamma + n) = make_float4(Q[0], Q[1], Q[2], Q[3]);
            } else {
              *reinterpret_cast<float4*>(p.ft.gbias + n) = make_float4(P[0], P[1], P[2], P[3]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            if (m >= p.M) continue;
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = has_bn ? ga[j] * rs[j] * (o[i][j] - P[j] * invB - xh[i][j] * (Q[j] * invB)) : o[i][j];
            __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.ft.dzb) + static_cast<size_t>(m) * p.N + n) = pk;
          }
        }
      }
    } else {
    // The epilogue variant is a template parameter and the bias sits in smem before the accumulator is complete:
    // the first version (runtime flags, generic smem accesses, bias re-loaded per chunk) spent 11 of the 20 us of a
    // 4096 x 1024 x 1024 GEMM here, one warp per scheduler crawling through ~2800 SASS instructions of branches.
    const int ew = warp & 3;                     // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;            // which half of the tile's columns
    const int hw = p.bn >= 64 ? p.bn / 2 : 32;                       // columns per half (BN = 32: the second half idles)
    const int cbeg = half * hw, cend = (cbeg + hw < p.bn) ? cbeg + hw : p.bn;
    const bool first_split = (blockIdx.z == 0);
    grid_dependency_wait();       // residual / alpha / C written by earlier kernels
    const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f);
    const float inv_keep2 = (CS == 2 && p.ft.dropout) ? static_cast<const train::StepScalars*>(p.ft.sc)->inv_keep : 1.f;
    const int et = (warp - 2) * 32 + lane;       // 0..255
    for (int j = et; j < p.bn; j += EPI_THREADS) {
      sbias[j] = (first_split && p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
    }
    named_bar_sync(1, EPI_THREADS);
    const uint32_t sbias_s = smem_u32(sbias);
    if constexpr (TS) {
      // ---------------------------------------------------------- TMA-store epilogue (OUT 0 / 1, no residual, CS 0 / 1)
      // tcgen05.ld hands every lane one ROW of a 32 x 32 chunk; that is exactly how a TMA box lies in shared memory, so
      // the chunk is written row-per-lane into a dense 32 x 128 B tile - 16-byte pieces XOR-swizzled by the row (the
      // SWIZZLE_128B pattern of tm_c: conflict-free for the per-lane st.shared.v4 AND for the column reads of the
      // statistics below) - and one lane sends it off.  No transpose, no per-lane global store, ragged M / N edges are
      // clipped by the tensor map.  Two staging tiles per warp: chunk i+1 is written while the TMA engine still reads
      // chunk i (cp.async.bulk.wait_group.read 1).  Column sums: lane = column, one conflict-free LDS per row.
      static_assert(OUT <= 1 && !RES && CS <= 1 && CG == 1 && OCC == 1, "TS epilogue: fp32 store / reduction without residual only");
      uint8_t* stg = smem + (warp - 2) * 8192;               // 2 x 4 KB per warp in the freed operand ring (1024-byte aligned)
      const uint32_t stg_s = smem_u32(stg);
      const int row_g = m0 + ew * 32;                        // first global row of this warp's TMEM lane quadrant
      const int rows_live = (p.M - row_g) < 32 ? (p.M - row_g) : 32;   // <= 0: the whole quadrant lies below the matrix
      mbar_wait(accf, 0, 3);
      if (warp == 2) P3D_STAMP(5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      uint32_t v[32];
      if (cbeg < cend && n0 + cbeg < p.N) tmem_ld_32x32b_x32(taddr + cbeg, v);
      int buf = 0;
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        if (n0 + c0 >= p.N) break;     // warp-uniform
        const bool more = (c0 + 32 < cend) && (n0 + c0 + 32 < p.N);
        tmem_ld_wait();
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = lds128(sbias_s + (c0 + 4 * j) * 4);      // same address in every lane: broadcast
          o[4 * j + 0] = __float_as_uint(alpha * __uint_as_float(v[4 * j + 0]) + b4.x);
          o[4 * j + 1] = __float_as_uint(alpha * __uint_as_float(v[4 * j + 1]) + b4.y);
          o[4 * j + 2] = __float_as_uint(alpha * __uint_as_float(v[4 * j + 2]) + b4.z);
          o[4 * j + 3] = __float_as_uint(alpha * __uint_as_float(v[4 * j + 3]) + b4.w);
        }
        if (more) tmem_ld_32x32b_x32(taddr + c0 + 32, v);      // in flight while this chunk is staged and sent
        if (lane == 0) tma_store_wait_read<1>();               // the store issued two chunks ago has read this staging tile
        __syncwarp();                                          // ... and every lane is done with its column reads of it
        const uint32_t tb = stg_s + buf * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(tb + lane * 128 + ((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async_smem();                              // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0 && rows_live > 0) {
          if (OUT == 1) tma_reduce_add_2d(&tm_c, stg + buf * 4096, n0 + c0, row_g);   // split-K / accumulate: C += tile, added in the L2
          else tma_store_2d(&tm_c, stg + buf * 4096, n0 + c0, row_g);
          tma_store_commit();
        }
        if (CS == 1) {
          // lane = column c0 + lane: piece (lane / 4) of row r sits at piece index (lane / 4) ^ (r % 8) -> the 32 lanes of
          // one row read hit 32 different banks.  Rows >= M hold bias (or stale operand rows of a short tile): skipped.
          float s1 = 0.f, s2 = 0.f;
          const uint32_t cb = tb + (lane & 3) * 4;
          const uint32_t piece = static_cast<uint32_t>(lane >> 2);
#pragma unroll 8
          for (int r = 0; r < rows_live; ++r) {
            const float q = lds32(cb + r * 128 + ((piece ^ static_cast<uint32_t>(r & 7)) << 4));
            s1 += q; s2 += q * q;
          }
          scol[ew * 512 + c0 + lane] = s1;                     // this quadrant's partial sums (one writer per slot)
          scol[ew * 512 + 256 + c0 + lane] = s2;
        }
        buf ^= 1;
      }
      if (lane == 0) tma_store_wait<0>();                      // all of this warp's stores are complete before the CTA retires
    } else {
    constexpr int TP = 36;                                   // tile pitch in floats (144 B): conflict-free both ways
    const uint32_t tile_s = smem_u32(smem) + (warp - 2) * 32 * TP * 4;
    const int rg = lane >> 3, cq = (lane & 7) * 4;           // row group (rows rg, rg+4, ...), first of this lane's 4 columns
    const int mrow0 = m0 + ew * 32 + rg;                     // this lane's rows: mrow0 + 4 i
    const bool use_res = RES && (OUT == 2 || first_split);
    // fast path: the whole tile is inside N and every pointer/pitch allows vector accesses
    bool fast = (n0 + p.bn <= p.N);
    if (OUT == 2) fast = fast && ((p.ldob & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_b) & 7) == 0) &&
                         (!RES || (((p.ldrb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res_b) & 7) == 0)));
    else fast = fast && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!RES || (((p.ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)));
    mbar_wait(accf, 0, 3);
    if (warp == 2) P3D_STAMP(5);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    uint32_t v[32];
    if (cbeg < cend && n0 + cbeg < p.N) tmem_ld_32x32b_x32(taddr + cbeg, v);
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      if (n0 + c0 >= p.N) break;     // warp-uniform
      const bool more = (c0 + 32 < cend) && (n0 + c0 + 32 < p.N);
      tmem_ld_wait();
      __syncwarp();                  // the previous chunk's readers are done with the tile
#pragma unroll
      for (int j = 0; j < 32; j += 4) sts128(tile_s + (lane * TP + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      if (more) tmem_ld_32x32b_x32(taddr + c0 + 32, v);      // in flight while this chunk is processed
      const int n = n0 + c0 + cq;                            // this lane's columns n .. n+3
      const float4 b4 = lds128(sbias_s + (c0 + cq) * 4);
      float cs1[4] = {0.f, 0.f, 0.f, 0.f}, cs2[4] = {0.f, 0.f, 0.f, 0.f};
      if (fast) {
        // CS == 2: this chunk's pre-activations / keep-mask and the per-column BatchNorm constants; none of them
        // depends on the accumulator, so the loads are in flight while the tile is read back from shared memory
        float4 z4[CS == 2 ? 8 : 1];
        uchar4 mk[CS == 2 ? 8 : 1];
        float4 bmu, brs, bga, bbe;
        if (CS == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            z4[i] = make_float4(0.f, 0.f, 0.f, 0.f); mk[i] = make_uchar4(1, 1, 1, 1);
            if (m < p.M) {
              z4[i] = __ldg(reinterpret_cast<const float4*>(p.ft.z + static_cast<size_t>(m) * p.N + n));
              if (p.ft.dropout) mk[i] = *reinterpret_cast<const uchar4*>(p.ft.mask + static_cast<size_t>(m) * p.N + n);
            }
          }
          bmu = __ldg(reinterpret_cast<const float4*>(p.ft.mean + n)); brs = __ldg(reinterpret_cast<const float4*>(p.ft.rstd + n));
          bga = __ldg(reinterpret_cast<const float4*>(p.ft.gamma + n)); bbe = __ldg(reinterpret_cast<const float4*>(p.ft.beta + n));
        }
        float4 t4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t4[i] = lds128(tile_s + ((i * 4 + rg) * TP + cq) * 4);
        float4 r4[8];
        uint2 rb[8];
        if (use_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            r4[i] = make_float4(0.f, 0.f, 0.f, 0.f); rb[i] = make_uint2(0u, 0u);
            if (m < p.M) {
              if (OUT == 2) rb[i] = *reinterpret_cast<const uint2*>(p.res_b + static_cast<size_t>(m) * p.ldrb + n);
              else r4[i] = __ldg(reinterpret_cast<const float4*>(p.res + static_cast<size_t>(m) * p.ldres + n));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = mrow0 + 4 * i;
          const bool live = m < p.M;
          float o[4] = {alpha * t4[i].x + b4.x, alpha * t4[i].y + b4.y, alpha * t4[i].z + b4.z, alpha * t4[i].w + b4.w};
          if (CS == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float q = live ? o[j] : 0.f; cs1[j] += q; cs2[j] += q * q; }
          }
          if (OUT == 2) {
            // bf16 activations of the layered inference path: same rounding points as the fused persistent kernel
            // (relu(.) rounded to bf16, then the residual added and rounded again)
            __nv_bfloat16 ob[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ob[j] = __float2bfloat16_rn(p.relu ? fmaxf(o[j], 0.f) : o[j]);
            if (use_res) {
              __nv_bfloat16 rv[4];
              memcpy(rv, &rb[i], 8);
#pragma unroll
              for (int j = 0; j < 4; ++j) ob[j] = __float2bfloat16_rn(__bfloat162float(ob[j]) + __bfloat162float(rv[j]));
            }
            uint2 q; memcpy(&q, ob, 8);
            if (live) *reinterpret_cast<uint2*>(p.out_b + static_cast<size_t>(m) * p.ldob + n) = q;
          } else {
            if (use_res) { o[0] += r4[i].x; o[1] += r4[i].y; o[2] += r4[i].z; o[3] += r4[i].w; }
            if (CS == 2) {
              // o = dh of this hidden layer (the gradient that bypassed the block included): pass A of the BatchNorm backward
              const float zz[4] = {z4[i].x, z4[i].y, z4[i].z, z4[i].w};
              const unsigned char kk[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
              const float mu[4] = {bmu.x, bmu.y, bmu.z, bmu.w}, rs[4] = {brs.x, brs.y, brs.z, brs.w};
              const float ga[4] = {bga.x, bga.y, bga.z, bga.w}, be[4] = {bbe.x, bbe.y, bbe.z, bbe.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float xh = (zz[j] - mu[j]) * rs[j];
                const float act = ga[j] * xh + be[j];
                float gr = o[j];
                if (p.ft.dropout) gr = kk[j] ? gr * inv_keep2 : 0.f;
                const float da = (live && act > 0.f) ? gr : 0.f;
                cs1[j] += da; cs2[j] += da * xh;
              }
            }
            float* crow = p.C + static_cast<size_t>(m) * p.ldc + n;
            if (live) {
              if (OUT == 1) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
              else *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      } else {
        // ragged / unaligned tile (N = 48 or 42 output layer, odd pitches): element by element; rows not unrolled.
        // The column loop IS unrolled: with a runtime column index, bias / tv and - shared with the fast path above -
        // cs1 / cs2 were demoted to local memory (ncu: STL/LDL of the column sums after every row of the hot loop).
        const float bias[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int m = mrow0 + 4 * i;
          if (m >= p.M) continue;
          const float4 t = lds128(tile_s + ((i * 4 + rg) * TP + cq) * 4);
          const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= p.N) continue;
            float o = alpha * tv[j] + bias[j];
            if (CS == 1) { cs1[j] += o; cs2[j] += o * o; }     // CS == 2 never takes the ragged path (checked by plan())
            if (OUT == 2) {
              __nv_bfloat16 ob = __float2bfloat16_rn(p.relu ? fmaxf(o, 0.f) : o);
              if (use_res) ob = __float2bfloat16_rn(__bfloat162float(ob) + __bfloat162float(p.res_b[static_cast<size_t>(m) * p.ldrb + n + j]));
              p.out_b[static_cast<size_t>(m) * p.ldob + n + j] = ob;
            } else {
              if (use_res) o += __ldg(p.res + static_cast<size_t>(m) * p.ldres + n + j);
              float* c = p.C + static_cast<size_t>(m) * p.ldc + n + j;
              if (OUT == 1) atomicAdd(c, o); else *c = o;
            }
          }
        }
      }
      if (CS) {
        // this lane holds 8 of the warp's 32 rows for its 4 columns: fold the 4 row groups, then lanes 0..7 publish
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cs1[j] += __shfl_xor_sync(0xffffffffu, cs1[j], 8);  cs2[j] += __shfl_xor_sync(0xffffffffu, cs2[j], 8);
          cs1[j] += __shfl_xor_sync(0xffffffffu, cs1[j], 16); cs2[j] += __shfl_xor_sync(0xffffffffu, cs2[j], 16);
        }
        if (rg == 0) {            // this quadrant's partial sums -> smem (plain stores: one writer per slot)
          float* q1 = scol + ew * 512 + c0 + cq;
          *reinterpret_cast<float4*>(q1) = make_float4(cs1[0], cs1[1], cs1[2], cs1[3]);
          *reinterpret_cast<float4*>(q1 + 256) = make_float4(cs2[0], cs2[1], cs2[2], cs2[3]);
        }
      }
    }
    }   // !TS
    if (CS) {
      named_bar_sync(1, EPI_THREADS);
      for (int j = et; j < p.bn; j += EPI_THREADS) {        // one global fp64 atomic per column per CTA
        if (n0 + j < p.N) {
          const float s1 = (scol[j] + scol[512 + j]) + (scol[1024 + j] + scol[1536 + j]);
          const float s2 = (scol[256 + j] + scol[768 + j]) + (scol[1280 + j] + scol[1792 + j]);
          atomicAdd(p.colsum + n0 + j, static_cast<double>(s1));
          atomicAdd(p.colsum + p.N + n0 + j, static_cast<double>(s2));
        }
      }
    }
    }   // plain epilogues
  }
  if (warp == 2) P3D_STAMP(6);
  tc_fence_before();
  if (p.cn > 1 || cg == 2) cluster_sync(); else __syncthreads();     // no CTA may leave while a peer can still arrive on its barriers
  if (warp == 0) P3D_STAMP(7);
  if (warp == 1) {
    tc_fence_after();
    if constexpr (cg == 2) tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.bn)); else tmem_dealloc(tmem_base, static_cast<uint32_t>(p.bn));
  }
}

// ----------------------------------------------------------------------------- MMA issue-rate probe
// One CTA, operands resident in shared memory (zeros), `iters` back-to-back tcgen05.mma 128 x N x 16 into one
// accumulator, timed with clock64 from the first issue to the completion of the commit: cycles per MMA as a function
// of N (what a k-block of four MMAs costs when nothing else is in the way).
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, N);
    const uint64_t ad = desc_k(smem_u32(smem)), bd = desc_k(smem_u32(smem + 16 * 1024));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_bf16_ss(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i ? 1u : 0u);
    umma_commit(bar);
    const long long t1 = clock64();
    mbar_wait(bar, 0, 9);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int mma_rate(int N, int iters, long long* out_dev, cudaStream_t st) {
  P3D_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && iters >= 1, "mma_rate: bad argument");
  const int smem = 1024 + 48 * 1024 + 64;
  P3D_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<1, 128, smem, st>>>(N, iters, out_dev);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// bf16 row-major [outer][inner] (pitch in elements), box [box_outer][64], 128B swizzle, zero fill out of range
static int make_map(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return P3D_ERR_CUDA; }
  if ((pitch * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tc_gemm: operand pitch/base must be 16-byte aligned (pitch %llu elements)", (unsigned long long)pitch);
    return P3D_ERR_ARG;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return P3D_ERR_CUDA; }
  return P3D_OK;
}

// fp32 row-major C [M][N] (pitch ldc elements) for the TS epilogue: box 32 columns (128 B) x 32 rows, 128B swizzle;
// stores / reductions beyond M or N are clipped by the TMA engine
static int make_map_c(CUtensorMap* out, const float* base, uint64_t N, uint64_t M, uint64_t ldc) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return P3D_ERR_CUDA; }
  cuuint64_t gdim[2] = {N, M};
  cuuint64_t gstride[1] = {ldc * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (C) failed with CUresult %d", (int)r); return P3D_ERR_CUDA; }
  return P3D_OK;
}

// A: a_mn ? [K][M] : [M][K];  B: b_mn ? [K][N] : [N][K]  (bf16, pitches lda/ldb in elements)
struct PlanData { CUtensorMap ta, tb, tc; Params p; dim3 grid; int pdl; int fused_mode; };
static_assert(sizeof(PlanData) <= sizeof(GemmPlan::blob), "GemmPlan::blob too small");

int plan(const GemmArgs& g, GemmPlan* out) {
  P3D_REQUIRE(g.M >= 1 && g.N >= 1 && g.K >= 1 && g.A && g.B && (g.C || g.out_bf16), "tc_gemm: bad argument");
  const int num_sms = 148;
  const int mt = (g.M + BM - 1) / BM;
  const int n64 = (g.N + 63) / 64 * 64;
  const int kblocks = (g.K + BK - 1) / BK;
  // tile width: the widest BN that still gives the grid about one CTA per SM; a split-K problem keeps the
  // wide tile (better MMA efficiency) and fills the machine through gridDim.z instead
  int bn = 256;
  while (bn > 64 && (bn > n64 || (!g.split_k && mt * ((g.N + bn - 1) / bn) < 100))) bn >>= 1;
  // a split-K problem whose K is only a few blocks long (weight gradients of a small batch) is all epilogue:
  // narrow tiles spread it over the machine instead
  if (g.split_k && kblocks <= 4) { while (bn > 64 && mt * ((g.N + bn - 1) / bn) < 100) bn >>= 1; }
  // one short M tile (small-batch inference): 32-wide tiles put twice as many SMs on the weight stream
  if (bn == 64 && !g.b_mn && !g.split_k && mt == 1 && g.N >= 256 && (g.N % 32) == 0) bn = 32;
  // Two CTAs per SM (P3D_GEMM_OCC2=1, opt-in until measured): for problems with more than one 128-wide tile per SM,
  // 128 x 128 tiles with a 96 KB ring; the second resident CTA hides the first one's prologue and epilogue.
  static const bool occ2_env = [] { const char* e = getenv("P3D_GEMM_OCC2"); return e && e[0] == '1'; }();
  int occ = 1;
  const int tiles128 = mt * ((g.N + 127) / 128);
  if (occ2_env && !g.pdl && !g.fused_mode && !g.out_bf16 && bn >= 128 &&
      (g.split_k ? (kblocks >= 16 && tiles128 >= 32) : (tiles128 > num_sms))) {
    occ = 2;
    bn = 128;
  }
  const int nt = (g.N + bn - 1) / bn;
  int splits = 1;
  if (g.split_k) {
    splits = occ * num_sms / (mt * nt);
    if (splits < 1) splits = 1;
    if (splits > kblocks) splits = kblocks;
    if (splits > 32) splits = 32;
  }
  const int kps = ((kblocks + splits - 1) / splits) * BK;
  splits = (g.K + kps - 1) / kps;
  PlanData* d = reinterpret_cast<PlanData*>(out->blob);
  // K-major A of a short problem: fetch only the rows that exist (the MMA still reads 128 smem rows; what it makes
  // of the stale ones lands in accumulator rows >= M, which are never stored)
  // The CTAs of one M tile (adjacent N tiles) read the same A tile: in a cluster of cn of them each fetches 1/cn of
  // the rows and TMA-multicasts it to all.  Measured on B200 (M=64, N=K=1024): the mainloop got SLOWER (5.7 -> 6.2 us)
  // although every CTA fetches half the bytes - a k-block costs ~0.2 us + 0.5 ns per 128-byte row whatever its size,
  // so bytes are not what bounds a short-M GEMM.  Kept behind P3D_GEMM_MCAST=1 for the large-M experiments of round 2.
  int cn = 1;
  static const bool mcast = [] { const char* e = getenv("P3D_GEMM_MCAST"); return e && e[0] == '1'; }();
  if (!g.a_mn && mcast) cn = (nt % 4 == 0) ? 4 : ((nt % 2 == 0) ? 2 : 1);
  // CTA pairs along M (cta_group::2): a 256 x BN tile per pair, each CTA fetches its own A rows and half of the B tile,
  // i.e. 32 KB instead of 48 KB per k-block at BN = 256.  The large-M GEMMs of the training step are bound by the
  // L2 -> SM operand stream, which this cuts by a third.  Opt-in (P3D_GEMM_CG2=1) until measured.
  static const bool pair_env = [] { const char* e = getenv("P3D_GEMM_CG2"); return e && e[0] == '1'; }();
  int cg = 1;
  if (occ == 2) cn = 1;
  if (pair_env && occ == 1 && cn == 1 && !g.pdl && !g.fused_mode && !g.out_bf16 && g.M >= 2 * BM && bn >= 128) cg = 2;   // the fp32-output epilogues have pair instantiations
  int a_rows = (!g.a_mn && g.M < BM) ? ((g.M + 7) / 8 * 8) : BM;
  if (cn > 1) a_rows = (a_rows + 8 * cn - 1) / (8 * cn) * (8 * cn);     // every share is whole 8-row swizzle groups
  if (!g.a_mn) P3D_TRY(make_map(&d->ta, g.A, g.K, g.M, g.lda, a_rows / cn));
  else P3D_TRY(make_map(&d->ta, g.A, g.M, g.K, g.lda, BK));
  if (!g.b_mn) P3D_TRY(make_map(&d->tb, g.B, g.K, g.N, g.ldb, bn / cg));
  else P3D_TRY(make_map(&d->tb, g.B, g.N, g.K, g.ldb, BK));
  Params& p = d->p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.bn = bn; p.k_per_split = kps; p.a_mn = g.a_mn; p.b_mn = g.b_mn;
  p.C = g.C; p.ldc = g.ldc; p.bias = g.bias; p.res = g.res; p.ldres = g.ldres; p.alpha_dev = g.alpha_dev; p.alpha = g.alpha;
  p.atomic = (splits > 1 || g.accumulate) ? 1 : 0;
  p.colsum = g.colsum;
  p.out_b = static_cast<__nv_bfloat16*>(g.out_bf16); p.ldob = g.ld_out_bf16;
  p.res_b = static_cast<const __nv_bfloat16*>(g.res_bf16); p.ldrb = g.ld_res_bf16; p.relu = g.relu;
  p.cn = cn;
  p.cg = cg;
  p.occ = occ;
  // TMA-store epilogue (P3D_GEMM_TMASTORE=1, opt-in until measured on the GPU): fp32 C without a residual operand, pitch and
  // base such that a tensor map can describe it; everything else keeps the st.global epilogues
  static const bool ts_env = [] { const char* e = getenv("P3D_GEMM_TMASTORE"); return e && e[0] == '1'; }();
  p.ts = (ts_env && g.C && !g.out_bf16 && !g.res && !g.fused_mode && cg == 1 && occ == 1 && cn == 1 && (g.ldc % 4) == 0 &&
          (reinterpret_cast<uintptr_t>(g.C) & 15) == 0) ? 1 : 0;
  // Lean MMA issue path (P3D_GEMM_FASTISSUE=1, opt-in until measured on the GPU): single-CTA, one-CTA-per-SM kernels
  static const bool fi_env = [] { const char* e = getenv("P3D_GEMM_FASTISSUE"); return e && e[0] == '1'; }();
  p.fi = (fi_env && cg == 1 && occ == 1 && g.fused_mode != 5) ? 1 : 0;
  memset(&d->tc, 0, sizeof(d->tc));
  if (p.ts) P3D_TRY(make_map_c(&d->tc, g.C, static_cast<uint64_t>(g.N), static_cast<uint64_t>(g.M), static_cast<uint64_t>(g.ldc)));
  p.a_bytes = a_rows * BK * 2;
  p.stages = (occ == 2 ? RING2_BYTES : RING_BYTES) / (p.a_bytes + (bn / cg) * BK * 2);
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.b_independent = g.pdl ? 1 : 0;
  p.dbg = static_cast<unsigned long long*>(g.dbg);
  d->pdl = g.pdl;
  p.ft = g.fused;
  d->fused_mode = g.fused_mode;
  if (g.fused_mode == 5) {
    // any batch: the dh-producing GEMM also accumulates the BatchNorm backward sums of the layer it feeds
    P3D_REQUIRE(g.colsum && g.C && !g.out_bf16 && splits == 1 && !g.accumulate, "tc_gemm: backward-sum epilogue needs colsum, fp32 C, unsplit K");
    P3D_REQUIRE((g.N % 256) == 0 && g.ldc == g.N && (!g.res || g.ldres == g.N) && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0,
                "tc_gemm: backward-sum epilogue needs N %% 256 == 0 and dense, aligned C / residual");
    P3D_REQUIRE(g.fused.z && g.fused.mean && g.fused.rstd && g.fused.gamma && g.fused.beta && g.fused.sc && (!g.fused.dropout || g.fused.mask),
                "tc_gemm: backward-sum epilogue operands missing");
  } else if (g.fused_mode) {
    P3D_REQUIRE(g.fused_mode == 3 || g.fused_mode == 4, "tc_gemm: unknown fused mode %d", g.fused_mode);
    P3D_REQUIRE(mt == 1 && splits == 1 && (g.N % 32) == 0 && g.ldc == g.N && !g.out_bf16 && !g.colsum,
                "tc_gemm: fused training epilogues need M <= 128, N %% 32 == 0, unsplit K, dense C");
    P3D_REQUIRE(!g.res || g.ldres == g.N, "tc_gemm: fused epilogue residual must be dense");
  }
  P3D_REQUIRE(!(p.colsum && splits > 1), "tc_gemm: column sums need an unsplit K");
  P3D_REQUIRE(!(p.out_b && splits > 1), "tc_gemm: bf16 output needs an unsplit K");
  if (cg == 2) d->grid = dim3((mt + 1) / 2 * 2, nt, splits);       // pairs: M tiles along x, an odd last one gets an all-out-of-range partner
  else d->grid = dim3(nt, mt, splits);
  out->valid = 1;
  return P3D_OK;
}

int launch(const GemmPlan& pl, cudaStream_t st) {
  P3D_REQUIRE(pl.valid, "tc_gemm: launch of an unplanned GEMM");
  const PlanData* d = reinterpret_cast<const PlanData*>(pl.blob);
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
  const Params& q = d->p;
  const int out = q.out_b ? 2 : (q.atomic ? 1 : 0);
  const bool res = q.out_b ? (q.res_b != nullptr) : (q.res != nullptr);
  const int cs = q.colsum ? (d->fused_mode == 5 ? 2 : 1) : 0;
  KernelFn fn = nullptr;
#define P3D_TCG_PICK(O, R, S) if (out == O && res == R && cs == S) fn = tc_gemm_kernel<O, R, S>;
#define P3D_TCG_PICK2(O, R, S) if (out == O && res == R && cs == S) \
    fn = q.cg == 2 ? tc_gemm_kernel<O, R, S, 2> : (q.occ == 2 ? tc_gemm_kernel<O, R, S, 1, 2> : tc_gemm_kernel<O, R, S, 1>);
  P3D_TCG_PICK2(0, false, 0) P3D_TCG_PICK2(0, false, 1) P3D_TCG_PICK2(0, true, 0) P3D_TCG_PICK2(0, true, 1)
  P3D_TCG_PICK(0, false, 2) P3D_TCG_PICK(0, true, 2)
  P3D_TCG_PICK2(1, false, 0) P3D_TCG_PICK2(1, true, 0)
  P3D_TCG_PICK(2, false, 0) P3D_TCG_PICK(2, true, 0)
#undef P3D_TCG_PICK
#undef P3D_TCG_PICK2
  if (q.ts) {      // plan() admits only these three combinations
    fn = nullptr;
    if (out == 0 && !res && cs == 0) fn = q.fi ? tc_gemm_kernel<0, false, 0, 1, 1, true, true> : tc_gemm_kernel<0, false, 0, 1, 1, true>;
    if (out == 0 && !res && cs == 1) fn = q.fi ? tc_gemm_kernel<0, false, 1, 1, 1, true, true> : tc_gemm_kernel<0, false, 1, 1, 1, true>;
    if (out == 1 && !res && cs == 0) fn = q.fi ? tc_gemm_kernel<1, false, 0, 1, 1, true, true> : tc_gemm_kernel<1, false, 0, 1, 1, true>;
  } else if (q.fi) {
#define P3D_TCG_PICKF(O, R, S) if (out == O && res == R && cs == S) fn = tc_gemm_kernel<O, R, S, 1, 1, false, true>;
    P3D_TCG_PICKF(0, false, 0) P3D_TCG_PICKF(0, false, 1) P3D_TCG_PICKF(0, true, 0) P3D_TCG_PICKF(0, true, 1)
    P3D_TCG_PICKF(1, false, 0) P3D_TCG_PICKF(1, true, 0) P3D_TCG_PICKF(2, false, 0) P3D_TCG_PICKF(2, true, 0)
#undef P3D_TCG_PICKF
  }
  if (d->fused_mode == 3) fn = q.fi ? tc_gemm_kernel<3, false, 0, 1, 1, false, true> : tc_gemm_kernel<3, false, 0>;
  if (d->fused_mode == 4) fn = q.fi ? tc_gemm_kernel<4, false, 0, 1, 1, false, true> : tc_gemm_kernel<4, false, 0>;
  P3D_REQUIRE(fn != nullptr, "tc_gemm: unsupported epilogue combination (out %d res %d colsum %d)", out, (int)res, (int)cs);
  static PerDeviceOnce attr;
  if (attr.needed()) {
    KernelFn all[] = {tc_gemm_kernel<0, false, 0>, tc_gemm_kernel<0, false, 1>, tc_gemm_kernel<0, true, 0>,
                      tc_gemm_kernel<0, true, 1>, tc_gemm_kernel<0, false, 2>, tc_gemm_kernel<0, true, 2>,
                      tc_gemm_kernel<1, false, 0>, tc_gemm_kernel<1, true, 0>,
                      tc_gemm_kernel<2, false, 0>, tc_gemm_kernel<2, true, 0>,
                      tc_gemm_kernel<3, false, 0>, tc_gemm_kernel<4, false, 0>,
                      tc_gemm_kernel<0, false, 0, 2>, tc_gemm_kernel<0, false, 1, 2>, tc_gemm_kernel<0, true, 0, 2>,
                      tc_gemm_kernel<0, true, 1, 2>, tc_gemm_kernel<1, false, 0, 2>, tc_gemm_kernel<1, true, 0, 2>,
                      tc_gemm_kernel<0, false, 0, 1, 1, true>, tc_gemm_kernel<0, false, 1, 1, 1, true>,
                      tc_gemm_kernel<1, false, 0, 1, 1, true>,
                      tc_gemm_kernel<0, false, 0, 1, 1, true, true>, tc_gemm_kernel<0, false, 1, 1, 1, true, true>,
                      tc_gemm_kernel<1, false, 0, 1, 1, true, true>,
                      tc_gemm_kernel<0, false, 0, 1, 1, false, true>, tc_gemm_kernel<0, false, 1, 1, 1, false, true>,
                      tc_gemm_kernel<0, true, 0, 1, 1, false, true>, tc_gemm_kernel<0, true, 1, 1, 1, false, true>,
                      tc_gemm_kernel<1, false, 0, 1, 1, false, true>, tc_gemm_kernel<1, true, 0, 1, 1, false, true>,
                      tc_gemm_kernel<2, false, 0, 1, 1, false, true>, tc_gemm_kernel<2, true, 0, 1, 1, false, true>,
                      tc_gemm_kernel<3, false, 0, 1, 1, false, true>, tc_gemm_kernel<4, false, 0, 1, 1, false, true>};
    for (KernelFn f : all) P3D_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    KernelFn two[] = {tc_gemm_kernel<0, false, 0, 1, 2>, tc_gemm_kernel<0, false, 1, 1, 2>, tc_gemm_kernel<0, true, 0, 1, 2>,
                      tc_gemm_kernel<0, true, 1, 1, 2>, tc_gemm_kernel<1, false, 0, 1, 2>, tc_gemm_kernel<1, true, 0, 1, 2>};
    for (KernelFn f : two) {
      P3D_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
      P3D_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    attr.mark();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = d->grid; cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = d->p.occ == 2 ? SMEM2_BYTES : SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  if (d->p.cn > 1 || d->p.cg == 2) {
    attrs[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    attrs[cfg.numAttrs].val.clusterDim.x = d->p.cg == 2 ? 2 : d->p.cn; attrs[cfg.numAttrs].val.clusterDim.y = 1;
    attrs[cfg.numAttrs].val.clusterDim.z = 1;
    cfg.attrs = attrs; ++cfg.numAttrs;
  }
  if (d->pdl) {
    // programmatic dependent launch: this kernel may start while its predecessor in the stream drains; it orders
    // itself behind the predecessor's memory with griddepcontrol.wait
    attrs[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs; ++cfg.numAttrs;
  }
  P3D_CUDA(cudaLaunchKernelEx(&cfg, fn, d->ta, d->tb, d->tc, d->p));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

int gemm(const GemmArgs& g, cudaStream_t st) {
  GemmPlan pl;
  P3D_TRY(plan(g, &pl));
  return launch(pl, st);
}

}  // namespace tcg
}  // namespace p3d
