// Small all-reduce over NVLink peer memory for the data-parallel training step.
//
// The step needs eleven tiny sum-reductions in strict sequence (SyncBN sums forward and backward: 2 x linear_size
// doubles per BN layer, and the loss): as NCCL calls they cost 20-40 us EACH (latency, not bandwidth) and made an
// 8-GPU step slower than a 1-GPU one.  Here every rank owns a small exchange buffer that all ranks of the node map
// through CUDA IPC; one kernel per reduction
//   1. stores the local vector into slot [seq % NSLOTS][my rank] of EVERY rank's buffer (NVLink peer stores),
//   2. fences system-wide and raises flag [slot][my rank] = seq on every rank,
//   3. spins until all ranks' flags of this slot carry seq,
//   4. adds the `world` vectors in rank order (bit-identical result on every rank - the optimizer stays replicated).
// One launch, one NVLink round trip.  The sequence number lives in device memory and is advanced by the kernel, so
// the launches replay inside the step's CUDA graph.  Slot reuse is safe with >= 2 slots: nobody can be more than one
// reduction ahead of the slowest rank (it needs that rank's flag), and a rank issues reduction seq+1 only after its
// own kernel for seq has finished reading.  The big gradient all-reduce stays with NCCL (bandwidth bound).
//
// The fused training epilogues of tc_gemm.cu run the same protocol on the same buffers from INSIDE the GEMM that needs
// the sums (FusedTrain in common.cuh); this kernel remains for the unfused path (batches whose tiles do not all fit the
// SMs) and for vectors that belong to no GEMM.  Exchanges of one model must not run concurrently (one sequence counter).
// Like an NCCL collective the kernels wait for their peers without a timeout: ranks must stay in lockstep (same number
// of steps in the same order).  P3D_SYNC_TIMEOUT_S=<seconds> turns a longer wait into a message + trap (diagnostics).
#include <cstring>

#include "common.cuh"

namespace p3d {
namespace p2p {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Optional tail of a forward SyncBN reduction: the BatchNorm finalisation (mean / biased variance over the global
// batch, moving averages with momentum .99 - train.cu's bn_finalize_kernel) runs on the freshly summed columns.
__global__ void __launch_bounds__(512) allreduce_kernel(const Peers peers, int rank, int world, double* __restrict__ buf, int n,
                                                        const BnFinalize fin, long long wait_limit_ns) {
  Layout* me = peers.p[rank];
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = me->seq + 1;           // sequence numbers start at 1 (flags are zero-initialised)
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int slot = static_cast<int>(seq % NSLOTS);
  // 1. my vector -> everybody (16-byte peer stores)
  const int n2 = n >> 1;
  for (int r = 0; r < world; ++r) {
    double* dst = peers.p[r]->data[slot][rank];
    for (int i = threadIdx.x; i < n2; i += blockDim.x) reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(buf)[i];
    if ((n & 1) && threadIdx.x == 0) dst[n - 1] = buf[n - 1];
  }
  // The block barrier orders every thread's peer stores before the flag threads; their st.release.sys is cumulative,
  // so one system-scope fence per flag thread publishes the whole vector (instead of 512 membar.sys).
  __syncthreads();
  // 2. raise my flag on every rank   3. wait for everybody's flag here
  if (threadIdx.x < world) {
    st_release_sys(&peers.p[threadIdx.x]->flag[slot][rank], seq);
    unsigned long long t0 = 0;
    if (wait_limit_ns > 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(&me->flag[slot][threadIdx.x]) < seq) {
      if (wait_limit_ns > 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > static_cast<unsigned long long>(wait_limit_ns)) {
          printf("p3d: peer all-reduce timed out (rank %d waits for %d, seq %llu)\n", rank, (int)threadIdx.x, seq); __trap();
        }
      }
    }
  }
  __syncthreads();
  // 4. sum in rank order
  if (fin.mean == nullptr) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < world; ++r) s += *reinterpret_cast<volatile double*>(&me->data[slot][r][i]);
      buf[i] = s;
    }
  } else {
    const int L = n >> 1;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
      double s0 = 0.0, s1 = 0.0;
      for (int r = 0; r < world; ++r) {
        s0 += *reinterpret_cast<volatile double*>(&me->data[slot][r][c]);
        s1 += *reinterpret_cast<volatile double*>(&me->data[slot][r][L + c]);
      }
      buf[c] = s0; buf[L + c] = s1;
      const double mu = s0 * fin.invB;
      double var = s1 * fin.invB - mu * mu;
      if (var < 0) var = 0;
      fin.mean[c] = static_cast<float>(mu);
      fin.rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
      fin.mm[c] = fin.mm[c] * kBnMomentum + static_cast<float>(mu) * (1.f - kBnMomentum);
      fin.mv[c] = fin.mv[c] * kBnMomentum + static_cast<float>(var) * (1.f - kBnMomentum);
    }
  }
  if (threadIdx.x == 0) me->seq = seq;
}

// Flat gradient all-reduce over NVLink peer memory (replaces ncclAllReduce of 17.2 MB, measured 99 us on 8 B200): every
// rank's gradient buffer is mapped by all ranks; rank r owns the r-th 1/world of it.  After an entry handshake (everybody's
// gradient is complete) each rank PULLS its shard from all ranks (16-byte system-scope loads, which do not hit the
// non-coherent L1), adds the `world` values in rank order and PUSHES the sum back into every rank's buffer - reduce-scatter
// and all-gather in one pass, every element reduced exactly once, so all ranks end up with bit-identical gradients.  An exit
// handshake (everybody has finished writing into my buffer) ends the kernel.  No CTA waits for another CTA of its own
// grid except through plain atomics, so the grid needs no co-residency.
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned long long* f, unsigned long long seq, long long wait_limit_ns, int rank, int peer) {
  unsigned long long t0 = 0;
  if (wait_limit_ns > 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(f) < seq) {
    if (wait_limit_ns > 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > static_cast<unsigned long long>(wait_limit_ns)) {
        printf("p3d: gradient all-reduce timed out (rank %d waits for %d, seq %llu)\n", rank, peer, seq); __trap();
      }
    }
  }
}
template <int W>      // world size as a template parameter: the per-rank values stay in registers
__global__ void __launch_bounds__(256) grad_allreduce_kernel(const Peers peers, const GradPeers gp, int rank, long long n4,
                                                             long long n, long long wait_limit_ns, const float* __restrict__ y_local,
                                                             long long y_count, long long y_off) {
  constexpr int world = W;
  Layout* me = peers.p[rank];
  __shared__ unsigned long long s_seq;
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_seq = me->gseq + 1;          // advanced by the last CTA at the very end, when every CTA has read it
  __syncthreads();
  const unsigned long long seq = s_seq;
  // entry: my gradient is complete (stream order); wait until everybody's is
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(&peers.p[threadIdx.x]->gflag[0][rank], seq);
  if (threadIdx.x < world) spin_until(&me->gflag[0][threadIdx.x], seq, wait_limit_ns, rank, threadIdx.x);
  __syncthreads();
  const long long per = (n4 + world - 1) / world;
  const long long beg = static_cast<long long>(rank) * per, end = (beg + per < n4) ? beg + per : n4;
  // U independent float4 positions per thread and trip: all their peer loads are in flight together (an NVLink round trip is
  // ~2-3 us; position by position the two-rank case spent 7 of them in sequence: 45 us for 2 x 8.6 MB)
  constexpr int U = (W <= 2) ? 4 : ((W <= 4) ? 2 : 1);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = beg + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < end; i0 += U * stride) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < end) {
#pragma unroll
        for (int r = 0; r < W; ++r) v[u][r] = ld_relaxed_sys_f4(gp.g[r] + 4 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < end) {
        float4 a = v[u][0];
#pragma unroll
        for (int r = 1; r < W; ++r) { a.x += v[u][r].x; a.y += v[u][r].y; a.z += v[u][r].z; a.w += v[u][r].w; }
#pragma unroll
        for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(gp.g[r] + 4 * i) = a;
      }
    }
  }
  // this rank's rows of the step outputs -> every rank's ybuf, spread over the whole grid as 8-byte stores (rows x 48 or 42
  // floats: always an even count at an even offset); the exit handshake below publishes them
  if (y_local != nullptr) {
    const long long c2 = y_count >> 1;
    const long long total = c2 * world;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int r = static_cast<int>(idx / c2);
      const long long i = idx - r * c2;
      reinterpret_cast<float2*>(peers.p[r]->ybuf + y_off)[i] = reinterpret_cast<const float2*>(y_local)[i];
    }
  }
  if (rank == 0 && blockIdx.x == 0) {                   // a length that is not a multiple of 4: the tail, element by element
    for (long long i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
      float a = 0.f;
      for (int r = 0; r < world; ++r) a += *reinterpret_cast<volatile float*>(gp.g[r] + i);
      for (int r = 0; r < world; ++r) gp.g[r][i] = a;
    }
  }
  // exit: the block barrier hands this CTA's stores to thread 0, whose system-scope fence publishes them before the counter
  // moves; the last CTA of this rank tells everybody and waits until everybody has finished writing into this rank's buffer
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(&me->gdone, 1u) == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < world) {
      st_release_sys(&peers.p[threadIdx.x]->gflag[1][rank], seq);
      spin_until(&me->gflag[1][threadIdx.x], seq, wait_limit_ns, rank, threadIdx.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) { me->gdone = 0u; me->gseq = seq; __threadfence(); }
  }
}

struct State {
  Layout* local = nullptr;
  Peers peers{};
  GradPeers gpeers{};
  bool grad_mapped = false;
  Peers* dev_peers = nullptr;       // device copy, for kernels that take the table by pointer (tc_gemm's fused epilogues)
  int world = 0, rank = 0;
  bool ready = false;
};

static State* state_of(p3d_model* m) { return static_cast<State*>(m->p2p_state); }

int local_handle(p3d_model* m, uint8_t* handle128) {
  if (!m->p2p_state) m->p2p_state = new State();
  State* s = state_of(m);
  if (!s->local) {
    P3D_CUDA(cudaMalloc(&s->local, sizeof(Layout)));
    P3D_CUDA(cudaMemset(s->local, 0, sizeof(Layout)));
    P3D_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  P3D_CUDA(cudaIpcGetMemHandle(&h, s->local));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle128, &h, 64);
  P3D_CUDA(cudaIpcGetMemHandle(&h, m->grad));           // the flat gradient buffer, for the peer-memory all-reduce
  memcpy(handle128 + 64, &h, 64);
  return P3D_OK;
}

int detach(p3d_model* m);

int attach(p3d_model* m, const uint8_t* handles, int rank, int world) {
  State* s = state_of(m);
  P3D_REQUIRE(s && s->local, "p2p attach: call p3d_model_p2p_handle first");
  P3D_REQUIRE(world >= 2 && world <= MAXW && rank >= 0 && rank < world, "p2p attach: bad rank/world");
  for (int r = 0; r < MAXW; ++r) { s->peers.p[r] = nullptr; s->gpeers.g[r] = nullptr; }
  s->grad_mapped = false;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { s->peers.p[r] = s->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 128 * r, 64);
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("peer memory of rank %d is not reachable (%s): the step keeps NCCL for the small reductions", r, cudaGetErrorString(e));
      detach(m);                        // close what was opened so far
      return P3D_ERR_CUDA;
    }
    s->peers.p[r] = static_cast<Layout*>(ptr);
  }
  // the gradient buffers: optional (without them the gradient keeps going through NCCL)
  bool gok = true;
  for (int r = 0; r < world && gok; ++r) {
    if (r == rank) { s->gpeers.g[r] = m->grad; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 128 * r + 64, 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); gok = false; break; }
    s->gpeers.g[r] = static_cast<float*>(ptr);
  }
  if (!gok) {
    for (int r = 0; r < world; ++r) { if (s->gpeers.g[r] && s->gpeers.g[r] != m->grad) cudaIpcCloseMemHandle(s->gpeers.g[r]); s->gpeers.g[r] = nullptr; }
  }
  s->grad_mapped = gok;
  s->world = world; s->rank = rank;
  if (!s->dev_peers) P3D_CUDA(cudaMalloc(&s->dev_peers, sizeof(Peers)));
  P3D_CUDA(cudaMemcpy(s->dev_peers, &s->peers, sizeof(Peers), cudaMemcpyHostToDevice));
  s->ready = true;
  return P3D_OK;
}

// Undo a (possibly partial) attach: close what was opened, keep the local buffer.  The step then uses NCCL.
int detach(p3d_model* m) {
  State* s = state_of(m);
  if (!s) return P3D_OK;
  for (int r = 0; r < MAXW; ++r) {
    if (s->peers.p[r] && s->peers.p[r] != s->local) cudaIpcCloseMemHandle(s->peers.p[r]);
    s->peers.p[r] = nullptr;
    if (s->gpeers.g[r] && s->gpeers.g[r] != m->grad) cudaIpcCloseMemHandle(s->gpeers.g[r]);
    s->gpeers.g[r] = nullptr;
  }
  s->world = 0; s->ready = false; s->grad_mapped = false;
  return P3D_OK;
}

bool ready(const p3d_model* m) { return m->p2p_state && static_cast<const State*>(m->p2p_state)->ready; }
bool grad_ready(const p3d_model* m) {
  static const bool on = [] { const char* e = getenv("P3D_P2P_GRAD"); return !(e && e[0] == '0'); }();
  if (!on || !ready(m)) return false;
  const State* s = static_cast<const State*>(m->p2p_state);
  return s->grad_mapped && ((s->world >= 2 && s->world <= 8) || s->world == 16);
}
const float* gathered_outputs(const p3d_model* m) { return ready(m) ? static_cast<const State*>(m->p2p_state)->local->ybuf : nullptr; }
int allreduce_grad(p3d_model* m, size_t n, const float* y_local, long long rows, long long row0, int out, cudaStream_t st) {
  State* s = state_of(m);
  P3D_REQUIRE(s && s->ready && s->grad_mapped, "peer gradient all-reduce: gradient buffers not mapped");
  static const long long wait_ns = [] { const char* e = getenv("P3D_SYNC_TIMEOUT_S"); return e ? static_cast<long long>(atof(e) * 1e9) : 0LL; }();
  const long long n4 = static_cast<long long>(n / 4), nn = static_cast<long long>(n);
  const dim3 grid(2 * m->num_sms), block(256);
  const long long y_count = y_local ? rows * out : 0, y_off = row0 * out;
  switch (s->world) {
#define P3D_GAR(W) case W: grad_allreduce_kernel<W><<<grid, block, 0, st>>>(s->peers, s->gpeers, s->rank, n4, nn, wait_ns, y_local, y_count, y_off); break;
    P3D_GAR(2) P3D_GAR(3) P3D_GAR(4) P3D_GAR(5) P3D_GAR(6) P3D_GAR(7) P3D_GAR(8) P3D_GAR(16)
#undef P3D_GAR
    default: set_error("peer gradient all-reduce: world size %d has no instantiation", s->world); return P3D_ERR_ARG;
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
const Peers* device_peers(const p3d_model* m) { return ready(m) ? static_cast<const State*>(m->p2p_state)->dev_peers : nullptr; }

int allreduce_small(p3d_model* m, double* buf, size_t n, cudaStream_t st, const BnFinalize* fin) {
  State* s = state_of(m);
  P3D_REQUIRE(s && s->ready && n >= 1 && n <= MAXN, "peer all-reduce: not attached or vector too long");
  P3D_REQUIRE(!fin || (n % 2) == 0, "peer all-reduce: the BatchNorm tail needs [sum | sumsq]");
  static const long long wait_ns = [] { const char* e = getenv("P3D_SYNC_TIMEOUT_S"); return e ? static_cast<long long>(atof(e) * 1e9) : 0LL; }();
  allreduce_kernel<<<1, 512, 0, st>>>(s->peers, s->rank, s->world, buf, static_cast<int>(n), fin ? *fin : BnFinalize(), wait_ns);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

void destroy(p3d_model* m) {
  State* s = state_of(m);
  if (!s) return;
  for (int r = 0; r < MAXW; ++r) {
    if (s->peers.p[r] && s->peers.p[r] != s->local) cudaIpcCloseMemHandle(s->peers.p[r]);
    if (s->gpeers.g[r] && s->gpeers.g[r] != m->grad) cudaIpcCloseMemHandle(s->gpeers.g[r]);
  }
  cudaFree(s->dev_peers);
  cudaFree(s->local);
  delete s;
  m->p2p_state = nullptr;
}

}  // namespace p2p
}  // namespace p3d

using namespace p3d;

extern "C" {

int p3d_model_p2p_handle(p3d_model* m, uint8_t* handle128_host) {
  P3D_REQUIRE(m && handle128_host, "p2p_handle: null argument");
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  return p2p::local_handle(m, handle128_host);
}

int p3d_model_p2p_attach(p3d_model* m, const uint8_t* handles_host, int rank, int world) {
  P3D_REQUIRE(m && handles_host, "p2p_attach: null argument");
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  return p2p::attach(m, handles_host, rank, world);
}

int p3d_model_p2p_detach(p3d_model* m) {
  P3D_REQUIRE(m, "p2p_detach: null argument");
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  return p2p::detach(m);
}

}  // extern "C"
