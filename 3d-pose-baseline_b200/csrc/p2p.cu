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

struct State {
  Layout* local = nullptr;
  Peers peers{};
  Peers* dev_peers = nullptr;       // device copy, for kernels that take the table by pointer (tc_gemm's fused epilogues)
  int world = 0, rank = 0;
  bool ready = false;
};

static State* state_of(p3d_model* m) { return static_cast<State*>(m->p2p_state); }

int local_handle(p3d_model* m, uint8_t* handle64) {
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
  memcpy(handle64, &h, 64);
  return P3D_OK;
}

int detach(p3d_model* m);

int attach(p3d_model* m, const uint8_t* handles, int rank, int world) {
  State* s = state_of(m);
  P3D_REQUIRE(s && s->local, "p2p attach: call p3d_model_p2p_handle first");
  P3D_REQUIRE(world >= 2 && world <= MAXW && rank >= 0 && rank < world, "p2p attach: bad rank/world");
  for (int r = 0; r < MAXW; ++r) s->peers.p[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { s->peers.p[r] = s->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
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
  }
  s->world = 0; s->ready = false;
  return P3D_OK;
}

bool ready(const p3d_model* m) { return m->p2p_state && static_cast<const State*>(m->p2p_state)->ready; }
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
  for (int r = 0; r < MAXW; ++r)
    if (s->peers.p[r] && s->peers.p[r] != s->local) cudaIpcCloseMemHandle(s->peers.p[r]);
  cudaFree(s->dev_peers);
  cudaFree(s->local);
  delete s;
  m->p2p_state = nullptr;
}

}  // namespace p2p
}  // namespace p3d

using namespace p3d;

extern "C" {

int p3d_model_p2p_handle(p3d_model* m, uint8_t* handle64_host) {
  P3D_REQUIRE(m && handle64_host, "p2p_handle: null argument");
  P3D_CUDA(cudaSetDevice(m->cfg.device));
  return p2p::local_handle(m, handle64_host);
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
