// S1 exchange step over NVLink peer memory: the path's only collective is a sum of <= 128 float64 moment partial sums
// (Standardizer / advantage statistics; where the reference reduces: normalize.py:48, ppo.py:336, gail_TRPO.py:128,
// networks.py:76-81).  At 8 GPUs an NCCL all-reduce of 544 bytes costs ~60 us per rollout step, launch, stream hand-off
// and protocol included; here every rank owns a mailbox in its own HBM, mapped into every peer process with CUDA IPC, and
// ONE small kernel per rank (a) stores its partial sums, tagged with the round's sequence number, into every rank's
// mailbox over NVLink, (b) spins on its own mailbox until every rank's words carry the tag, (c) adds the world's
// contributions in rank order -- every rank gets bit-identical sums.  Two parity slots let a fast rank start the
// next round while a slow one still reads the previous.  A spin that exceeds the timeout (30 s by default, 0 = wait for
// ever like NCCL) gives the round up: the whole output vector is NaN on that rank -- never a partial sum -- and an error
// word is set that the host reads (and clears) with om_mailbox_timed_out.
#include <cstdlib>
#include <cstring>

#include "om_common.cuh"

namespace om {

constexpr int MB_MAX_WORLD = 16, MB_MAX_N = 128;

// Wire format (NCCL's "LL" idea): every 32-bit half of a value travels in ONE 8-byte word together with the low 32 bits
// of the round's sequence number, so a reader that sees the flag also sees the data -- no fences and no separate flag
// word.  (A first version used __threadfence_system + release / acquire flags: the system-scope fence stalled behind
// concurrent device->host copies of the rollout and halved the end-to-end throughput.)
struct MailBox {
  unsigned long long word[2][MB_MAX_WORLD][MB_MAX_N][2];   // [parity][source rank][value][lo / hi half]: data | flag << 32
  unsigned long long seq;            // rounds this rank has started (device-side: graph-replay safe)
  unsigned int timed_out;
};

struct MailPeers { MailBox* box[MB_MAX_WORLD]; };

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// in / out may alias (no __restrict__): thread tid reads in[tid] before it writes out[tid].
// timeout_ns == 0 waits for ever (what NCCL does).  A round that times out publishes NO partial sum: every value of the
// round's output is NaN on the rank that gave up, and its timed_out word is set for the host (om_mailbox_timed_out).
__global__ void __launch_bounds__(MB_MAX_N) mailbox_allreduce_kernel(MailPeers peers, MailBox* mine, int world, int rank,
                                                                     const double* in, double* out, int n,
                                                                     unsigned long long timeout_ns) {
  __shared__ unsigned long long s_seq;
  const int tid = threadIdx.x;
  if (tid == 0) s_seq = ++mine->seq;
  __syncthreads();
  const unsigned long long flag = (s_seq & 0xffffffffull) << 32;
  const int parity = (int)(s_seq & 1ull);
  double s = 0.0;
  int gave_up = 0;
  if (tid < n) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(in[tid]);
    const unsigned long long w0 = (bits & 0xffffffffull) | flag, w1 = (bits >> 32) | flag;
    for (int p = 0; p < world; ++p) {                               // NVLink stores (p == rank: local)
      unsigned long long* dst = peers.box[p]->word[parity][rank][tid];
      st_relaxed_sys(dst, w0);
      st_relaxed_sys(dst + 1, w1);
    }
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < world && !gave_up; ++r) {                   // rank order: identical bits on every rank
      const unsigned long long* src = mine->word[parity][r][tid];
      unsigned long long a, b;
      for (;;) {
        a = ld_relaxed_sys(src);
        b = ld_relaxed_sys(src + 1);
        if ((a & 0xffffffff00000000ull) == flag && (b & 0xffffffff00000000ull) == flag) break;
        if (timeout_ns && global_timer_ns() - t0 > timeout_ns) { gave_up = 1; break; }
      }
      s += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    }
  }
  if (__syncthreads_or(gave_up)) {                                  // all or nothing: never a partly summed vector
    if (tid == 0) mine->timed_out = 1u;
    s = __longlong_as_double(0x7ff8000000000000ll);
  }
  if (tid < n) out[tid] = s;
}

}  // namespace om

using namespace om;

struct OmMailbox {
  int world = 1, rank = 0, connected = 0;
  unsigned long long timeout_ns = 30000000000ull;     // 30 s; OM_MAILBOX_TIMEOUT_MS / om_mailbox_set_timeout_ms, 0 = never
  MailBox* mine = nullptr;
  MailPeers peers{};
  bool opened[MB_MAX_WORLD] = {};
};

extern "C" int om_mailbox_create(int world, int rank, OmMailbox** out, unsigned char* handle_out) {
  OM_REQUIRE(out && handle_out, "om_mailbox_create: null argument");
  OM_REQUIRE(world >= 1 && world <= MB_MAX_WORLD && rank >= 0 && rank < world, "om_mailbox_create: need 1 <= world <= %d, 0 <= rank < world", MB_MAX_WORLD);
  static_assert(sizeof(cudaIpcMemHandle_t) == OM_MAILBOX_HANDLE_BYTES, "handle size");
  OmMailbox* mb = new OmMailbox();
  mb->world = world; mb->rank = rank;
  if (const char* f = getenv("OM_MAILBOX_TIMEOUT_MS")) mb->timeout_ns = (unsigned long long)(atof(f) * 1e6);   // read once, here
  cudaError_t e = cudaMalloc((void**)&mb->mine, sizeof(MailBox));
  if (e == cudaSuccess) e = cudaMemset(mb->mine, 0, sizeof(MailBox));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, mb->mine);
  if (e != cudaSuccess) {
    if (mb->mine) cudaFree(mb->mine);
    delete mb;
    return fail("om_mailbox_create: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle_out, &h, sizeof h);
  mb->peers.box[rank] = mb->mine;
  *out = mb;
  return 0;
}

extern "C" int om_mailbox_connect(OmMailbox* mb, const unsigned char* all_handles) {
  OM_REQUIRE(mb && all_handles, "om_mailbox_connect: null argument");
  for (int p = 0; p < mb->world; ++p) {
    if (p == mb->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)p * OM_MAILBOX_HANDLE_BYTES, sizeof h);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail("om_mailbox_connect: cannot map the mailbox of rank %d (%s); use the NCCL all-reduce", p, cudaGetErrorString(e));
    }
    mb->peers.box[p] = (MailBox*)ptr;
    mb->opened[p] = true;
  }
  mb->connected = 1;
  return 0;
}

extern "C" int om_mailbox_allreduce(OmMailbox* mb, const double* in, double* out, int n, void* stream) {
  OM_REQUIRE(mb && mb->connected, "om_mailbox_allreduce: mailbox not connected");
  OM_REQUIRE(n >= 0 && n <= MB_MAX_N, "om_mailbox_allreduce: at most %d values (n=%d)", MB_MAX_N, n);
  if (n == 0) return 0;
  OM_REQUIRE(in && out, "om_mailbox_allreduce: null argument");
  mailbox_allreduce_kernel<<<1, MB_MAX_N, 0, (cudaStream_t)stream>>>(mb->peers, mb->mine, mb->world, mb->rank, in, out, n,
                                                                     mb->timeout_ns);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_mailbox_set_timeout_ms(OmMailbox* mb, double ms) {
  OM_REQUIRE(mb && ms >= 0.0, "om_mailbox_set_timeout_ms: bad argument");
  mb->timeout_ns = (unsigned long long)(ms * 1e6);
  return 0;
}

// Synchronises with the device (a host read of the word).  Reports whether any round since the last call gave up, and
// clears the word so that the next report is about the rounds after this one.
extern "C" int om_mailbox_timed_out(OmMailbox* mb, int* flag) {
  OM_REQUIRE(mb && flag, "om_mailbox_timed_out: null argument");
  unsigned int v = 0;
  OM_CUDA_OK(cudaMemcpy(&v, &mb->mine->timed_out, sizeof v, cudaMemcpyDeviceToHost));
  if (v) OM_CUDA_OK(cudaMemset(&mb->mine->timed_out, 0, sizeof v));
  *flag = (int)v;
  return 0;
}

extern "C" void om_mailbox_destroy(OmMailbox* mb) {
  if (!mb) return;
  for (int p = 0; p < mb->world; ++p)
    if (mb->opened[p]) cudaIpcCloseMemHandle(mb->peers.box[p]);
  if (mb->mine) cudaFree(mb->mine);
  delete mb;
}
