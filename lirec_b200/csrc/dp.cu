// dp.cu — data-parallel gradient exchange over NVSwitch multicast, bucket by bucket.
//
// The reference is single-process; data-parallel training over clips adds exactly one
// exchange per step: the sum of the flat fp32 gradient buffer over ranks, followed by Adam
// (reference optimizer: torch.optim.Adam, mlp/model.py:599-601).  The stock way is ncclAllReduce and
// then the Adam kernel, both after the last backward launch.  Here the flat gradient buffer is cut into
// buckets (lirec_b200/dp.py: gate + heads, 53 % of the bytes and final ~0.6 ms before backward ends, and the
// encoders) and each bucket runs, on whatever stream the caller chooses, the chain
//
//   barrier   cross-GPU: every rank's backward has written this bucket (flags in peer memory,
//             release/acquire at system scope; one CTA);
//   reduce    each rank reduces ITS 1/world shard of the bucket inside the switch
//             (multimem.ld_reduce.add.v4.f32 on the multicast address pulls the shard from all ranks and
//             sums it in the NVSwitch) and broadcasts the sum to every rank with one multimem.st — the
//             bucket is reduced in place, 1/world of it leaves each GPU once and arrives once, both link
//             directions busy at the same time;
//   barrier   all shards have landed everywhere and nobody reads peer memory any more (so the next
//             backward may overwrite the gradients);
//
// followed by lirec_adam_flat over the bucket.  The first bucket's chain is launched on a side stream as
// soon as backward has recorded its "head gradients final" event (lirec_model_backward_ex), so its exchange
// and its Adam pass overlap the second-layer / first-layer backward stages.
//
// Round 1 did all of this in ONE kernel whose CTAs spun on each other (grid-wide flags): that needs every
// CTA co-resident, which a plain launch does not guarantee.  The chain above is separate
// launches ordered by the stream: the only spin left is the single-CTA cross-GPU barrier, which waits for
// PEERS and never for another CTA of its own grid.
//
// The gradient buffer must be symmetric memory mapped into a multicast object on every rank
// (lirec_b200/dp.py allocates it with torch.distributed._symmetric_memory; torch only provides the
// allocation and the rendezvous, no arithmetic).  Every spin is bounded: a dead peer traps (kernel error)
// instead of hanging the GPU forever.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace lirec {
namespace dp {

constexpr int THREADS = 256;
constexpr int MAX_CHANNELS = 4;                 // independent chains in flight (one per bucket)
constexpr uint32_t SPIN_LIMIT = 1u << 28;       // minutes of NVLink round trips

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag[slot * world + src] in the flag buffer of rank dst: "src has reached `slot`".  A put flips the
// peer's slot 0 -> 1 (spinning while the previous use is still unconsumed), a wait flips the local slot
// 1 -> 0, so the slots reset themselves and the barrier is reusable without an epoch.
__global__ void __launch_bounds__(32) barrier_kernel(uint32_t* const* flags, int rank, int world, int slot) {
  const int t = threadIdx.x;
  __threadfence_system();        // everything earlier kernels of this stream wrote is visible before the signal
  if (t < world) {
    uint32_t spins = 0;
    uint32_t* remote = flags[t] + slot * world + rank;
    while (cas_release_sys(remote, 0u, 1u) != 0u) {
      if (++spins > SPIN_LIMIT) __trap();
      __nanosleep(64);
    }
    uint32_t* local = flags[rank] + slot * world + t;
    spins = 0;
    while (cas_acquire_sys(local, 1u, 0u) != 1u) {
      if (++spins > SPIN_LIMIT) __trap();
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

// This rank's shard [beg4, end4) (in float4 units) of a bucket: reduced in the switch, broadcast to all.
// Four independent 16-byte reductions in flight per thread: an in-switch reduction is a ~3 us round trip, so
// the links are only covered by many outstanding requests.
__global__ void __launch_bounds__(THREADS) reduce_bcast_kernel(float* __restrict__ g_mc, int64_t beg4, int64_t end4) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = beg4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < end4; i += 4 * stride) {
    const float4 a = multimem_ld_reduce_add(g_mc + 4 * i);
    const float4 b = multimem_ld_reduce_add(g_mc + 4 * (i + stride));
    const float4 c = multimem_ld_reduce_add(g_mc + 4 * (i + 2 * stride));
    const float4 d = multimem_ld_reduce_add(g_mc + 4 * (i + 3 * stride));
    multimem_st(g_mc + 4 * i, a);
    multimem_st(g_mc + 4 * (i + stride), b);
    multimem_st(g_mc + 4 * (i + 2 * stride), c);
    multimem_st(g_mc + 4 * (i + 3 * stride), d);
  }
  for (; i < end4; i += stride) multimem_st(g_mc + 4 * i, multimem_ld_reduce_add(g_mc + 4 * i));
  __threadfence_system();        // this thread's broadcasts are performed before the kernel retires
}

}  // namespace dp
}  // namespace lirec

using namespace lirec;

extern "C" int lirec_dp_flag_words(int32_t world) { return 2 * dp::MAX_CHANNELS * std::max(world, 1); }

extern "C" int lirec_dp_exchange(void* grad_multicast, int64_t offset, int64_t n, int32_t rank, int32_t world,
                                 const void* flag_ptrs_dev, int32_t channel, void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(grad_multicast && flag_ptrs_dev, "dp_exchange: null argument");
  LIREC_REQUIRE(n > 0 && n % 4 == 0 && offset >= 0 && offset % 4 == 0,
                "dp_exchange: offset=%lld n=%lld must be multiples of 4 floats", (long long)offset, (long long)n);
  LIREC_REQUIRE(world >= 2 && world <= 32 && rank >= 0 && rank < world, "dp_exchange: rank %d of %d", rank, world);
  LIREC_REQUIRE(channel >= 0 && channel < dp::MAX_CHANNELS, "dp_exchange: channel %d", channel);
  LIREC_REQUIRE((reinterpret_cast<uintptr_t>(grad_multicast) & 15) == 0, "dp_exchange: buffer must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* const* flags = static_cast<uint32_t* const*>(flag_ptrs_dev);
  float* mc = static_cast<float*>(grad_multicast) + offset;
  const int64_t n4 = n / 4;
  const int64_t beg = n4 * rank / world, end = n4 * (rank + 1) / world;
  dp::barrier_kernel<<<1, 32, 0, s>>>(flags, rank, world, 2 * channel);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  if (end > beg) {
    const int64_t per_cta = static_cast<int64_t>(dp::THREADS) * 4;
    const int grid = static_cast<int>(std::min<int64_t>((end - beg + per_cta - 1) / per_cta, 148 * 4));
    dp::reduce_bcast_kernel<<<grid, dp::THREADS, 0, s>>>(mc, beg, end);
    LIREC_CUDA_OK(cudaGetLastError());
    note_launch();
  }
  dp::barrier_kernel<<<1, 32, 0, s>>>(flags, rank, world, 2 * channel + 1);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}
