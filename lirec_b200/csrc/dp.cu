// dp.cu — data-parallel gradient exchange fused with the optimizer step, over NVSwitch multicast.
//
// The reference is single-process; data-parallel training over clips adds exactly one
// exchange per step: the sum of the flat fp32 gradient buffer over ranks, followed by Adam
// (reference optimizer: torch.optim.Adam, mlp/model.py:599-601).  The stock way is ncclAllReduce and
// then the Adam kernel: 0.23 ms + 0.10 ms per step at 8 GPUs.  Here ONE kernel per rank does both:
//
//   phase 0  cross-GPU barrier: every rank's backward has written its gradients (flags in peer memory,
//            release/acquire at system scope);
//   phase 1  each rank reduces ITS 1/world shard of the buffer inside the switch
//            (multimem.ld_reduce.add.v4.f32 on the multicast address pulls the shard from all ranks and
//            sums it in the NVSwitch) and broadcasts the sum to every rank with one multimem.st — the
//            gradient buffer is reduced in place, 1/world of the buffer leaves each GPU once and
//            arrives once, both link directions busy at the same time;
//   barrier  grid-wide, then cross-GPU: all shards have landed everywhere, nobody reads peer memory
//            any more (so the next backward may overwrite the gradients);
//   phase 2  Adam (coupled L2, bias-corrected) over the rank's full replica + bf16 weight shadow.
//
// The gradient buffer must be symmetric memory mapped into a multicast object on every rank
// (lirec_b200/dp.py allocates it with torch.distributed._symmetric_memory; torch only provides the
// allocation and the rendezvous, no arithmetic).  Every spin is bounded: a protocol bug traps instead
// of hanging the GPU.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace lirec {
namespace dp {

constexpr int THREADS = 512;
constexpr uint32_t SPIN_LIMIT = 1u << 28;   // ~ a few seconds

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
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.acquire.gpu.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.global.release.gpu.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.global.release.gpu.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

// flag[phase * world + src] in the flag buffer of rank dst: "src has reached `phase`".  A put flips the
// peer's slot 0 -> 1 (spinning while the previous use is still unconsumed), a wait flips the local slot
// 1 -> 0, so the slots reset themselves and the barrier is reusable without an epoch.
__device__ __forceinline__ void cross_gpu_barrier(uint32_t* const* flags, int rank, int world, int phase) {
  const int t = threadIdx.x;
  if (t < world) {
    uint32_t spins = 0;
    uint32_t* remote = flags[t] + phase * world + rank;
    while (cas_release_sys(remote, 0u, 1u) != 0u)
      if (++spins > SPIN_LIMIT) __trap();
    uint32_t* local = flags[rank] + phase * world + t;
    spins = 0;
    while (cas_acquire_sys(local, 1u, 0u) != 1u)
      if (++spins > SPIN_LIMIT) __trap();
  }
}

struct Args {
  float* p;
  float* g;          // this rank's gradient buffer (symmetric memory), reduced in place
  float* g_mc;       // multicast address of the same buffer
  float* m;
  float* v;
  __nv_bfloat16* pb;
  int64_t n;
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale;
  int rank, world;
  uint32_t* const* flags;   // device array [world]: every rank's flag buffer (peer-mapped)
  uint32_t* ws;             // local: [0] arrival counter, [1] phase-0 done epoch, [2] barrier done epoch
  uint32_t epoch;           // 1, 2, 3, ... per call (the grid size must not change between calls)
};

__global__ void __launch_bounds__(THREADS) allreduce_adam_kernel(const Args a) {
  const bool leader = blockIdx.x == 0;
  // ---- phase 0: all ranks' gradients are complete --------------------------------------------------
  if (leader) {
    cross_gpu_barrier(a.flags, a.rank, a.world, 0);
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.ws + 1, a.epoch);
  } else {
    if (threadIdx.x == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(a.ws + 1) != a.epoch)
        if (++spins > SPIN_LIMIT) __trap();
    }
    __syncthreads();
  }
  // ---- phase 1: my shard, reduced in the switch and broadcast -------------------------------------------
  const int64_t n4 = a.n / 4;
  const int64_t beg = n4 * a.rank / a.world, end = n4 * (a.rank + 1) / a.world;
  const int64_t gstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t gtid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (int64_t i = beg + gtid; i < end; i += gstride) {
    const float4 s = multimem_ld_reduce_add(a.g_mc + 4 * i);
    multimem_st(a.g_mc + 4 * i, s);
  }
  __threadfence_system();
  __syncthreads();
  // ---- grid barrier + cross-GPU barrier ----------------------------------------------------------------
  if (threadIdx.x == 0) red_release_gpu_add(a.ws, 1u);
  if (leader) {
    if (threadIdx.x == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(a.ws) != gridDim.x * a.epoch)
        if (++spins > SPIN_LIMIT) __trap();
    }
    __syncthreads();
    cross_gpu_barrier(a.flags, a.rank, a.world, 1);
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.ws + 2, a.epoch);
  } else {
    if (threadIdx.x == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(a.ws + 2) != a.epoch)
        if (++spins > SPIN_LIMIT) __trap();
    }
    __syncthreads();
  }
  // ---- phase 2: Adam over the full replica (same arithmetic as loss.cu:adam_kernel) -----------------------
  const float step = a.lr / a.bc1;
  for (int64_t i = gtid; i < n4; i += gstride) {
    float4 pv = reinterpret_cast<const float4*>(a.p)[i];
    const float4 gr = __ldcg(reinterpret_cast<const float4*>(a.g) + i);   // written by peers: read through L2
    float4 mv = reinterpret_cast<const float4*>(a.m)[i];
    float4 vv = reinterpret_cast<const float4*>(a.v)[i];
    float* pp = &pv.x; const float* gg = &gr.x; float* mm = &mv.x; float* vvp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gv = gg[k] * a.grad_scale + a.wd * pp[k];
      mm[k] = a.beta1 * mm[k] + (1.f - a.beta1) * gv;
      vvp[k] = a.beta2 * vvp[k] + (1.f - a.beta2) * gv * gv;
      pp[k] -= step * mm[k] / (sqrtf(vvp[k]) / a.bc2_sqrt + a.eps);
    }
    reinterpret_cast<float4*>(a.m)[i] = mv;
    reinterpret_cast<float4*>(a.v)[i] = vv;
    reinterpret_cast<float4*>(a.p)[i] = pv;
    if (a.pb) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&lo);
      w.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(a.pb)[i] = w;
    }
  }
}

}  // namespace dp
}  // namespace lirec

using namespace lirec;

extern "C" int lirec_dp_grid_size(void) {
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp::allreduce_adam_kernel, dp::THREADS, 0);
    grid = sms * std::max(1, std::min(per_sm, 2));    // all CTAs must be co-resident (they spin on each other)
  }
  return grid;
}

extern "C" int lirec_dp_allreduce_adam(float* param, float* grad, void* grad_multicast, float* exp_avg,
                                       float* exp_avg_sq, void* param_bf16, int64_t n, float lr, float beta1,
                                       float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                                       int32_t rank, int32_t world, const void* flag_ptrs_dev, void* sync_ws,
                                       uint32_t epoch, void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(param && grad && grad_multicast && exp_avg && exp_avg_sq && flag_ptrs_dev && sync_ws,
                "dp_allreduce_adam: null argument");
  LIREC_REQUIRE(n > 0 && n % 4 == 0, "dp_allreduce_adam: n=%lld must be a positive multiple of 4", (long long)n);
  LIREC_REQUIRE(world >= 2 && world <= 64 && rank >= 0 && rank < world, "dp_allreduce_adam: rank %d of %d", rank, world);
  LIREC_REQUIRE(step >= 1 && epoch >= 1, "dp_allreduce_adam: step=%d epoch=%u", step, epoch);
  LIREC_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                  reinterpret_cast<uintptr_t>(grad_multicast) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0,
                "dp_allreduce_adam: buffers must be 16-byte aligned");
  const int grid = lirec_dp_grid_size();
  LIREC_REQUIRE(grid > 0, "dp_allreduce_adam: no device");
  dp::Args a;
  a.p = param; a.g = grad; a.g_mc = static_cast<float*>(grad_multicast); a.m = exp_avg; a.v = exp_avg_sq;
  a.pb = reinterpret_cast<__nv_bfloat16*>(param_bf16);
  a.n = n; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  a.bc2_sqrt = sqrtf(static_cast<float>(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  a.grad_scale = grad_scale;
  a.rank = rank; a.world = world;
  a.flags = static_cast<uint32_t* const*>(flag_ptrs_dev);
  a.ws = static_cast<uint32_t*>(sync_ws);
  a.epoch = epoch;
  dp::allreduce_adam_kernel<<<grid, dp::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}
