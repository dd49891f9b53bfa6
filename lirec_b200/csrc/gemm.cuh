// gemm.cuh — host entry of the grouped tcgen05 GEMM (gemm_tcgen05.cu).
#pragma once
#include "common.cuh"

namespace lirec {
namespace gemm {
// Encodes the tensor maps, builds the tile table and launches ONE persistent kernel.
int run_grouped(const lirec_gemm_problem* probs, int nprobs, cudaStream_t stream);
}  // namespace gemm
}  // namespace lirec
