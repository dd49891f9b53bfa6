#!/bin/bash
# Build A/B variants of liblirec_b200.so into variants/<name>.so (git-ignored, travels with gpurun), then
# restore the default build.  tools/ab_bench.sh runs the GEMM probe and the step bench for each of them:
#   bash tools/build_variants.sh && gpurun --timeout 900 -- 'bash tools/ab_bench.sh'
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() {   # name, defines
  LIREC_NVCC_DEFINES="$2" python -m lirec_b200.build --force > /dev/null
  cp lirec_b200/liblirec_b200.so variants/$1.so
  echo "built variants/$1.so  [$2]"
}
build default ""
build ebt_unroll2 "-DLIREC_EBT_UNROLL=2"
build ebt_unroll4 "-DLIREC_EBT_UNROLL=4"
build ebt_zsplit2 "-DLIREC_EBT_ZSPLIT=2"
build ebt_minblocks3 "-DLIREC_EBT_MIN_BLOCKS=3"
build epi_warps4 "-DLIREC_EPI_WARPS=4"
python -m lirec_b200.build --force > /dev/null    # the default build is what ships
