#!/bin/bash
# Build A/B variants of liblirec_b200.so into variants/<name>.so (git-ignored, travels with gpurun), then
# restore the default build.  tools/ab_bench.sh runs the step bench for each of them on ONE box, interleaved:
#   bash tools/build_variants.sh "name1:-DX=1 -DY=2" "name2:..." && gpurun --timeout 900 -- 'bash tools/ab_bench.sh'
set -e
cd "$(dirname "$0")/.."
rm -rf variants
mkdir -p variants
build() {   # name, defines
  LIREC_NVCC_DEFINES="$2" python -m lirec_b200.build --force > /dev/null
  cp lirec_b200/liblirec_b200.so variants/$1.so
  echo "built variants/$1.so  [$2]"
}
build default ""
for spec in "$@"; do
  build "${spec%%:*}" "${spec#*:}"
done
python -m lirec_b200.build --force > /dev/null    # the default build is what ships
