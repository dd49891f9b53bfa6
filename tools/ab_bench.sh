#!/bin/bash
# A/B: run the isolated GEMM bench and the step bench for each variant library in variants/
for lib in variants/*.so; do
  echo "=== $lib"
  LIREC_B200_LIB=$PWD/$lib python tools/gemm_probe.py 2>&1 | grep -E "^bench|FAIL"
  LIREC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 40 --warmup 3 --no_cpu_baseline --dump_profile gpurun_out/ab_$(basename $lib .so).txt 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('clips/s %.0f ms/step %.3f gemm_ms %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['clocks']))"
  cut -d' ' -f1,2,4 gpurun_out/ab_$(basename $lib .so).txt | tr '\n' ';'; echo
done
