#!/bin/bash
# A/B on one box: the device-resident step bench for each variant library in variants/, ROUNDS times interleaved
# (box-to-box and minute-to-minute drift is ~1-2 %, more than most kernel changes are worth).
ROUNDS=${ROUNDS:-3}
ARGS=${ARGS:---steps 60 --warmup 5}
for r in $(seq $ROUNDS); do
  for lib in variants/*.so; do
    LIREC_B200_LIB=$PWD/$lib timeout 300 python bench.py --only_value $ARGS 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-22s round $r  clips/s %.0f  ms/step %.4f  gemm_ms %.4f  %s' % ('$(basename $lib .so)', d['value'], d['ms_per_step'], d['gemm_ms_per_step'], d['gemm_launch_ms']))"
  done
done
