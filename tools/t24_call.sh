#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_pooling_gpu.py -m gpu -q -x > $out/t24_pytest_pool.log 2>&1; tail -5 $out/t24_pytest_pool.log
for L in 0 1 4; do
  echo "## LIREC_SP_LANES=$L" | tee -a $out/t24_softmax_sweep.txt
  LIREC_SP_LANES=$L LIREC_SWEEP_ONLY=softmax timeout 300 python tools/stress_sweep.py 2>>$out/t24_err.log | tee -a $out/t24_softmax_sweep.txt
done
LIREC_SWEEP_ONLY=softmax timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 20 -c 2 \
  -o $out/t24_softpool_fwd -f python tools/stress_sweep.py > $out/t24_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $out/t24_pytest.log 2>&1; tail -3 $out/t24_pytest.log
tail -5 $out/t24_err.log
