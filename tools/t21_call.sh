#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_pooling_gpu.py tests/test_loss_gpu.py tests/test_model_gpu.py tests/test_rows_gpu.py -m gpu -q -x > $out/t21_pytest.log 2>&1; tail -3 $out/t21_pytest.log
for co in 1 0; do for B in 1024 64; do
LIREC_DP_CORESIDENT=$co timeout 300 python tools/overlap_probe.py --batch $B 2>>$out/t21_err.log | tee -a $out/t21_overlap_probe.txt
done; done
LIREC_SWEEP_ONLY=roi,softmax timeout 300 python tools/stress_sweep.py 2>&1 | tee $out/t21_stress.txt | grep -v single-pass
tail -5 $out/t21_err.log
