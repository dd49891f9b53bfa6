#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/t26_pytest.log 2>&1; tail -5 $out/t26_pytest.log
LIREC_SWEEP_ONLY=softmax timeout 300 python tools/stress_sweep.py 2>>$out/t26_err.log | tee $out/t26_softmax_sweep.txt
b() { # name, args
  timeout 300 python bench.py --only_value $2 2>>$out/t26_err.log | tail -1 > $out/t26_$1.json
  python -c "import json;d=json.load(open('$out/t26_$1.json'));print('%-28s %.0f clips/s %.4f ms gemm %.4f'%('$1',d['value'],d['ms_per_step'],d['gemm_ms_per_step']))"
}
b base "--steps 100 --warmup 5"
b b64 "--batch 64 --steps 400 --warmup 10"
tail -5 $out/t26_err.log
