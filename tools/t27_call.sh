#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_resident_gpu.py -m gpu -q -x > $out/t27_pytest_resident.log 2>&1; tail -5 $out/t27_pytest_resident.log
LIREC_SWEEP_ONLY=gather timeout 300 python tools/stress_sweep.py 2>>$out/t27_err.log | tee $out/t27_gather_sweep.txt
timeout 600 python tools/library_path_probe.py --batches 64,256 2>>$out/t27_err.log | tee $out/t27_library_path.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no_configs --no_cpu_baseline --no_traffic > $out/t27_bench.json 2>>$out/t27_err.log
python -c "
import json;d=json.loads(open('$out/t27_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'pre',round(d['e2e_precollated']['value']),'streamed',round(d['e2e_streamed']['value']))"
timeout 900 python -m pytest tests -m gpu -q -x > $out/t27_pytest.log 2>&1; tail -3 $out/t27_pytest.log
tail -5 $out/t27_err.log
