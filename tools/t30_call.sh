#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/t30_pytest.log 2>&1; tail -3 $out/t30_pytest.log
for r in 1 2 3; do
  LIREC_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no_configs --no_cpu_baseline --no_traffic > $out/t30_bench_$r.json 2>$out/t30_bench_$r.err
  python -c "
import json;d=json.loads(open('$out/t30_bench_$r.json').read().strip().splitlines()[-1])
print('run $r value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),d['e2e'].get('legs_clips_per_s'),d['e2e'].get('legs_longest_batch_wait_ms'),'pre',round(d['e2e_precollated']['value']))"
done
timeout 600 python tools/library_path_probe.py --batches 64,256 2>>$out/t30_err.log | tee $out/t30_library_path.txt
tail -3 $out/t30_err.log
