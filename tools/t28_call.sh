#!/bin/bash
out=gpurun_out; mkdir -p $out
b() { # name, env
  env $2 LIREC_BENCH_DEBUG=1 timeout 600 python bench.py --steps 100 --warmup 5 --no_configs --no_cpu_baseline --no_traffic > $out/t28_$1.json 2>$out/t28_$1.err
  python -c "
import json;d=json.loads(open('$out/t28_$1.json').read().strip().splitlines()[-1])
print('$1','value',round(d['value']),'e2e',round(d['e2e']['value']),'pre',round(d['e2e_precollated']['value']),'streamed',round(d['e2e_streamed']['value']))"
  grep "e2e loader leg" $out/t28_$1.err | tail -1
}
b new "X=1"
b flat "LIREC_GATHER_FLAT=1"
b nofreeze "LIREC_BENCH_NO_FREEZE=1"
b flat_nofreeze "LIREC_GATHER_FLAT=1 LIREC_BENCH_NO_FREEZE=1"
b new2 "X=1"
