#!/bin/bash
# 8-GPU call: the long-clip stress preset (BASELINE config 5) and the default preset, data-parallel over 8 B200.
out=gpurun_out; mkdir -p $out
run() { # name, args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 \
    bench.py --gpus 8 --steps 20 --warmup 3 --no_configs --no_cpu_baseline --no_traffic $2 > $out/n8_$1.json 2> $out/n8_$1.err
  python - <<P
import json
try:
    d = json.loads(open('$out/n8_$1.json').read().strip().splitlines()[-1])
    print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'], 4), 'ms', 'dp_parity', d.get('dp_parity'), 'e2e', round(d['e2e']['value']))
except Exception as e:
    print('$1 FAILED', e)
P
}
run n8_stress "--preset stress" 29531
run n8 "" 29532
tail -3 $out/n8_n8_stress.err $out/n8_n8.err
