#!/bin/bash
# 8-GPU call: the default preset (and, with "stress" as first argument, the long-clip stress preset = BASELINE
# config 5 as well), data-parallel over 8 B200.
#   gpurun --gpus 8 --timeout 700 -- 'bash tools/call_n8_default_and_stress.sh [stress]'
out=gpurun_out; mkdir -p $out
run() { # name, args, port
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 \
    bench.py --gpus 8 --steps 20 --warmup 3 --no_configs --no_cpu_baseline --no_traffic $2 > $out/n8_$1.json 2> $out/n8_$1.err
  python - <<P
import json
try:
    d = json.loads(open('$out/n8_$1.json').read().strip().splitlines()[-1])
    print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'], 4), 'ms', 'dp_parity', d.get('dp_parity'), 'e2e', round(d['e2e']['value']), d['e2e'].get('legs_clips_per_s'), d['e2e'].get('legs_longest_batch_wait_ms'))
except Exception as e:
    print('$1 FAILED', e)
P
}
if [ "$1" = "stress" ]; then run stress "--preset stress" 29531; fi
run default "" 29532
tail -n 3 $out/n8_default.err
