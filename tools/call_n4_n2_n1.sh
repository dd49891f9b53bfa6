#!/bin/bash
# 4-GPU box: the default bench at 4 and at 2 ranks (the driver's scaling run covers N = 1, 2, 4, 8)
out=gpurun_out; mkdir -p $out
run() { # name, nproc, port
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 \
    bench.py --gpus $2 --steps 20 --warmup 3 --no_configs --no_cpu_baseline --no_traffic > $out/scale_$1.json 2> $out/scale_$1.err
  python - <<P
import json
try:
    d = json.loads(open('$out/scale_$1.json').read().strip().splitlines()[-1])
    print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'], 4), 'ms', 'dp_parity', d.get('dp_parity'), 'e2e', round(d['e2e']['value']), d['e2e'].get('legs_clips_per_s'), d['e2e'].get('legs_longest_batch_wait_ms'))
except Exception as e:
    print('$1 FAILED', e)
P
}
run n4 4 29541
run n2 2 29542
timeout 300 python bench.py --steps 20 --warmup 3 --no_configs --no_cpu_baseline --no_traffic > $out/scale_n1.json 2> $out/scale_n1.err
python -c "
import json;d=json.loads(open('$out/scale_n1.json').read().strip().splitlines()[-1])
print('n1',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),d['e2e'].get('legs_clips_per_s'),d['e2e'].get('legs_longest_batch_wait_ms'))"
tail -n 3 $out/scale_n4.err; tail -n 3 $out/scale_n2.err
