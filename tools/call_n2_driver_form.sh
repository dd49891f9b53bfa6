#!/bin/bash
# the driver's own form at N = 2: default bench (configs table included) under torchrun
out=gpurun_out; mkdir -p $out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $out/driver_form_n2_default.json 2> $out/driver_form_n2_default.err
echo "rc $?"
python - <<P
import json
d = json.loads(open('$out/driver_form_n2_default.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'], 4), d.get('dp_parity'), 'e2e', round(d['e2e']['value']), d['e2e'].get('legs_clips_per_s'))
for c in d.get('configs') or []:
    print('  ', c.get('preset'), c.get('clips_per_gpu'), round(c.get('value', 0)), c.get('ms_per_step'), c.get('error'))
P
tail -n 5 $out/driver_form_n2_default.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
  bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $out/driver_form_n2_reference.json 2> $out/driver_form_n2_reference.err
echo "reference arm rc $?"; tail -c 400 $out/driver_form_n2_reference.json
