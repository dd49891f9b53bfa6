#!/bin/bash
# last check of a build on a fresh box, in the driver's order: GPU suite, smoke, reference arm, bench (driver form)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/final_pytest.log 2>&1; tail -2 $out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $out/final_reference.json 2>$out/final_reference.err; tail -c 300 $out/final_reference.json; echo
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/final_bench.json 2>$out/final_bench.err
python -c "
import json;d=json.loads(open('$out/final_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),d['e2e'].get('legs_clips_per_s'),d['e2e'].get('legs_longest_batch_wait_ms'),'pre',round(d['e2e_precollated']['value']),'clocks',d['clocks']['sm_mhz'],d['clocks']['samples'],d['clocks']['reasons'],'frac',round(d['roofline']['frac'],3),d['roofline']['peak_source'],'launches',d['gpu_launches'],'configs',len(d.get('configs') or []))"
tail -3 $out/final_bench.err
