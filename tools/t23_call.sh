#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/t23_pytest.log 2>&1; tail -3 $out/t23_pytest.log
b() { # name, env, args
  env $2 timeout 300 python bench.py --only_value $3 2>>$out/t23_err.log | tail -1 > $out/t23_$1.json
  python -c "import json;d=json.load(open('$out/t23_$1.json'));print('%-28s %.0f clips/s %.4f ms gemm %.4f %s'%('$1',d['value'],d['ms_per_step'],d['gemm_ms_per_step'],d['gemm_launch_ms']))"
}
for r in 1 2; do
b base_r$r "X=1" "--steps 100 --warmup 5"
b b64_r$r "X=1" "--batch 64 --steps 400 --warmup 10"
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/t23_launches.csv python bench.py --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/t23_launches_b64.csv python bench.py --batch 64 --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
python tools/ncu_extract.py launches $out/t23_launches.csv 2 | grep -i 'expand\|total\|gemm' | head -14
python tools/ncu_extract.py launches $out/t23_launches_b64.csv 2 | grep -i 'expand\|total' | head
tail -5 $out/t23_err.log
