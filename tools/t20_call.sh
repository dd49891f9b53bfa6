#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/t20_pytest.log 2>&1; tail -3 $out/t20_pytest.log
b() { # name, env, args
  env $2 timeout 300 python bench.py --only_value $3 2>>$out/t20_err.log | tail -1 > $out/t20_$1.json
  python -c "import json;d=json.load(open('$out/t20_$1.json'));print('%-28s %.0f clips/s %.4f ms gemm %.4f'%('$1',d['value'],d['ms_per_step'],d['gemm_ms_per_step']))"
}
for r in 1 2; do
b base_r$r "X=1" "--steps 100 --warmup 5"
b ovl_co_r$r "LIREC_DP_CORESIDENT=1" "--steps 100 --warmup 5 --overlap_adam"
b ovl_noco_r$r "LIREC_DP_CORESIDENT=0" "--steps 100 --warmup 5 --overlap_adam"
b b64_base_r$r "X=1" "--batch 64 --steps 400 --warmup 10"
b b64_ovl_co_r$r "LIREC_DP_CORESIDENT=1" "--batch 64 --steps 400 --warmup 10 --overlap_adam"
b b64_ovl_noco_r$r "LIREC_DP_CORESIDENT=0" "--batch 64 --steps 400 --warmup 10 --overlap_adam"
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/t20_launches.csv python bench.py --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/t20_launches_b64.csv python bench.py --batch 64 --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
timeout 300 python tools/stress_sweep.py 2>&1 | grep -i 'roi\|gather' > $out/t20_stress_roi.txt; cat $out/t20_stress_roi.txt
