#!/bin/bash
out=gpurun_out; mkdir -p $out
b() { # name, env, args
  env $2 timeout 300 python bench.py --only_value $3 2>>$out/t22_err.log | tail -1 > $out/t22_$1.json
  python -c "import json;d=json.load(open('$out/t22_$1.json'));print('%-28s %.0f clips/s %.4f ms gemm %.4f'%('$1',d['value'],d['ms_per_step'],d['gemm_ms_per_step']))"
}
for r in 1 2; do
b base_r$r "X=1" "--steps 100 --warmup 5"
b ovl_r$r "X=1" "--steps 100 --warmup 5 --overlap_adam"
b b64_base_r$r "X=1" "--batch 64 --steps 400 --warmup 10"
b b64_ovl_r$r "X=1" "--batch 64 --steps 400 --warmup 10 --overlap_adam"
b b64_base_nosampler_r$r "LIREC_BENCH_NO_SAMPLER=1" "--batch 64 --steps 400 --warmup 10"
done
for B in 1024 64; do
timeout 300 python tools/overlap_probe.py --batch $B --no_overlap 2>>$out/t22_err.log | tee -a $out/t22_overlap_probe.txt
timeout 300 python tools/overlap_probe.py --batch $B 2>>$out/t22_err.log | tee -a $out/t22_overlap_probe.txt
done
timeout 300 python bench.py --batch 64 --steps 400 --no_cpu_baseline --no_configs --no_traffic > $out/t22_bench_b64_full.json 2>>$out/t22_err.log
tail -c 400 $out/t22_bench_b64_full.json
tail -5 $out/t22_err.log
