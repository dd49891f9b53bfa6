#!/bin/bash
# Copy / summarise what tools/round_close.sh left in gpurun_out/ into profiles/ (tracked).  Run here, no GPU.
#   bash tools/collect_profiles.sh r02
tag=${1:-rNN}
src=gpurun_out
dst=profiles
for f in bench_n1.json bench_n1_b64.json bench_n1_steps20.json bench_reference_arm.json \
         gemm_event_timings_per_launch.txt gemm_event_timings_per_launch_b64.txt gemm_trace.txt stress_sweep.txt \
         pytest_gpu.log smoke.log; do
  [ -f $src/${tag}_$f ] && cp $src/${tag}_$f $dst/${tag}_$f
done
for p in int_rel_ch modalities int_rels int_ch stress int_rel_ch_b64; do
  [ -f $src/${tag}_launches_$p.csv ] && python tools/ncu_extract.py launches $src/${tag}_launches_$p.csv 2 > $dst/${tag}_launches_one_step_$p.txt
done
if [ -f $src/${tag}_step_full.ncu-rep ]; then
  python tools/ncu_extract.py full $src/${tag}_step_full.ncu-rep > $dst/${tag}_step_ncu_full.txt
  for k in expand_bwd_t_kernel expand_fwd_kernel adam_kernel; do
    python tools/ncu_stalls.py $src/${tag}_step_full.ncu-rep $k 16 | awk '!seen[$0]++' > $dst/${tag}_stalls_$k.txt
  done
fi
python tools/res_usage.py > $dst/${tag}_resource_usage.txt 2>/dev/null
ls $dst | grep "^${tag}_" | wc -l
