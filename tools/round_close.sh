#!/bin/bash
# One gpurun call that records the state of the tree on a fresh B200 at the end of a round: GPU parity suite,
# smoke, the default bench line (with the per-configuration table and the live DRAM-traffic capture), the
# reference arm, the 64-clip bench, ncu launch lists of one step for every BASELINE preset, one --set full
# capture of the step's kernels, the in-kernel GEMM timeline (variants/trace.so if present).
#
#   bash tools/build_variants.sh "trace:-DLIREC_GEMM_TRACE=1"
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round_close.sh r02'
#
# Afterwards, here:  python tools/ncu_extract.py launches gpurun_out/<tag>_launches_<preset>.csv 2
#                    python tools/ncu_extract.py full gpurun_out/<tag>_step_full.ncu-rep
tag=${1:-rNN}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py --dump_profile $out/${tag}_gemm_event_timings_per_launch.txt > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 300 $out/${tag}_bench_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --no_configs > $out/${tag}_bench_n1_steps20.json 2>> $out/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
timeout 300 python bench.py --batch 64 --steps 400 --no_cpu_baseline --no_configs --no_traffic \
  --dump_profile $out/${tag}_gemm_event_timings_per_launch_b64.txt > $out/${tag}_bench_n1_b64.json 2>> $out/${tag}_bench_n1.err
for p in int_rel_ch modalities int_rels int_ch stress; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches_$p.csv python bench.py --preset $p --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/${tag}_launches_int_rel_ch_b64.csv python bench.py --batch 64 --steps 2 --warmup 3 --ncu_window --only_value > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -c 40 \
  -o $out/${tag}_step_full -f python bench.py --steps 1 --warmup 3 --ncu_window --only_value > $out/${tag}_ncu_full.log 2>&1
if [ -f variants/trace.so ]; then
  LIREC_B200_LIB=$PWD/variants/trace.so timeout 300 python tools/gemm_trace.py > $out/${tag}_gemm_trace.txt 2>/dev/null
fi
timeout 600 python tools/stress_sweep.py > $out/${tag}_stress_sweep.txt 2>&1
ls -la $out | tail -30
