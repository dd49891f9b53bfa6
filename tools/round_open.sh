#!/bin/bash
# One gpurun call that re-establishes the state of the tree on a fresh B200 at the start of a round:
# GPU parity suite, smoke, the default bench line and the reference arm, the 64-clip bench, the ncu launch
# list of one step and one --set full capture of the step's kernels.  Everything lands in gpurun_out/.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_open.sh r02'
#
# Afterwards, here:  python tools/ncu_extract.py launches gpurun_out/<tag>_launches.csv 2
#                    python tools/ncu_extract.py full gpurun_out/<tag>_step_full.ncu-rep
# and copy the summaries to profiles/.
tag=${1:-rNN}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 600 python bench.py --dump_profile $out/${tag}_gemm_event_timings_per_launch.txt > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 600 $out/${tag}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
timeout 300 python bench.py --batch 64 --steps 400 --no_cpu_baseline --dump_profile $out/${tag}_gemm_event_timings_per_launch_b64.txt \
  > $out/${tag}_bench_n1_b64.json 2>> $out/${tag}_bench_n1.err
# ncu: launch list of one step (cheap), then the full set over the same window
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --ncu_window --no_cpu_baseline > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -c 40 \
  -o $out/${tag}_step_full -f python bench.py --steps 1 --warmup 3 --ncu_window --no_cpu_baseline > $out/${tag}_ncu_full.log 2>&1
python tools/loader_probe.py > $out/${tag}_loader_probe.txt 2>&1
timeout 600 python tools/e2e_loader_probe.py > $out/${tag}_e2e_loader_probe.json 2> $out/${tag}_e2e_loader_probe.err; tail -c 700 $out/${tag}_e2e_loader_probe.json
ls -la $out | tail -20
