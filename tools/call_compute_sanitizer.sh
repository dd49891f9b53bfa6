#!/bin/bash
# compute-sanitizer over the single-GPU suite and smoke (round 2 build)
out=gpurun_out; mkdir -p $out
f=$out/r02_compute_sanitizer.txt
echo "# compute-sanitizer on B200 (round 2 build)" > $f
echo "## --tool memcheck python -m pytest tests -m gpu -q  (whole single-GPU suite)" >> $f
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q 2>&1 | grep -v "^\.\|^$" | tail -25 >> $f
echo "## --tool memcheck: __graft_entry__.smoke()" >> $f
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -6 >> $f
echo "## --tool racecheck python -m pytest tests/test_pooling_gpu.py tests/test_rows_gpu.py tests/test_resident_gpu.py::test_gather_rows_is_bit_exact -m gpu" >> $f
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_pooling_gpu.py tests/test_rows_gpu.py tests/test_resident_gpu.py::test_gather_rows_is_bit_exact -m gpu -q 2>&1 | grep -v "^\.\|^$" | tail -15 >> $f
tail -40 $f
