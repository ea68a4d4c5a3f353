#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2n_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2n_smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 > $O/r2n_bench_driver_style.json 2> $O/r2n_bench_driver_style.err
echo done > $O/r2n_done.txt
