#!/bin/bash
# GPU run 13: last validation of HEAD (all kernels incl. the policy rollout): full GPU suite, smoke, both bench arms
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2m_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2m_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2m_smoke.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench_reference_arm.json 2> $O/r2m_bench_reference_arm.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench_driver_style.json 2> $O/r2m_bench_driver_style.err
echo done > $O/r2m_done.txt
