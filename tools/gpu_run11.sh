#!/bin/bash
# GPU run 11: final validation of HEAD (packed arithmetic, compact staging, 80-register variant, 1 warp per block)
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2k_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2k_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2k_bench_reference_arm.json 2> $O/r2k_bench_reference_arm.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2k_bench_driver_style.json 2> $O/r2k_bench_driver_style.err
timeout 900 python bench.py > $O/r2k_bench_default.json 2> $O/r2k_bench_default.err
timeout 900 python bench.py --steps 12000 --warmup 600 --scenario simple_world_comm --num-envs 32768 --cpu-seconds 0 > $O/r2k_bench_C5shard_worldcomm_32768.json 2> $O/r2k_bench_C5.err
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_tag --num-envs 262144 --cpu-seconds 0 > $O/r2k_bench_C3_tag_262144.json 2> $O/r2k_bench_C3.err
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_spread --num-agents 6 --num-envs 131072 --cpu-seconds 0 > $O/r2k_bench_C4shard_spread6_131072.json 2> $O/r2k_bench_C4.err
timeout 1200 python tools/sweep.py --out $O/r2k_sweep_n1.jsonl > $O/r2k_sweep_n1.log 2>&1
echo done > $O/r2k_done.txt
