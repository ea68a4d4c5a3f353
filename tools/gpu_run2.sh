#!/bin/bash
# GPU run 2 of round 2: HOT kernel parity + A/B, new bench methodology, range-replay traffic, BASELINE configs at N=1, sweep
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2b_bench_driver_style.json 2> $O/r2b_bench_driver_style.err
timeout 900 python bench.py > $O/r2b_bench_default.json 2> $O/r2b_bench_default.err
MPE_B200_HOT=0 timeout 900 python bench.py --cpu-seconds 0 > $O/r2b_bench_default_general_kernel.json 2> $O/r2b_bench_default_general_kernel.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
timeout 900 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r2b_traffic_spread3_65536.csv python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 2 > $O/r2b_traffic_run.log 2>&1
timeout 900 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r2b_traffic_spread3_65536_ring31.csv python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 4 --ring 31 >> $O/r2b_traffic_run.log 2>&1
timeout 900 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r2b_traffic_tag_262144.csv python tools/traffic.py --scenario simple_tag --num-envs 262144 --passes 2 >> $O/r2b_traffic_run.log 2>&1
timeout 900 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r2b_traffic_spread6_131072.csv python tools/traffic.py --scenario simple_spread --num-agents 6 --num-envs 131072 --passes 2 >> $O/r2b_traffic_run.log 2>&1
timeout 900 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r2b_traffic_worldcomm_32768.csv python tools/traffic.py --scenario simple_world_comm --num-envs 32768 --passes 2 >> $O/r2b_traffic_run.log 2>&1
# launch list of the driver-style command + full capture of one steady-state HOT fused step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2b_launches_spread3_65536.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --e2e-steps 3 > $O/r2b_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 200 -c 1 -f -o $O/r2b_spread3_65536_hot_full python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 1 > $O/r2b_full_run.log 2>&1
ncu -i $O/r2b_spread3_65536_hot_full.ncu-rep --page details > $O/r2b_ncu_details_spread3_65536_hot.txt 2>&1
# BASELINE configs 3, 4 (one shard), 5 (one shard) on one GPU
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_tag --num-envs 262144 > $O/r2b_bench_C3_tag_262144.json 2> $O/r2b_bench_C3.err
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_spread --num-agents 6 --num-envs 131072 > $O/r2b_bench_C4shard_spread6_131072.json 2> $O/r2b_bench_C4.err
timeout 900 python bench.py --steps 12000 --warmup 600 --scenario simple_world_comm --num-envs 32768 > $O/r2b_bench_C5shard_worldcomm_32768.json 2> $O/r2b_bench_C5.err
# sweep, all four scenarios, HOT (default) and general kernel
timeout 1200 python tools/sweep.py --out $O/r2b_sweep_n1.jsonl > $O/r2b_sweep_n1.log 2>&1
MPE_B200_HOT=0 timeout 900 python tools/sweep.py --scenarios simple_spread,simple_world_comm --sizes 32768,65536,262144 --out $O/r2b_sweep_n1_general.jsonl > $O/r2b_sweep_n1_general.log 2>&1
echo done > $O/r2b_done.txt
