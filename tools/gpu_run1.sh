#!/bin/bash
# GPU run 1 of round 2: parity tests, driver-style bench, CPU arm, steady-state traffic, ncu captures, SPLIT A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2a_bench_driver_style.json 2> $O/r2a_bench_driver_style.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2a_bench_reference_arm.json 2> $O/r2a_bench_reference_arm.err
timeout 900 python bench.py > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err
# steady-state DRAM traffic (cache state left alone between kernels)
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:mpe_kernel --csv --log-file $O/r2a_traffic_spread3_65536.csv python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 3 > $O/r2a_traffic_run.log 2>&1
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:mpe_kernel --csv --log-file $O/r2a_traffic_tag_262144.csv python tools/traffic.py --scenario simple_tag --num-envs 262144 --passes 3 >> $O/r2a_traffic_run.log 2>&1
# launch list of the driver-style command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2a_launches_spread3_65536.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --e2e-steps 3 > $O/r2a_launches_run.log 2>&1
# full capture of one steady-state fused step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 60 -c 1 -f -o $O/r2a_spread3_65536_full python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 2 > $O/r2a_full_run.log 2>&1
ncu -i $O/r2a_spread3_65536_full.ncu-rep --page details > $O/r2a_ncu_details_spread3_65536.txt 2>&1
# warp-pair A/B
for sp in 0 1; do
  MPE_B200_SPLIT=$sp timeout 900 python tools/sweep.py --scenarios simple_world_comm,simple_spread_n6,simple_tag,simple_spread --sizes 8192,32768,65536,131072,262144 --out $O/r2a_sweep_split$sp.jsonl > $O/r2a_sweep_split$sp.log 2>&1
done
# software-pipelined persistent kernel A/B
for tpw in 2 3 4; do
  MPE_B200_SPLIT=0 MPE_B200_PIPE=1 MPE_B200_PIPE_TPW=$tpw timeout 600 python tools/sweep.py --scenarios simple_spread,simple_tag --sizes 65536,262144 --out $O/r2a_sweep_pipe_tpw$tpw.jsonl > $O/r2a_sweep_pipe_tpw$tpw.log 2>&1
done
MPE_B200_SPLIT=0 MPE_B200_PIPE=1 MPE_B200_WPB=1 timeout 600 python tools/sweep.py --scenarios simple_spread --sizes 65536 --out $O/r2a_sweep_pipe_wpb1.jsonl > $O/r2a_sweep_pipe_wpb1.log 2>&1
MPE_B200_SPLIT=0 MPE_B200_PIPE=1 MPE_B200_WPB=4 timeout 600 python tools/sweep.py --scenarios simple_spread --sizes 65536 --out $O/r2a_sweep_pipe_wpb4.jsonl > $O/r2a_sweep_pipe_wpb4.log 2>&1
# K-step rollout kernel vs K fused steps
for sc in simple_spread simple_tag simple_world_comm; do
  timeout 600 python tools/rollout_bench.py --scenario $sc --num-envs 65536 >> $O/r2a_rollout_bench.jsonl 2>> $O/r2a_rollout_bench.err
done
# racecheck / memcheck of the warp-pair kernel
cat > /tmp/split_small.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import make_product_env
for tag, n in (("simple_world_comm", 257), ("simple_spread_n3", 100), ("simple_tag", 65)):
    env = make_product_env(tag, num_envs=n); env.reset(); nw = env.world.native
    acts = [torch.rand(n, d, device="cuda") for d in nw.act_dims]
    for _ in range(2): env.step(acts)
torch.cuda.synchronize(); print("ok")
PY
MPE_B200_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck python /tmp/split_small.py > $O/r2a_sanitizer_split.txt 2>&1
MPE_B200_SPLIT=1 timeout 900 compute-sanitizer --tool memcheck python /tmp/split_small.py >> $O/r2a_sanitizer_split.txt 2>&1
echo done > $O/r2a_done.txt
