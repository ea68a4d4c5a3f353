#!/bin/bash
# GPU run 8: the final kernels of round 2 (HOT + compact staging + 80-register variant): parity, bench of every BASELINE
# config on one GPU, sweep, ncu captures
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2h_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2h_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2h_bench_driver_style.json 2> $O/r2h_bench_driver_style.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2h_bench_reference_arm.json 2> $O/r2h_bench_reference_arm.err
timeout 900 python bench.py > $O/r2h_bench_default.json 2> $O/r2h_bench_default.err
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_tag --num-envs 262144 > $O/r2h_bench_C3_tag_262144.json 2> $O/r2h_bench_C3.err
timeout 900 python bench.py --steps 6000 --warmup 300 --scenario simple_spread --num-agents 6 --num-envs 131072 > $O/r2h_bench_C4shard_spread6_131072.json 2> $O/r2h_bench_C4.err
timeout 900 python bench.py --steps 12000 --warmup 600 --scenario simple_world_comm --num-envs 32768 > $O/r2h_bench_C5shard_worldcomm_32768.json 2> $O/r2h_bench_C5.err
timeout 1200 python tools/sweep.py --out $O/r2h_sweep_n1.jsonl > $O/r2h_sweep_n1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2h_launches_spread3_65536.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --e2e-steps 3 > $O/r2h_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 200 -c 1 -f -o $O/r2h_spread3_65536_full python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 1 > $O/r2h_full_run.log 2>&1
ncu -i $O/r2h_spread3_65536_full.ncu-rep --page details > $O/r2h_ncu_details_spread3_65536.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 40 -c 1 -f -o $O/r2h_tag_262144_full python tools/traffic.py --scenario simple_tag --num-envs 262144 --passes 1 > $O/r2h_full_run_tag.log 2>&1
ncu -i $O/r2h_tag_262144_full.ncu-rep --page details > $O/r2h_ncu_details_tag_262144.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 200 -c 1 -f -o $O/r2h_worldcomm_32768_full python tools/traffic.py --scenario simple_world_comm --num-envs 32768 --passes 1 > $O/r2h_full_run_wc.log 2>&1
ncu -i $O/r2h_worldcomm_32768_full.ncu-rep --page details > $O/r2h_ncu_details_worldcomm_32768.txt 2>&1
for sc in simple_spread simple_tag simple_world_comm; do
  timeout 600 python tools/rollout_bench.py --scenario $sc --num-envs 65536 >> $O/r2h_rollout_bench.jsonl 2>> $O/r2h_rollout_bench.err
done
echo done > $O/r2h_done.txt
