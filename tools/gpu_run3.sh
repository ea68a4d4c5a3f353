#!/bin/bash
# GPU run 3 of round 2: bench with lead-in steps + copy probe, wide-register A/B, world_comm capture
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2c_bench_driver_style.json 2> $O/r2c_bench_driver_style.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 > $O/r2c_bench_driver_style_repeat.json 2> $O/r2c_bench_driver_style_repeat.err
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --cpu-seconds 0 > $O/r2c_bench_k200.json 2> $O/r2c_bench_k200.err
timeout 900 python bench.py > $O/r2c_bench_default.json 2> $O/r2c_bench_default.err
for lib in default wide; do
  if [ $lib = wide ]; then export MPE_B200_LIB=$PWD/multiagent_particle_envs_b200/csrc/ab/libmpe_b200_wide.so; else unset MPE_B200_LIB; fi
  timeout 900 python tools/sweep.py --scenarios simple_world_comm,simple_spread_n6,simple_tag --sizes 16384,32768,65536,262144 --out $O/r2c_sweep_regs_$lib.jsonl > $O/r2c_sweep_regs_$lib.log 2>&1
done
unset MPE_B200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpe_kernel -s 200 -c 1 -f -o $O/r2c_worldcomm_32768_hot_full python tools/traffic.py --scenario simple_world_comm --num-envs 32768 --passes 1 > $O/r2c_full_run.log 2>&1
ncu -i $O/r2c_worldcomm_32768_hot_full.ncu-rep --page details > $O/r2c_ncu_details_worldcomm_32768_hot.txt 2>&1
echo done > $O/r2c_done.txt
