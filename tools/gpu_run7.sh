#!/bin/bash
# GPU run 7: compact staging as the default: full parity suite, smoke, bench, register-budget A/B on top of it
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2g_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2g_bench_driver_style.json 2> $O/r2g_bench_driver_style.err
AB=$PWD/multiagent_particle_envs_b200/csrc/ab
for lib in default regs96 regs80; do
  if [ $lib = default ]; then unset MPE_B200_LIB; else export MPE_B200_LIB=$AB/libmpe_b200_$lib.so; fi
  MPE_B200_WPB=2 timeout 900 python tools/sweep.py --scenarios simple_tag,simple_world_comm,simple_spread_n6 --sizes 32768,65536,131072,262144 --seconds 0.1 --out $O/r2g_sweep_$lib.jsonl > $O/r2g_sweep_$lib.log 2>&1
done
unset MPE_B200_LIB
timeout 900 python bench.py --steps 12000 --warmup 600 --scenario simple_world_comm --num-envs 32768 --cpu-seconds 0 > $O/r2g_bench_C5shard_worldcomm_32768.json 2> $O/r2g_bench_C5.err
timeout 900 python tools/sweep.py --out $O/r2g_sweep_n1.jsonl > $O/r2g_sweep_n1.log 2>&1
echo done > $O/r2g_done.txt
