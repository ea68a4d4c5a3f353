#!/bin/bash
# GPU run 9: packed fp32x2 (FADD2 / FMUL2 / FFMA2) arithmetic: parity (bit-exactness tests) + A/B against the scalar build
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2i_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2i_pytest_gpu.log
AB=$PWD/multiagent_particle_envs_b200/csrc/ab
for lib in packed scalar packed2 scalar2; do
  case $lib in scalar*) export MPE_B200_LIB=$AB/libmpe_b200_scalar.so;; *) unset MPE_B200_LIB;; esac
  timeout 900 python tools/sweep.py --scenarios simple_spread,simple_tag,simple_world_comm,simple_spread_n6 --sizes 32768,65536,262144 --seconds 0.15 --out $O/r2i_sweep_$lib.jsonl > $O/r2i_sweep_$lib.log 2>&1
done
unset MPE_B200_LIB
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 > $O/r2i_bench_driver_style.json 2> $O/r2i_bench_driver_style.err
echo done > $O/r2i_done.txt
