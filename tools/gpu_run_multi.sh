#!/bin/bash
# multi-GPU measurements of round 2 (run with: gpurun --gpus N -- 'bash tools/gpu_run_multi.sh N')
# N=8: BASELINE configs 4 (simple_spread N=6, 1 048 576 worlds = 131 072 per GPU) and 5 (simple_world_comm, 262 144
# worlds = 32 768 per GPU) at their stated GPU count, the headline at N, and the 1k-1M sweep of spread / tag.
set -u
N=$1
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/r2_topo_n$N.txt 2>&1
timeout 600 $TR --master-port 29431 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_headline_driver_style_n$N.json 2> $O/r2_bench_headline_driver_style_n$N.err
timeout 600 $TR --master-port 29432 bench.py --gpus $N --steps 6000 --warmup 300 > $O/r2_bench_headline_n$N.json 2> $O/r2_bench_headline_n$N.err
if [ "$N" = "8" ]; then
  timeout 900 $TR --master-port 29433 bench.py --gpus $N --steps 3000 --warmup 100 --scenario simple_spread --num-agents 6 --num-envs 131072 > $O/r2_bench_C4_spread6_1M_n8.json 2> $O/r2_bench_C4_n8.err
  timeout 900 $TR --master-port 29434 bench.py --gpus $N --steps 6000 --warmup 300 --scenario simple_world_comm --num-envs 32768 > $O/r2_bench_C5_worldcomm_262k_n8.json 2> $O/r2_bench_C5_n8.err
fi
timeout 900 $TR --master-port 29435 tools/sweep.py --scenarios simple_spread,simple_tag --sizes 1024,4096,16384,65536,262144,1048576 --out $O/r2_sweep_n$N.jsonl > $O/r2_sweep_n$N.log 2>&1
echo done > $O/r2_multi_done_n$N.txt
