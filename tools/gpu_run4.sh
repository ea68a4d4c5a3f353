#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for w in 0 1; do
  MPE_BENCH_TLB_WARM=$w timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 --e2e-steps 3 > $O/r2d_bench_k20_tlb$w.json 2> $O/r2d_bench_k20_tlb$w.err
  MPE_BENCH_TLB_WARM=$w timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 --e2e-steps 3 > $O/r2d_bench_k20_tlb${w}_b.json 2>> $O/r2d_bench_k20_tlb$w.err
done
echo done > $O/r2d_done.txt
