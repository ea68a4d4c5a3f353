#!/bin/bash
# GPU run 10: warps per block at the BASELINE shard sizes of the heavy scenarios (final kernels)
set -u
O=gpurun_out; mkdir -p $O
for wpb in 1 2 4; do
  MPE_B200_WPB=$wpb timeout 600 python tools/sweep.py --scenarios simple_world_comm --sizes 32768,65536 --seconds 0.15 --out $O/r2j_wc_wpb$wpb.jsonl > $O/r2j_wc_wpb$wpb.log 2>&1
  MPE_B200_WPB=$wpb timeout 600 python tools/sweep.py --scenarios simple_spread_n6 --sizes 65536,131072 --seconds 0.15 --out $O/r2j_s6_wpb$wpb.jsonl > $O/r2j_s6_wpb$wpb.log 2>&1
  MPE_B200_WPB=$wpb timeout 600 python tools/sweep.py --scenarios simple_tag --sizes 65536,262144 --seconds 0.15 --out $O/r2j_tag_wpb$wpb.jsonl > $O/r2j_tag_wpb$wpb.log 2>&1
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j_smoke.log 2>&1
echo done > $O/r2j_done.txt
