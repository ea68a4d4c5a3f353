#!/bin/bash
# GPU run 6: launch geometry x PDL release point, and compact observation staging (more co-residency for the next grid)
set -u
O=gpurun_out; mkdir -p $O
AB=$PWD/multiagent_particle_envs_b200/csrc/ab
for lib in default compact compact72; do
  if [ $lib = default ]; then unset MPE_B200_LIB; else export MPE_B200_LIB=$AB/libmpe_b200_$lib.so; fi
  for wpb in 1 2 4; do
    if [ $lib = compact72 ] && [ $wpb = 4 ]; then continue; fi
    for pdl in 3 5 2; do
      MPE_B200_WPB=$wpb MPE_B200_PDL=$pdl timeout 300 python tools/sweep.py --scenarios simple_spread --sizes 65536 --seconds 0.1 > $O/tmp_sweep.log 2>&1
      echo "{\"lib\": \"$lib\", \"wpb\": $wpb, \"pdl\": $pdl, \"point\": $(grep '^{' $O/tmp_sweep.log | tail -1)}" >> $O/r2f_geometry_spread3_65536.jsonl
    done
  done
  timeout 600 python tools/sweep.py --scenarios simple_spread,simple_tag,simple_world_comm --sizes 32768,65536,262144 --seconds 0.1 --out $O/r2f_sweep_$lib.jsonl > $O/r2f_sweep_$lib.log 2>&1
done
unset MPE_B200_LIB
echo done > $O/r2f_done.txt
