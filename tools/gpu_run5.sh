#!/bin/bash
# GPU run 5: warp-pair kernel v2 (contact forces split between the two warps): parity + A/B
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alternative or golden_fixtures_single or fused_step_equals" ) > $O/r2e_pytest_split.log 2>&1
echo "pytest rc=$?" >> $O/r2e_pytest_split.log
for sp in 0 1; do
  MPE_B200_SPLIT=$sp MPE_B200_HOT=0 timeout 900 python tools/sweep.py --scenarios simple_world_comm,simple_spread_n6,simple_tag,simple_spread --sizes 8192,16384,32768,65536,131072 --out $O/r2e_sweep_split$sp.jsonl > $O/r2e_sweep_split$sp.log 2>&1
done
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
MPE_B200_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck python /tmp/split_small.py > $O/r2e_sanitizer_split.txt 2>&1
MPE_B200_SPLIT=1 timeout 900 compute-sanitizer --tool memcheck python /tmp/split_small.py >> $O/r2e_sanitizer_split.txt 2>&1
echo done > $O/r2e_done.txt
