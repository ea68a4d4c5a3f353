#!/bin/bash
# GPU run 12: closed-loop policy rollout kernel: parity tests + timing against the graphed torch rollout
set -u
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "policy_rollout or open_loop" ) > $O/r2l_pytest_policy.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest_policy.log
for sc in simple_spread simple_tag; do for h in 32 64; do
  timeout 600 python tools/policy_rollout_bench.py --scenario $sc --hidden $h >> $O/r2l_policy_rollout_bench.jsonl 2>> $O/r2l_policy_rollout_bench.err
done; done
echo done > $O/r2l_done.txt
