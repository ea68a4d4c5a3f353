#!/usr/bin/env python
"""Reads the ncu CSV written by the range-replay command in tools/traffic.py and records the steady-state DRAM bytes
per fused-step launch of that config in profiles/traffic.json (bench.py reports it as roofline.traffic)."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "%": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--launches", type=int, required=True, help="fused-step launches inside the profiled range")
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--num-agents", type=int, default=None)
    ap.add_argument("--ring", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    import bench
    kw = {"num_agents": args.num_agents} if args.num_agents is not None else {}
    with open(args.csv) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    tot = {}
    n_results = set()
    for r in csv.DictReader(lines):
        n_results.add(r["ID"])
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"].lower(), 1)
        tot.setdefault(r["Metric Name"], []).append(v)
    assert len(n_results) == 1, "expected ONE range result, got %d (did ncu run with --replay-mode range?)" % len(n_results)
    rd = sum(tot["dram__bytes_read.sum"]) / args.launches
    wr = sum(tot["dram__bytes_write.sum"]) / args.launches
    hit = tot.get("lts__t_sector_hit_rate.pct", [0.0])[0]
    ns = tot.get("gpu__time_duration.sum", [0.0])[0] / args.launches
    sh = bench.scenario_world(args.scenario, kw).native_shapes()      # device-less library handle: shapes only
    bpe, ibpe = sh.bytes_per_env_step, bench.input_bytes_from_shapes(sh)
    R = bench.ring_size(ibpe, args.num_envs, args.ring)
    alg = bpe * args.num_envs
    key = bench.traffic_key(args.scenario, kw, args.num_envs)
    try:
        doc = json.load(open(args.out))
    except Exception:  # noqa: BLE001
        doc = {}
    doc.setdefault("steady_state", {})[key] = rd + wr
    doc.setdefault("detail", {})[key] = {
        "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr, "l2_sector_hit_rate_pct": hit,
        "algorithmic_bytes_per_launch": alg, "algorithmic_read_bytes": ibpe * args.num_envs,
        "algorithmic_write_bytes": (bpe - ibpe) * args.num_envs,
        "traffic_over_algorithmic": (rd + wr) / alg, "read_over_algorithmic_read": rd / (ibpe * args.num_envs),
        "launches_in_range": args.launches, "ring": R, "range_ns_per_launch_under_ncu": ns,
        "how": "ncu --replay-mode range --cache-control none --clock-control none over whole ring passes of un-serialised "
               "fused-step launches (tools/traffic.py), dram__bytes_read.sum + dram__bytes_write.sum of the range / launches; "
               "source csv: " + os.path.basename(args.csv)}
    json.dump(doc, open(args.out, "w"), indent=1)
    print(key, json.dumps(doc["detail"][key]))


if __name__ == "__main__":
    main()
