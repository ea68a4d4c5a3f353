#!/usr/bin/env python
"""Reads the ncu CSV written by the command in tools/traffic.py and records the steady-state DRAM bytes per
fused-step launch of that config in profiles/traffic.json (bench.py reports it as roofline.traffic)."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--num-agents", type=int, default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    import bench
    kw = {"num_agents": args.num_agents} if args.num_agents is not None else {}
    rows = []
    with open(args.csv) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = {}
    order = []
    for r in csv.DictReader(lines):
        name = r["Kernel Name"]
        if "mpe_kernel" not in name or not name.rstrip().endswith(", 0>(StepArgs)"):
            continue   # fused step only (mode 0); observe (mode 3) and reset kernels belong to ring construction
        k = int(r["ID"])
        if k not in per:
            per[k] = {"name": name}
            order.append(k)
        per[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[k][r["Metric Name"] + "/unit"] = r["Metric Unit"]
    launches = [per[k] for k in order]
    w = bench.scenario_world(args.scenario, kw)
    _, _, bpe, ibpe = bench.shapes_from_oracle(w.descriptor())
    R = bench.ring_size(ibpe, args.num_envs)
    assert len(launches) >= 3 * R, "need ring construction + >= 2 passes, got %d fused launches for ring %d" % (len(launches), R)
    steady = launches[2 * R:]            # drop the construction pass and the first measured pass

    def to_bytes(rec, m):
        u = rec.get(m + "/unit", "byte").lower()
        return rec[m] * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    rd = sum(to_bytes(x, "dram__bytes_read.sum") for x in steady) / len(steady)
    wr = sum(to_bytes(x, "dram__bytes_write.sum") for x in steady) / len(steady)
    hit = sum(x.get("lts__t_sector_hit_rate.pct", 0.0) for x in steady) / len(steady)
    ns = sum(x.get("gpu__time_duration.sum", 0.0) for x in steady) / len(steady)
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
        "traffic_over_algorithmic": (rd + wr) / alg, "launches_averaged": len(steady), "ring": R,
        "gpu_time_ns_under_ncu": ns,
        "how": "ncu --cache-control none --clock-control none, metrics dram__bytes_read.sum + dram__bytes_write.sum per "
               "mpe_kernel<..., kFusedStep> launch, averaged over whole ring passes after dropping the first two "
               "(tools/traffic.py); source csv: " + os.path.basename(args.csv)}
    json.dump(doc, open(args.out, "w"), indent=1)
    print(key, json.dumps(doc["detail"][key]))


if __name__ == "__main__":
    main()
