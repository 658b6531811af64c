"""Per-layer table from scripts/ncu_replay_metrics.sh output: `python scripts/layer_table.py <csv> <frames>`."""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from analyze_launches import name_ops  # noqa: E402


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1].split("<")[0],
                                           "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(byid.values())


def main():
    path, B = sys.argv[1], int(sys.argv[2])
    import json
    ops_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(path)), "ops.json")
    ks = load(path)
    tot = sum(k["gpu__time_duration.sum"] for k in ks) / 1e3
    agg = collections.defaultdict(float)
    for k in ks:
        agg[k["name"]] += k["gpu__time_duration.sum"] / 1e3
    print(f"{len(ks)} launches, {tot:.1f} us for {B} frames = {tot / B:.2f} us/frame:",
          ", ".join(f"{n} {v:.1f}" for n, v in sorted(agg.items(), key=lambda x: -x[1])))
    convs = [k for k in ks if k["name"].startswith("conv_") or k["name"].startswith("sppf")]
    L = name_ops(json.load(open(ops_path)))
    assert len(convs) == len(L), (len(convs), len(L))
    ctot = sum(k["gpu__time_duration.sum"] for k in convs) / 1e3
    print(f"{'layer':15s} {'kernel':11s} {'hw':>4s} {'cin':>4s} {'cout':>4s} k s {'us':>7s} {'share':>6s} {'TFLOP/s':>8s} {'tensor%':>7s} {'DRAM MB':>8s} {'L2 MB':>8s} {'smem KB':>7s} grid")
    for k, d in zip(convs, L):
        name, hw, cin, cout, kk, s, flf = d
        t = k["gpu__time_duration.sum"] / 1e3
        if name == "POOL":
            print(f"{name:15s} {k['name'][:11]:11s} {'':21s} {t:7.1f} {100 * t / ctot:5.1f}%")
            continue
        fl = flf * B
        print(f"{name:15s} {k['name'][:11]:11s} {hw:4d} {cin:4d} {cout:4d} {kk} {s} {t:7.1f} {100 * t / ctot:5.1f}% {fl / t / 1e6:8.1f} "
              f"{k.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):7.1f} "
              f"{(k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0)) / 1e6:8.1f} {k.get('lts__t_bytes.sum', 0) / 1e6:8.1f} "
              f"{k.get('launch__shared_mem_per_block_dynamic', 0) / 1e3:7.0f} {k['grid']}x{k['block']}")


if __name__ == "__main__":
    main()
