#!/usr/bin/env python3
"""Share of device time per kernel over an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`), and the
timed region of bench.py's device-resident arm identified in it (the last run of K consecutive fused launches of the
headline kernel before the other shapes start).  usage: launch_shares.py <launches.csv> [K]"""
import collections
import csv
import sys


def main():
    path, K = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    launches = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit.startswith("ns") else v * 1e3 if unit.startswith("ms") else v
        launches.append((r["Kernel Name"], r["Grid Size"], us))
    tot = sum(u for _, _, u in launches)
    agg = collections.defaultdict(list)
    for k, _, u in launches:
        agg[k].append(u)
    print(f"# share of device time by kernel over {path} ({len(launches)} launches; cold-cache, serialised: under ncu launches do not overlap)")
    for k, us in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{100 * sum(us) / tot:5.1f}%  n={len(us):4d}  avg={sum(us) / len(us):8.1f} us  {k[:120]}")
    fused = "grl_turn_kernel<2, 20, 20, 32, 1, 1, 0>"
    idx = [i for i, (k, _, _) in enumerate(launches) if fused in k]
    # the first run of >= W+K consecutive fused launches holds warm-up then the timed region
    run, best = [], None
    for i in idx:
        if run and i == run[-1] + 1:
            run.append(i)
        else:
            run = [i]
        if len(run) >= K and best is None or (best and run[0] == best[0]):
            best = list(run)
    if best:
        reg = best[-K:] if len(best) >= K else best
        us = [launches[i][2] for i in reg]
        print(f"\n# first run of consecutive {fused} launches: {best[0] + 1}..{best[-1] + 1} ({len(best)} launches" + (" = the timed region of the device-resident arm" if len(best) == K else " = warm-up + timed") + "); "
              f"its last {len(reg)}: avg {sum(us) / len(us):.1f} us (min {min(us):.1f}, max {max(us):.1f}), grid {launches[reg[0]][1]}")


if __name__ == "__main__":
    main()
