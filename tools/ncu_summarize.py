#!/usr/bin/env python3
"""Turn a round's ncu artefacts (gpurun_out/, scratch) into the committed summaries under profiles/.

usage: ncu_summarize.py <tag> [games W H P]
  gpurun_out/prof_<tag>.ncu-rep     (ncu --set full ... --import-source on)
  gpurun_out/launches_<tag>.csv     (ncu --metrics gpu__time_duration.sum launch list)
writes profiles/<tag>_ncu_full_summary.json, <tag>_launches.csv, <tag>_launch_shares.txt,
       <tag>_phases.txt and profiles/traffic.json (dram bytes per launch of the turn kernel).
"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes_per_env_step  # noqa: E402

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct']


def main():
    tag = sys.argv[1]
    games, W, H, P = (int(v) for v in sys.argv[2:6]) if len(sys.argv) >= 6 else (65536, 20, 20, 2)
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def num(r, k):
        i = hdr.index(k)
        return float(r[i]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(units[i], 1)

    stall = [h for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h]
    out, traffic = [], []
    for r in rows[2:]:
        d = {k: f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip() for k in KEYS if k in hdr}
        top = sorted(stall, key=lambda h: -float(r[hdr.index(h)]))[:6]
        d['top_stalls_per_issue'] = {h.split('issue_stalled_')[1].split('_per_issue')[0]: round(float(r[hdr.index(h)]), 2) for h in top}
        out.append(d)
        traffic.append(num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum'))
    prof = os.path.join(ROOT, "profiles")
    json.dump({"source": f"ncu --set full --clock-control none --import-source on -k regex:grl_turn_kernel -s 30 -c 2, "
                         f"bench.py --steps 20 --warmup 3 --quick; {games} games {W}x{H}x{P}p; report gpurun_out/prof_{tag}.ncu-rep (scratch)",
               "launches": out}, open(os.path.join(prof, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    alg = algorithmic_bytes_per_env_step(W, H, P)["total"] * games
    json.dump({"kernel": f"grl_turn_kernel<{P},{W},{H},true,true>", "dram_bytes_per_launch": sum(traffic) / len(traffic),
               "dram_bytes_read": num(rows[2], 'dram__bytes_read.sum'), "dram_bytes_write": num(rows[2], 'dram__bytes_write.sum'),
               "algorithmic_bytes_per_launch": alg, "traffic_over_algorithmic": sum(traffic) / len(traffic) / alg,
               "source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full, {games} games {W}x{H}x{P}p)"},
              open(os.path.join(prof, "traffic.json"), "w"), indent=1)
    # launch list
    src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(prof, f"{tag}_launches.csv"))
        lr = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
        agg, cnt = collections.Counter(), collections.Counter()
        for r in lr:
            agg[r['Kernel Name'][:70]] += float(r['Metric Value'])
            cnt[r['Kernel Name'][:70]] += 1
        tot = sum(agg.values())
        with open(os.path.join(prof, f"{tag}_launch_shares.txt"), "w") as f:
            f.write(f"# share of device time by kernel over the ncu launch list ({len(lr)} launches; cold-cache, serialised)\n")
            for k, v in agg.most_common():
                f.write(f"{v / tot * 100:5.1f}%  n={cnt[k]:4d}  avg={v / cnt[k] / 1e3:8.1f} us  {k}\n")
    # per-phase shares
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/ncu_src.csv", "w").write(srcp)
    os.makedirs("/tmp/dis", exist_ok=True)
    for f in glob.glob("/tmp/dis/*"):
        os.remove(f)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "generalsreinforcementlearning_b200", "csrc", "libgrlcuda.so")],
                   cwd="/tmp/dis", capture_output=True)
    cub = [f for f in glob.glob("/tmp/dis/grl_kernels.sm_100a.cubin")][0]
    open("/tmp/dis/k.dis", "w").write(subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout)
    mangled = f"_Z15grl_turn_kernelILi{2 if P <= 2 else 4}ELi{W}ELi{H}ELi32ELb1ELb1ELb0EEv10GrlKParams7GrlGymK"
    ph = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), "/tmp/ncu_src.csv", "/tmp/dis/k.dis", mangled,
                         os.path.join(ROOT, "generalsreinforcementlearning_b200", "csrc", "grl_kernels.cu"), str(games)],
                        capture_output=True, text=True)
    open(os.path.join(prof, f"{tag}_phases.txt"), "w").write(ph.stdout + ph.stderr)
    print(json.dumps(out[0], indent=1))
    print(open(os.path.join(prof, "traffic.json")).read())
    print(ph.stdout[-900:], ph.stderr[-300:])


if __name__ == "__main__":
    main()
