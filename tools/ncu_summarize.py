#!/usr/bin/env python3
"""Turn one ncu report (gpurun_out/, scratch) into the committed summaries under profiles/.

usage: ncu_summarize.py <report.ncu-rep> <tag> <kernel substring> <games> <W> <H> <P> [--traffic] [--gym]
  report: ncu --set full --clock-control none --import-source on ... -o <report>
writes profiles/<tag>_ncu_summary.json (per-launch metrics, top stalls, dram traffic vs algorithmic bytes) and
       profiles/<tag>_regions.txt (per-region instruction / stall shares, tools/ncu_regions.py);
with --traffic also profiles/traffic.json, keyed by the hash of the sources the profiled library was built from
(bench.py prints roofline.traffic only when that hash matches the library it runs).
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes_per_env_step  # noqa: E402
from generalsreinforcementlearning_b200 import build as grl_build  # noqa: E402

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_warps']


def main():
    rep, tag, kernel = sys.argv[1:4]
    games, W, H, P = (int(v) for v in sys.argv[4:8])
    gym = "--gym" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def num(r, k):
        i = hdr.index(k)
        return float(r[i]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9,
                              's': 1}.get(units[i], 1)

    stall = [h for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h]
    out = []
    sel = [r for r in rows[2:] if kernel in r[hdr.index('Kernel Name')] or kernel in r[hdr.index('Function Name')] if 'Function Name' in hdr] \
        if 'Function Name' in hdr else [r for r in rows[2:] if kernel in r[hdr.index('Kernel Name')]]
    sel = sel or rows[2:]
    alg = algorithmic_bytes_per_env_step(W, H, P)["total"] * games
    if gym:  # slab + static + indices read; slab + obs + N*5 mask + stats + scalars written
        a = algorithmic_bytes_per_env_step(W, H, P)
        alg = (a["read"] + a["write"] - a["mask"] + 5 * W * H * P + 16 * P + 24) * games
    for r in sel:
        d = {k: f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip() for k in KEYS if k in hdr}
        top = sorted(stall, key=lambda h: -float(r[hdr.index(h)]))[:6]
        d['top_stalls_per_issue'] = {h.split('issue_stalled_')[1].split('_per_issue')[0]: round(float(r[hdr.index(h)]), 2) for h in top}
        traffic = num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum')
        dur = num(r, 'gpu__time_duration.sum')
        d['dram_bytes'] = traffic
        d['dram_GBps'] = round(traffic / dur / 1e9, 1)
        d['algorithmic_bytes'] = alg
        d['algorithmic_GBps'] = round(alg / dur / 1e9, 1)
        d['warp_inst_per_game'] = round(num(r, 'smsp__inst_executed.sum') / games, 1)
        out.append(d)
    prof = os.path.join(ROOT, "profiles")
    json.dump({"source": f"ncu --set full --clock-control none --import-source on; {games} games {W}x{H}x{P}p; kernel filter {kernel!r}; "
                         f"report {os.path.relpath(rep, ROOT)} (scratch)", "lib_source_hash": grl_build.source_hash(), "launches": out},
              open(os.path.join(prof, f"{tag}_ncu_summary.json"), "w"), indent=1)
    if "--traffic" in sys.argv:
        t = sum(d['dram_bytes'] for d in out) / len(out)
        json.dump({"kernel": f"grl_turn_kernel<{P},{W},{H}>", "dram_bytes_per_launch": t, "algorithmic_bytes_per_launch": alg,
                   "traffic_over_algorithmic": t / alg, "lib_source_hash": grl_build.source_hash(),
                   "turn_source_hash": grl_build.turn_source_hash(),
                   "source": f"profiles/{tag}_ncu_summary.json (ncu --set full, {games} games {W}x{H}x{P}p)"},
                  open(os.path.join(prof, "traffic.json"), "w"), indent=1)
    # per-region shares
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/ncu_src.csv", "w").write(srcp)
    os.makedirs("/tmp/dis", exist_ok=True)
    for f in glob.glob("/tmp/dis/*"):
        os.remove(f)
    subprocess.run(["cuobjdump", "-xelf", "all", grl_build.LIB], cwd="/tmp/dis", capture_output=True)
    # "void grl_turn_kernel<2, 15, 15, 8, 1, 1, 0>(GrlKParams, GrlGymK)" -> the mangled template arguments
    import re
    m = re.search(r"grl_turn_kernel<([^>]*)>", out[0]['Kernel Name']) if out else None
    if m:
        a = [v.strip() for v in m.group(1).split(",")]
        kernel = "grl_turn_kernelI" + "".join(f"Li{v}E" for v in a[:4]) + "".join(f"Lb{v}E" for v in a[4:6]) + "".join(f"Li{v}E" for v in a[6:]) + "E"   # <PT,TW,TH,LG, bool STEP, bool OUT, int GYM>
    for cub in sorted(glob.glob("/tmp/dis/*.cubin")):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        secs = [l for l in dis.split("\n") if l.startswith("\t.section\t.text.") and kernel in l]
        if secs:
            open("/tmp/dis/k.dis", "w").write(dis)
            name = secs[0].split(".text.")[1].split(",")[0]
            print("kernel section", name)
            ph = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), "/tmp/ncu_src.csv", "/tmp/dis/k.dis", name,
                                 str(games)], capture_output=True, text=True)
            open(os.path.join(prof, f"{tag}_regions.txt"), "w").write(ph.stdout + ph.stderr)
            print(ph.stdout[-3500:], ph.stderr[-300:])
            break
    print(json.dumps(out[0] if out else {}, indent=1))


if __name__ == "__main__":
    main()
