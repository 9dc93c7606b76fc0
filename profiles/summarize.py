#!/usr/bin/env python
"""Turns ncu outputs into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/X_launches.txt
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep      > profiles/X_full.txt
  python profiles/summarize.py traffic gpurun_out/prof_X.ncu-rep   > profiles/r02_traffic.json
      (dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel of every stage, stamped with
       the sha of the kernel sources it was captured from: bench.py drops `roofline.traffic` when they differ)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        # L2 atomic throughput and residency (north_star: "L2 atomic throughput"): sectors of ATOM / RED requests, how
        # many found their line in L2, and how busy the L2 atomic units were
        "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_cas.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_atom.sum.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ours = [r for r in rows if any(k in r["Kernel Name"] for k in ("vx_", "bev_", "pillar_", "pfn_", "ingest_"))]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in ours:
        name = r["Kernel Name"].split("(")[0]
        d = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += d
        tot += d
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# %d launches of this library's kernels, %d launches in total" % (len(ours), len(rows)))
    print("%-48s %8s %12s %8s %10s" % ("kernel", "launches", "total_us", "share", "avg_us"))
    for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %8d %12.1f %7.1f%% %10.1f" % (k, c, d, 100 * d / tot, d / c))
    print("%-48s %8d %12.1f" % ("TOTAL", len(ours), tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = collections.OrderedDict()
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        seen.setdefault(name, []).append(d)
    print("# ncu --set full --clock-control none; one line block per kernel (first captured launch), n = launches captured")
    for name, ds in seen.items():
        d = ds[0]
        print("== %s (n=%d)" % (name, len(ds)))
        for k in KEYS:
            if k in idx:
                print("   %-72s %16s %s" % (k, d[idx[k]], units[idx[k]]))


def _to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def traffic(path):
    import datetime
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        b = sum(_to_bytes(d[idx[k]], units[idx[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per.setdefault(name, []).append(b)
    first = {k: v[0] for k, v in per.items()}

    def pick(sub):
        return [(k, v) for k, v in first.items() if sub in k]
    stages = {
        "scatter": {"kernel": "pillar_canvas_q_kernel", "dram_bytes_per_launch": int(sum(v for _, v in pick("pillar_canvas_q")))},
        "pillarize": {"kernel": "vx_bins_kernel<decorate>", "dram_bytes_per_launch": int(sum(v for _, v in pick("vx_bins_kernel<1")))},
        "bev": {"kernel": "bev_hist_kernel + bev_finalize_flat4_kernel",
                "dram_bytes_per_launch": int(sum(v for _, v in pick("bev_hist") + pick("bev_finalize")))},
    }
    print(json.dumps({"frames_per_step": 128, "source_sha16": bench.source_sha(),
                      "captured": datetime.date.today().isoformat(),
                      "source": "ncu --set full --clock-control none on `python bench.py --steps 1 --warmup 3 --no-e2e "
                                "--no-cpu-baseline --no-other-configs --no-verify --no-eager-ref --serial`",
                      "stages": stages}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
