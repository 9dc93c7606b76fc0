#!/usr/bin/env python
"""Per-CUDA-line stall samples of one kernel from an ncu report (needs -lineinfo).

  python profiles/srclines.py gpurun_out/prof_X.ncu-rep vx_bins [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg, total, seen_kernel = None, None, {}, 0, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or not r[0].isdigit():
            continue
        try:
            samples = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        stalls = {}
        for h, i in hdr.items():
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                if v:
                    stalls[h[6:]] = v
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, {}])
        a[0] += samples
        for k, v in stalls.items():
            a[1][k] = a[1].get(k, 0) + v
        total += samples
    print("# %s: %d samples" % (pat, total))
    for (f, ln, src), (s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%%  %s:%d  %s\n         [%s]" % (100.0 * s / max(total, 1), f, ln, src, tops))


if __name__ == "__main__":
    main()
