"""Per-CUDA-source-line share of stall samples and executed instructions from an ncu report (read on the CPU box).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top]

Uses `ncu --page source --print-source cuda,sass`, where every CUDA line row carries the aggregate of its SASS rows.
Needs -lineinfo at compile time and --import-source on at capture time.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) >= 2 and r[0] == "File Path":
            cur, hdr = r[1], None
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")

            def num(x):
                try:
                    return float(x or 0)
                except ValueError:
                    return 0.0
            s, n = num(r[i_s]), num(r[i_n])
            if s > 0 or n > 0:
                agg.append((s, n, cur.split("/")[-1], r[0], r[1].strip()[:100]))
    tot, totn = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    print("# %s  kernel ~ %s: %d stall samples, %d warp instructions" % (rep, kern, tot, totn))
    for a in sorted(agg, reverse=True)[:top]:
        print("%5.1f%% samples %5.1f%% inst  %s:%s  %s" % (100 * a[0] / tot, 100 * a[1] / totn, a[2], a[3], a[4]))


if __name__ == "__main__":
    main()
