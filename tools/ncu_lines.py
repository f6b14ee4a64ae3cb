"""Hot CUDA source lines of one kernel launch of an .ncu-rep (captured with --import-source on, built with -lineinfo).

usage: python tools/ncu_lines.py REPORT KERNEL_REGEX [top_n]
Per source line: share of warp-stall samples (and the long-scoreboard part of it), share of executed warp
instructions, average active lanes per instruction.
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + rx, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    tot = [0, 0, 0]
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        try:
            smp = int(r[hdr.index("# Samples")] or 0)
            ins = int(r[hdr.index("Instructions Executed")] or 0)
            thr = int(r[hdr.index("Thread Instructions Executed")] or 0)
            lsb = int(r[hdr.index("stall_long_sb")] or 0)
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((cur_file, int(r[0]), r[1].strip()[:96]), [0, 0, 0, 0])
        a[0] += smp; a[1] += ins; a[2] += thr; a[3] += lsb
        tot[0] += smp; tot[1] += ins; tot[2] += thr
    print("samples %d  warp instructions %d  lanes per instruction %.2f" % (tot[0], tot[1], tot[2] / max(tot[1], 1)))
    for (f, ln, src), (smp, ins, thr, lsb) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp (%4.1f%% long_sb) %5.1f%% inst lanes %5.1f  %s:%d  %s"
              % (100 * smp / max(tot[0], 1), 100 * lsb / max(tot[0], 1), 100 * ins / max(tot[1], 1), thr / max(ins, 1), f, ln, src))


if __name__ == "__main__":
    main()
