"""Hot source lines of one kernel of an .ncu-rep (needs --import-source on and -lineinfo).

usage: python tools/ncu_hot_lines.py REPORT KERNEL_REGEX [launch_index] [top_n]
Prints, per CUDA source line, the share of warp-stall samples, of executed warp instructions, and the
average number of active lanes per instruction.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    idx = sys.argv[3] if len(sys.argv) > 3 else "1"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-id", "::regex:%s:%s" % (rx, idx)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    lines = []
    for r in rows:
        if len(r) < 10 or r[0] in ("", "Line No"):
            continue
        try:
            lines.append((int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_t] or 0), r[0], r[1].strip()[:110]))
        except ValueError:
            pass
    ts, ti, tt = (sum(l[k] for l in lines) for k in range(3))
    print("samples %d  warp instructions %d  lanes per instruction %.2f" % (ts, ti, tt / max(ti, 1)))
    lines.sort(reverse=True)
    for s, i, t, ln, src in lines[:top]:
        print("%5.1f%% smp %5.1f%% inst  lanes %5.1f  L%-4s %s" % (100 * s / ts, 100 * i / max(ti, 1), t / max(i, 1), ln, src))


if __name__ == "__main__":
    main()
