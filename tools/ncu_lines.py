"""Summarise an ncu report's source page: warp-stall samples and executed instructions per CUDA source line.

usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          ("regex:" + kern) if not kern.startswith("=") else kern[1:], "--launch-skip", skip, "--launch-count", "1"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, lines, total, total_inst = None, [], 0, 0
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] in ("Function Name",) or hdr is None:
            continue
        if r[0] != "" and r[0].isdigit():
            try:
                s = int(r[4]); inst = int(r[7])
            except ValueError:
                continue
            lines.append((s, inst, cur_file, int(r[0]), r[1].strip()))
            total += s; total_inst += inst
    lines.sort(reverse=True)
    print("total samples %d, warp-instructions %d" % (total, total_inst))
    for s, inst, f, ln, src in lines[:top]:
        print("%6d %5.1f%% inst %9d  %s:%d  %s" % (s, 100.0 * s / max(total, 1), inst, f, ln, src[:110]))


if __name__ == "__main__":
    main()
