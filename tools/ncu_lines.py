#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an .ncu-rep captured with
--import-source on (library built with -lineinfo). Usage: tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = "?"; hdr = None; lines = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or r[0] == "": continue
        try:
            ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ii]), int(r[si])))
        except (ValueError, IndexError):
            pass
    ti = sum(l[3] for l in lines) or 1; ts = sum(l[4] for l in lines) or 1
    print("total warp instructions %d, samples %d" % (ti, ts))
    print("--- by instructions")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * l[3] / ti, 100.0 * l[4] / ts, l[0], l[1], l[2][:110]))
    print("--- by stall samples")
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * l[4] / ts, 100.0 * l[3] / ti, l[0], l[1], l[2][:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
