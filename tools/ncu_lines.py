#!/usr/bin/env python3
"""Hottest SOURCE lines of an .ncu-rep captured with --import-source on (needs ncu on PATH, -lineinfo build):
warp-stall samples summed per (file, line), with the executed-instruction count."""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = "?", None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            s, x = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                key = (cur_file, int(r[0]))
                v = lines.setdefault(key, [0, 0, r[1].strip()])
                v[0] += int(r[s])
                v[1] += int(r[x])
            except (ValueError, IndexError):
                pass
    tot = sum(v[0] for v in lines.values()) or 1
    print(f"{tot} samples")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * v[0] / tot:5.1f}%  x{v[1]:>10d}  {f}:{ln:<5d} {v[2][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
