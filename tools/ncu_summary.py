#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics per kernel + top stall instructions (needs ncu on PATH)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__inst_executed_pipe_tmem.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_active.avg", "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print(f"   {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")
        for h in hdr:
            if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") or \
               (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                try:
                    v = float(d[h])
                except Exception:
                    continue
                if v > 0.3:
                    print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for n, i in enumerate(idx):
        h = rows[i]
        col = {name: j for j, name in enumerate(h)}
        end = idx[n + 1] - 1 if n + 1 < len(idx) else len(rows)
        data = [r for r in rows[i + 1:end] if len(r) > col["# Samples"] and r[col["# Samples"]].isdigit()]
        tot = sum(int(r[col["# Samples"]]) for r in data) or 1
        print(f"-- kernel {n}: {tot} samples; top instructions")
        for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]:
            print(f"   {100.0 * int(r[col['# Samples']]) / tot:5.1f}%  x{r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
