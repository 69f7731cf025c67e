#!/usr/bin/env python3
"""Regenerates profiles/README.md from the bench JSON lines committed under profiles/."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*.json"))):
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        try:
            j = json.loads(line)
        except Exception:
            continue
        r = j.get("roofline") or {}
        cb = j.get("cpu_baseline") or {}
        rows.append((os.path.basename(path), j, r, cb))

out = ["# profiles/ — measured on B200 (gpurun), round 1", "",
       "Every line below is one `bench.py` JSON line committed in this directory (file name in the first column).",
       "`GB/s` = algorithmic bytes of the scanned image per step / CUDA-event time of the scan kernels;",
       "`frac` = that / 6535 GB/s (measured copy bandwidth, MEASURED_PEAKS.json) or, for tensor-bound lines, TOP/s / peak.",
       "", "| file | workload | GPUs | queries/s | e2e queries/s | ms/step | scan kernel | scan ms | GB/s | TOP/s | bound | frac | CPU baseline q/s (cores) |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, j, r, cb in rows:
    out.append("| {} | {} | {} | {:.0f} | {:.0f} | {:.3f} | {} | {:.3f} | {:.0f} | {:.0f} | {} | {:.3f} | {} |".format(
        name, j["config"]["workload"], j.get("n_gpus"), j["value"], (j.get("e2e") or {}).get("value", 0),
        j["ms_per_step"], (r.get("kernel") or "").split(" (")[0], r.get("kernel_ms_per_step") or 0,
        r.get("achieved_gbs") or 0, r.get("achieved_tops") or 0, r.get("bound"), r.get("frac") or 0,
        "{:.2f} ({})".format(cb["value"], cb["cores"]) if cb else "-"))
out += ["", "Other evidence here: `*.ncu.txt` (ncu summaries: key raw metrics + hottest SASS lines), "
        "`*launches*.csv` (ncu launch lists), `r01_multi_parity_n2.log` (2-GPU sharded search == oracle), "
        "`tma_stream.txt` (TMA streaming microbenchmark: 7.3 TB/s read ceiling with 6 x 16 KB stages)."]
open(os.path.join(ROOT, "profiles", "README.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
