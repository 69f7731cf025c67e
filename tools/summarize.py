#!/usr/bin/env python3
"""Prints one line per bench JSON file given on the command line."""
import json
import sys

for path in sys.argv[1:]:
    try:
        for l in open(path):
            l = l.strip()
            if not l.startswith("{"):
                continue
            j = json.loads(l)
            r = j.get("roofline", {})
            print(f"{path.split('/')[-1]:36s} {j['config']['workload']:52s} qps {j['value']:10.0f} e2e {j['e2e']['value']:10.0f} "
                  f"ms {j['ms_per_step']:8.3f} {r.get('kernel')} scan_ms {r.get('kernel_ms_per_step', 0):8.3f} "
                  f"GB/s {r.get('achieved_gbs', 0):7.0f} TOP/s {r.get('achieved_tops', 0):7.0f} frac {r.get('frac') or 0:.3f} "
                  f"launches {r.get('launches_per_step')} rescans {j.get('overflow_rescans')} parity {j.get('parity')} clk {j.get('clocks', {}).get('sm_mhz') if j.get('clocks') else None}")
    except Exception as e:
        print(path, "ERR", e)
