"""Markdown table of the bench lines committed under profiles/ (one row per JSON line, sub-records of `configs` too)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def row(name, d):
    r = d.get("roofline") or {}
    st = d.get("search_stats") or {}
    cpu = d.get("cpu_baseline") or {}
    sus = d.get("sustained") or {}
    e2e = d.get("e2e")
    e2e = e2e.get("value") if isinstance(e2e, dict) else e2e
    wl = d.get("workload") or (d.get("config") or {}).get("workload", "")
    frac = r.get("frac")
    return ("| {} | {} | {} | {:.0f} | {} | {} | {:.3f} | {} | {} | {} | {} | {} | {} |".format(
        name, wl, d.get("n_gpus", 1), d.get("value", 0), f"{e2e:.0f}" if e2e else "-",
        f"{sus['value']:.0f}" if sus.get("value") else "-", d.get("ms_per_step", 0),
        f"{r.get('kernel_ms_per_step', 0):.3f}" if r else "-",
        f"{r.get('achieved', 0):.0f} {r.get('unit', '')}" if r else "-", r.get("bound", "-"),
        f"{frac:.3f}" if frac is not None else "-",
        "{:.0f} / {:.0f}".format(st.get("rescored_rows_per_query", 0), st.get("deferred_rows_per_query", 0)) if st else "-",
        f"{cpu['value']:.2f} ({cpu['cores']})" if cpu else "-"))


def main(prefix):
    print("| file | workload | GPUs | queries/s | e2e q/s | sustained q/s | ms/step | scan ms | achieved | bound | frac of peak | "
          "re-scored rows/query (in-chunk+in-kernel / deferred) | CPU baseline q/s (cores) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", prefix + "*.json"))):
        d = last_json(path)
        if not d or "value" not in d:
            continue
        name = os.path.basename(path)
        print(row(name, d))
        for key, sub in (d.get("configs") or {}).items():
            if "value" in sub:
                print(row(f"{name} :: configs.{key}", sub))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02_")
