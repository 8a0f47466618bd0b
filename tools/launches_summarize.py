"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/launches_summarize.py gpurun_out/launches.csv > profiles/r01_launches_summary.md
"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    agg, total, n = {}, 0.0, 0
    with open(path, newline="") as f:
        rows = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(rows)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
                  "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|unnamed>::|void |pi05::", "", r["Kernel Name"])
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
        n += 1
    print(f"Source: `{path}` — {n} launches, {total:.2f} ms of kernel time (ncu serialises launches and runs them "
          "cold-cache: compare SHARES with the CUDA-event numbers, not absolutes).\n")
    print("| kernel | launches | total ms | share % | avg us |")
    print("|---|---|---|---|---|")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"| `{name}` | {cnt} | {ms:.3f} | {100 * ms / total:.2f} | {1e3 * ms / cnt:.1f} |")


if __name__ == "__main__":
    main()
