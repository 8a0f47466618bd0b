"""Summarise an .ncu-rep (ncu --set full) into one markdown table per kernel class:
duration, DRAM bytes read+written per launch, achieved DRAM GB/s and % of peak, tensor-pipe %, registers.

    python tools/ncu_summarize.py gpurun_out/prof_layer.ncu-rep > profiles/r01_ncu_layer.md
"""
import csv
import io
import re
import subprocess
import sys

COLS = {
    "name": "Kernel Name",
    "ms": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    rep = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6571.0
    if rep.endswith(".csv"):  # already exported: ncu -i X.ncu-rep --page raw --csv > X.csv
        out = "".join(ln for ln in open(rep) if ln.startswith('"'))
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def col(key):
        want = COLS[key]
        for i, h in enumerate(hdr):
            if h == want or h.endswith("." + want) or h.endswith(want):
                return i
        return None

    idx = {k: col(k) for k in COLS}
    agg = {}
    for r in rows[2:]:
        def val(k, default=0.0):
            i = idx[k]
            if i is None or r[i] in ("", "n/a"):
                return default
            v = float(r[i].replace(",", ""))
            return v * SCALE.get(units[i], 1.0)

        name = r[idx["name"]]
        name = re.sub(r"\(anonymous namespace\)::|unnamed>::|void |pi05::", "", name)
        name = re.sub(r"\(.*", "", name)[:64]
        key = (name, int(val("grid")), int(val("block")))
        a = agg.setdefault(key, {"n": 0, "ms": 0.0, "bytes": 0.0, "dram": 0.0, "tensor": 0.0, "sm": 0.0,
                                 "regs": int(val("regs")), "occ": 0.0})
        a["n"] += 1
        a["ms"] += val("ms")
        a["bytes"] += val("rd") + val("wr")
        a["dram"] += val("dram_pct")
        a["tensor"] += val("tensor_pct")
        a["sm"] += val("sm_pct")
        a["occ"] += val("occ")
    print(f"Source: `{rep}` (ncu --set full --clock-control none; per-launch averages; cold-cache, serialised launches).")
    print(f"GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; % of the measured {peak:.0f} GB/s "
          "copy peak (MEASURED_PEAKS.json).\n")
    print("| kernel | grid x block | launches | avg ms | DRAM MB/launch | DRAM GB/s | % of measured HBM peak | "
          "ncu dram % | tensor pipe % | SM % | warps active % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for (name, grid, block), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        n = a["n"]
        ms = a["ms"] / n
        by = a["bytes"] / n
        gbs = by / 1e9 / (ms / 1e3) if ms > 0 else 0.0
        print(f"| `{name}` | {grid} x {block} | {n} | {ms:.4f} | {by / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} | "
              f"{a['dram'] / n:.1f} | {a['tensor'] / n:.1f} | {a['sm'] / n:.1f} | {a['occ'] / n:.1f} | {a['regs']} |")


if __name__ == "__main__":
    main()
