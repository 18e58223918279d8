"""Per-kernel table from an `ncu -i X.ncu-rep --page raw --csv` dump (one row per profiled launch).
usage: python tools/summarize_ncu_raw.py raw.csv [labels.txt] > profiles/<name>.md
HBM peak for the GB/s fraction comes from MEASURED_PEAKS.json (driver-written) when present."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_act": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "l2_pct": "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:   # noqa: BLE001
        return float("nan")


def main(path, labels=None):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: (hdr.index(v) if v in hdr else -1) for k, v in COLS.items()}
    kn, grid, block = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peak.get("hbm_gbs") or 0) or None
    except Exception:   # noqa: BLE001
        peak = None
    lab = [l.strip() for l in open(labels)] if labels else []
    print("| # | op | kernel | grid x block | us | DRAM MB (rd+wr) | GB/s | % of measured HBM peak | tensor pipe % (elapsed / active) | SM % | regs |")
    print("|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|")
    for i, r in enumerate(data):
        def g(k):
            return num(r[ix[k]]) if ix[k] >= 0 else float("nan")
        dur = g("dur")
        du = units[ix["dur"]]
        us = dur / 1e3 if du in ("ns", "nsecond") else (dur if du in ("us", "usecond") else dur * 1e3)
        by = g("rd") + g("wr")
        bu = units[ix["rd"]]
        mb = by / 1e6 if bu == "byte" else (by / 1e3 if bu == "Kbyte" else (by if bu == "Mbyte" else by * 1e3))
        gbs = mb * 1e6 / (us * 1e-6) / 1e9 if us > 0 else float("nan")
        name = r[kn].split("(")[0]
        name = name.replace("mugd::", "").replace("void ", "")
        op = lab[i] if i < len(lab) else ""
        pct = f"{100 * gbs / peak:.1f}" if peak else "n/a"
        print(f"| {i} | {op} | `{name}` | {r[grid]} x {r[block]} | {us:.2f} | {mb:.2f} | {gbs:.0f} | {pct} | "
              f"{g('tensor_pct'):.1f} / {g('tensor_act'):.1f} | {g('sm_pct'):.1f} | {g('regs'):.0f} |")
    if peak:
        print(f"\nHBM peak: {peak:.0f} GB/s (MEASURED_PEAKS.json, driver-measured copy bandwidth); cold-cache, serialised ncu launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
