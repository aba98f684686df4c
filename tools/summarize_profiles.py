"""Turn the ncu reports brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_profiles.py r01b        # prefix of the reports / launch list in gpurun_out/
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main(prefix):
    OUT.mkdir(exist_ok=True)
    for rep in sorted((ROOT / "gpurun_out").glob(prefix + "_*.ncu-rep")):
        hdr, units, data = raw_rows(rep)
        ik = hdr.index("Kernel Name")
        cols = [hdr.index(k) for k in KEEP if k in hdr]
        with open(OUT / (rep.stem + "_raw.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["Kernel Name"] + [hdr[c] for c in cols])
            w.writerow(["(unit)"] + [units[c] for c in cols])
            for r in data:
                w.writerow([r[ik]] + [r[c] for c in cols])
        print("wrote", rep.stem + "_raw.csv", len(data), "launches")
    ll = ROOT / "gpurun_out" / (prefix + "_launches_bench.csv")
    if ll.exists():
        rows = list(csv.reader(open(ll)))
        h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr = rows[h]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = defaultdict(list)
        for r in rows[h + 1:]:
            if len(r) == len(hdr):
                agg[r[ik]].append(float(r[iv].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        with open(OUT / (prefix + "_launches_bench_by_kernel.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "mean_ns", "total_ns", "share_of_profiled_time"])
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                w.writerow([k, len(v), round(sum(v) / len(v), 1), round(sum(v), 1), round(sum(v) / tot, 4)])
        (OUT / (prefix + "_launches_bench.csv")).write_text(ll.read_text())
        print("wrote launch lists;", len(agg), "distinct kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01b")
