#!/usr/bin/env python
"""Key metrics per kernel from an `ncu --set full` report:  python profiles/summarize_ncu.py <rep | raw.csv.gz> <out.md> [names...]

<raw.csv.gz> = `ncu -i <rep> --page raw --csv | gzip` made on the GPU box (reports above ~60 MB do not travel back).
A name may end in `*2` / `*3`: the operator launches that many kernels (two-kernel GroupNorm), all labelled with it."""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("dur us", "gpu__time_duration.sum", 1e-3),
    ("dram rd MB", "dram__bytes_read.sum", None),
    ("dram wr MB", "dram__bytes_write.sum", None),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("tensor-mem %", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("smem KB", "launch__shared_mem_per_block_dynamic", None),
    ("grid", "launch__grid_size", 1),
]


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(v) * f


def main():
    rep, out = sys.argv[1], sys.argv[2]
    labels = []
    for name in sys.argv[3:]:
        base, _, rep_n = name.partition("*")
        labels += [base] * (int(rep_n) if rep_n else 1)
    if rep.endswith(".gz"):
        import gzip
        raw = gzip.open(rep, "rt").read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["| # | kernel | " + " | ".join(m[0] for m in METRICS) + " |", "|---|---|" + "---:|" * len(METRICS)]
    data = [r for r in data if "at::" not in r[col["Kernel Name"]]]     # drop torch's L2-flush fills between the operators
    for k, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        if k < len(labels):
            name = f"{labels[k]}: {name}"
        cells = []
        for label, key, scale in METRICS:
            if key not in col or r[col[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v, u = r[col[key]].replace(",", ""), units[col[key]]
            if scale is None:
                b = to_bytes(v, u)
                cells.append(f"{b / 1e6:.1f}" if "MB" in label else f"{b / 1e3:.1f}")
            else:
                x = float(v) * scale
                if label == "dur us" and u != "ns":
                    x = float(v) * {"us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
                cells.append(f"{x:.1f}" if x < 1000 else f"{x:.0f}")
        lines.append(f"| {k} | `{name}` | " + " | ".join(cells) + " |")
    txt = "\n".join(lines) + "\n"
    open(out, "w").write(f"# ncu --set full summary of {rep}\n\n" + txt)
    print(txt)


if __name__ == "__main__":
    main()
