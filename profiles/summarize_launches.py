#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table.

  python profiles/summarize_launches.py gpurun_out/r1_launches.csv profiles/r1_launches_step [--step 1 | --all]

Picks one full DDIM step (the launches between two consecutive cfg_ddim_step kernels), writes
<out>.csv.gz (the raw rows of that step) and <out>.md (time share per kernel, per template instance)."""
import csv
import gzip
import re
import sys
from collections import defaultdict


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    n = re.sub(r"\((CUtensorMap_st|const|T1|float|int|long|double|unsigned|__nv).*$", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.strip()


def main():
    src, out = sys.argv[1], sys.argv[2]
    step = int(sys.argv[sys.argv.index("--step") + 1]) if "--step" in sys.argv else 1
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    rows = [r for r in rd]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    marks = [i for i, r in enumerate(rows) if "ddim" in r[ki]]
    if "--all" in sys.argv:          # list captured with `bench.py --ncu-step`: exactly one step between profiler start/stop
        lo, hi = 0, len(rows)
    else:
        if len(marks) <= step:
            raise SystemExit(f"only {len(marks)} step marker(s) in the list")
        lo, hi = marks[step - 1] + 1, marks[step] + 1
    sel = rows[lo:hi]
    with gzip.open(out + ".csv.gz", "wt", newline="") as g:
        w = csv.writer(g)
        w.writerow(hdr)
        w.writerows(sel)
    agg = defaultdict(lambda: [0, 0.0])
    fam = defaultdict(lambda: [0, 0.0])
    for r in sel:
        n = short(r[ki])
        ns = float(r[vi])
        agg[n][0] += 1
        agg[n][1] += ns
        f = re.sub(r"<.*", "", n)
        fam[f][0] += 1
        fam[f][1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(out + ".md", "w") as m:
        m.write(f"# ncu launch list, one DDIM step of `bench.py` (launches {lo}..{hi - 1} of {src})\n\n")
        m.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised;\n"
                "compare SHARES with bench.py's event-timed table, not absolutes.\n\n")
        m.write(f"{len(sel)} launches, {tot / 1e6:.1f} ms summed kernel time.\n\n## by kernel family\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            m.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% |\n")
        m.write("\n## by template instance\n\n| kernel | launches | ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            m.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
