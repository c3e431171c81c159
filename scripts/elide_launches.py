"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list for profiles/: one line per launch (id, kernel, block,
grid, duration_ns); runs of more than 8 launches of the same kernel and grid keep their first 3 and last 3 + a count and the mean.
usage: python scripts/elide_launches.py <ncu csv log> <out csv> "<header comment>" """
import csv, sys

src, dst, header = sys.argv[1:4]
rows = []
with open(src, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], r["Block Size"], r["Grid Size"], int(round(ns))))
out = [f"# {header}",
       "# runs of more than 8 identical kernels (the 512-step pre-rolls, the single-walker cfg1 loop) are elided to their first 3 and last 3 launches + a count and the mean",
       "id,kernel,block,grid,duration_ns"]
i = 0
while i < len(rows):
    j = i
    while j < len(rows) and rows[j][1:4] == rows[i][1:4]:
        j += 1
    run = rows[i:j]
    def fmt(r):
        return f'{r[0]},"{r[1]}","{r[2]}","{r[3]}",{r[4]}'
    if len(run) > 8:
        out += [fmt(r) for r in run[:3]]
        out.append(f"# ... {len(run) - 6} more launches of the same kernel and grid; mean of the run {sum(r[4] for r in run) / len(run):.0f} ns")
        out += [fmt(r) for r in run[-3:]]
    else:
        out += [fmt(r) for r in run]
    i = j
open(dst, "w").write("\n".join(out) + "\n")
print(len(rows), "launches ->", len(out), "lines")
