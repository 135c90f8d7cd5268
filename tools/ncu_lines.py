"""Hottest CUDA source lines of every kernel in an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [top=30]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = hdr = None
agg = {}
num = lambda x: float(x) if x not in ("", "-") else 0.0
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": cur_fn = r[1]; continue
    if r and r[0] == "Line No": hdr = r; isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > 10 and r[0] != "":
        key = (cur_fn, cur_file, int(r[0]))
        old = agg.get(key, (0.0, 0.0, ""))
        agg[key] = (old[0] + num(r[isamp]), old[1] + num(r[iex]), r[1].strip()[:100])
for fn in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot = sum(v[0] for k, v in items) or 1; totx = sum(v[1] for k, v in items) or 1
    print(fn, "samples %d" % tot)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print("  %-16s %4d smp %5.2f%% ex %5.2f%% | %s" % (k[1], k[2], 100 * v[0] / tot, 100 * v[1] / totx, v[2]))
