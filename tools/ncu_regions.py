"""Region view of every kernel in an .ncu-rep captured with --import-source on: executed
instructions, stall samples and opcode mix per block of SASS instructions.
usage: python tools/ncu_regions.py report.ncu-rep [block=200]"""
import collections, csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = set()
for si, st in enumerate(starts):
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    name = rows[st][1]
    if name in seen:          # ncu prints every kernel once per result id and once per launch
        continue
    seen.add(name)
    hdr, data = rows[st + 1], [r for r in rows[st + 2:end] if len(r) > 10]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot_ex = sum(float(r[iex]) for r in data)
    tot_s = sum(float(r[isamp]) for r in data)
    print(name)
    print("instructions %.3e  samples %d  sass lines %d" % (tot_ex, tot_s, len(data)))
    op = lambda r: re.sub(r"^@!?U?P\d+\s+", "", r[ia]).split()[0].split(".")[0]
    for b in range(0, len(data), B):
        blk = data[b:b + B]
        ex = sum(float(r[iex]) for r in blk); s = sum(float(r[isamp]) for r in blk)
        if ex < 0.004 * tot_ex: continue
        dyn = collections.Counter()
        for r in blk: dyn[op(r)] += float(r[iex])
        stc = collections.Counter()
        for i, h in stalls: stc[h[6:]] += sum(float(r[i] or 0) for r in blk)
        print("%5d ex %5.1f%% smp %5.1f%% | %s | %s" % (
            b, 100 * ex / tot_ex, 100 * s / tot_s,
            " ".join("%s:%.0f%%" % (k, 100 * v / ex) for k, v in dyn.most_common(6)),
            " ".join("%s:%.0f%%" % (k, 100 * v / max(s, 1)) for k, v in stc.most_common(3))))
    print()
