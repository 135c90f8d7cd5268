#!/usr/bin/env python
"""Static resource table of every kernel in cpp_ls_lib.so from the `-Xptxas -v` logs the build
leaves in movie_recommender_b200/csrc/build/*.ptxas.log: registers, spill bytes, static shared
memory, per translation unit.  No GPU needed.

    python tools/ptxas_table.py > profiles/ptxas_r02.txt
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "movie_recommender_b200", "csrc", "build")


def short(name):
    """Demangled kernel name without the parameter list and the anonymous-namespace noise."""
    name = name.replace("(anonymous namespace)::", "").replace("mrb::", "")
    name = re.sub(r"^void ", "", name)
    depth = 0
    for i, ch in enumerate(name):           # cut at the '(' that opens the parameter list
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    logs = sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log")))
    if not logs:
        sys.exit("no ptxas logs under %s: run __graft_entry__.build() first" % BUILD)
    rows = []
    for path in logs:
        unit = os.path.basename(path).replace(".ptxas.log", ".cu")
        cur, spill = None, (0, 0, 0)
        for line in open(path):
            m = re.search(r"Compiling entry function '([^']+)' for 'sm_100a'", line)
            if m:
                cur, spill = m.group(1), (0, 0, 0)
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                spill = tuple(int(x) for x in m.groups())
                continue
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(.*)", line)
            if m and cur:
                smem = re.search(r"(\d+) bytes smem", m.group(3))
                rows.append([unit, cur, int(m.group(1)), int(m.group(2) or 0), spill,
                             int(smem.group(1)) if smem else 0])
                cur = None
    names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True,
                           text=True).stdout.splitlines()
    print("# -Xptxas -v summary of movie_recommender_b200/cpp_ls_lib.so (sm_100a), %d kernels" % len(rows))
    print("# unit | kernel | registers | barriers | stack / spill stores / spill loads (bytes) | static smem (bytes)")
    spilled = 0
    for r, d in sorted(zip(rows, names), key=lambda x: (x[0][0], short(x[1]))):
        if r[4][1] or r[4][2]:
            spilled += 1
        print("%-18s %-78s %4d %3d  %4d/%4d/%4d  %6d" % (r[0], short(d)[:78], r[2], r[3],
                                                       r[4][0], r[4][1], r[4][2], r[5]))
    print("# kernels with register spills: %d of %d" % (spilled, len(rows)))


if __name__ == "__main__":
    main()
