#!/usr/bin/env python
"""Attribute executed warp instructions and stall samples of one kernel in an .ncu-rep to source lines, by
joining the SASS page (per-instruction counters) with nvdisasm's line info for the SAME library build.

  python tools/ncu_by_line.py <rep> <library.so> <kernel name substring as ncu prints it> <mangled substring> [launch index]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_map(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    m, cur, on = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("//-") and ".text." in ln:
            on = kernel in ln
        if not on:
            continue
        f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if a:
            m[int(a.group(1), 16)] = cur
    return m


def main():
    rep, lib, kernel, mangled = sys.argv[1:5]
    launch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(raw.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    blocks = [b for b in blocks if kernel in b["name"]]
    b = blocks[launch]
    hdr = b["rows"][0]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    rows = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    base = min(int(r[ia], 16) for r in rows)
    lm = line_map(lib, mangled)
    by = defaultdict(lambda: [0, 0])
    for r in rows:
        key = lm.get(int(r[ia], 16) - base, ("?", 0))
        by[key][0] += int(r[iex])
        by[key][1] += int(r[isamp])
    tot_e = sum(v[0] for v in by.values())
    tot_s = sum(v[1] for v in by.values())
    byfile = defaultdict(lambda: [0, 0])
    for (f, l), v in by.items():
        byfile[f][0] += v[0]
        byfile[f][1] += v[1]
    print(f"kernel {b['name'][:70]}: {tot_e} warp instructions, {tot_s} samples")
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:22s} exec {100 * v[0] / tot_e:5.1f}%  samples {100 * v[1] / tot_s:5.1f}%")
    print("top lines:")
    for (f, l), v in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {f}:{l:<5d} exec {100 * v[0] / tot_e:5.2f}%  samples {100 * v[1] / tot_s:5.2f}%")


if __name__ == "__main__":
    main()
