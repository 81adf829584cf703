#!/usr/bin/env python
"""Attribute an ncu SASS-page export to source lines / phases using nvdisasm line info.
usage: tools/ncu_lines.py <obj-with-kernel.o> <mangled-kernel-substring> <prof.ncu-rep> <source.cu> [lo:hi:name ...]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kern, rep, srcfile = sys.argv[1:5]
    phases = [p.split(":") for p in sys.argv[5:]]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    # OUTER=1: attribute inlined helpers to their call site in the kernel body (nvdisasm prints the inline chain innermost first;
    # the last annotation before an instruction is the outermost frame)
    outer = os.environ.get("OUTER", "0") == "1"
    sass = subprocess.run(["nvdisasm", "--print-line-info-inline" if outer else "--print-line-info", os.path.join(tmp, cubin)],
                          capture_output=True, text=True).stdout.split("\n")
    addr2line, cur, inside = {}, None, False
    base_name = os.path.basename(srcfile)
    for ln in sass:
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith(base_name) else -1
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ai, si, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    samples, instr, stalls, base = collections.Counter(), collections.Counter(), collections.Counter(), None
    line_stalls = collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) < len(hdr):
            continue
        try:
            a = int(r[ai], 16)
        except ValueError:
            continue
        base = a if base is None else base
        l = addr2line.get(a - base)
        samples[l] += int(r[si] or 0)
        instr[l] += int(r[ii] or 0)
        for c in stall_cols:
            if r[c] and r[c] != "0":
                stalls[hdr[c]] += int(r[c])
                line_stalls[l][hdr[c][6:]] += int(r[c])
    tot, toti = sum(samples.values()), sum(instr.values())
    print(f"samples {tot}  warp-instructions {toti}")
    print("stall mix:", ", ".join(f"{k[6:]} {100 * v / sum(stalls.values()):.1f}%" for k, v in stalls.most_common(8)))
    for lo, hi, name in phases:
        lo, hi = int(lo), int(hi)
        s = sum(v for l, v in samples.items() if l and lo <= l <= hi)
        i = sum(v for l, v in instr.items() if l and lo <= l <= hi)
        print(f"  {name:28s} time {100 * s / tot:5.1f}%   instructions {100 * i / toti:5.1f}%")
    src = open(srcfile).read().split("\n")
    print("hottest lines:")
    by_instr = os.environ.get("SORT", "time") == "instr"
    for l, v in (instr if by_instr else samples).most_common(int(os.environ.get("NLINES", "22"))):
        v = samples[l]
        text = src[l - 1].strip()[:100] if l and l > 0 else "<inlined helper / other file>"
        why = ", ".join(f"{k} {100 * c / max(1, sum(line_stalls[l].values())):.0f}%" for k, c in line_stalls[l].most_common(3)) if os.environ.get("WHY") else ""
        print(f"  line {l}: time {100 * v / tot:4.1f}%  instr {100 * instr[l] / toti:4.1f}%  {text}" + (f"   [{why}]" if why else ""))


if __name__ == "__main__":
    main()
