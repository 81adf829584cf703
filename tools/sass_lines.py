#!/usr/bin/env python
"""Static SASS of a source-line range of k_substeps: opcode mix and instructions per line (nvdisasm line info).
usage: tools/sass_lines.py <lo> <hi> [obj]      (SRC=policy_tc.cu KERNEL=k_policy_tail for another kernel)"""
import subprocess, re, os, tempfile, collections, sys
lo, hi = int(sys.argv[1]), int(sys.argv[2])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SRC", "physics.cu"); KERNEL = os.environ.get("KERNEL", "k_substeps")
obj = sys.argv[3] if len(sys.argv) > 3 else os.path.join(root, 'multiagent-quadruped-environment_b200/csrc', SRC.replace(".cu", ".o"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(root, 'multiagent-quadruped-environment_b200/csrc', SRC)).read().split('\n')
inside = False; cur = None; curfile = None
ops = collections.Counter(); lines = collections.Counter()
for ln in sass:
    if ln.startswith(".text."):
        inside = KERNEL in ln; continue
    if not inside: continue
    m = re.search(r'//## File "(.*)", line (\d+)(.*)', ln)
    if m:
        curfile = os.path.basename(m.group(1)); cur = int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and curfile == SRC and lo <= cur <= hi:
        t = m.group(2).split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += 1; lines[cur] += 1
print(sum(ops.values()), "instructions attributed directly to lines", lo, "-", hi)
print(ops.most_common(40))
for l, c in sorted(lines.items()):
    print(l, c, src[l - 1].strip()[:110])
