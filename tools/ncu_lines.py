#!/usr/bin/env python
"""Per-source-line totals of an ncu --set full capture: joins the SASS page of one captured launch with the line table of the
matching cubin (nvdisasm -g), so that the share of warp instructions and the lanes per instruction can be read per phase.
usage: tools/ncu_lines.py prof.ncu-rep traverse.sm_100a.cubin LAUNCH_INDEX trace_kernelILi0ELb0ELb0E   (a substring of the mangled name;
       cubins: cuobjdump -xelf all imbatracer_b200/libimba_b200.so)"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, launch, kname = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seg = rows[starts[launch]:(starts[launch + 1] if launch + 1 < len(starts) else len(rows))]
hdr = seg[1]; ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
sass = [(r[si].strip(), int(r[ci]), int(r[ti])) for r in seg[2:] if len(r) > ti]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.splitlines()
# find the function section of the kernel
mangled = None
lines = []  # (line number, opcode text) in order for the kernel
cur_line, in_fn = None, False
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        in_fn = kname in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File ".*/([^/"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        lines.append((cur_line, m.group(1).strip()))
assert len(lines) == len(sass), (len(lines), len(sass))
tot = collections.Counter(); thr = collections.Counter()
for (ln, _), (_, c, t) in zip(lines, sass):
    tot[ln] += c; thr[ln] += t
S = sum(tot.values())
print(f"total warp instructions {S}, lanes/instruction {sum(thr.values()) / S:.2f}")
for ln, c in sorted(tot.items(), key=lambda kv: (kv[0][0], kv[0][1]) if kv[0] else ("", 0)):
    if c * 1000 >= S:
        print(f"{ln[0] if ln else '?'}:{ln[1] if ln else 0:4d}  {100.0 * c / S:5.1f}%  lanes {thr[ln] / max(c, 1):5.1f}")
