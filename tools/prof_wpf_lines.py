#!/usr/bin/env python3
"""Per-source-line instruction / stall / shared-memory breakdown of an ncu source-page dump of wpf1920_kernel.
usage: prof_wpf_lines.py <src csv> <nvdisasm -g -c dump> <frames in the launch>"""
import csv, re, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
frames = float(sys.argv[3])
lines = open(sys.argv[2]).read().split('\n')
in_fn = False; cur = ('?', 0, '', 0); seq = []
stack_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for ln in lines:
    if ln.startswith('.text.') or ln.strip().startswith('.section'):
        in_fn = ('wpf1920_kernelILi16ELb1' in ln); continue
    if not in_fn: continue
    m = stack_re.search(ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: seq.append((cur, m.group(2)))
n = min(len(seq), len(data))
by = Counter(); bys = Counter(); byc = Counter(); byw = Counter(); bystall = {}
stallcols = [h for h in hdr if h.startswith('stall_')]
for i in range(n):
    cur, ins = seq[i]
    key = (cur[0], cur[1]) if cur[0] == 'wpf1920.cu' else (cur[0], 0)
    ex = int(data[i][ix['Instructions Executed']]); sm = int(data[i][ix['# Samples']])
    by[key] += ex; bys[key] += sm
    byw[key] += int(data[i][ix['L1 Wavefronts Shared']] or 0); byc[key] += int(data[i][ix['L1 Wavefronts Shared Excessive']] or 0)
    d = bystall.setdefault(key, Counter())
    for h in stallcols: d[h] += int(data[i][ix[h]] or 0)
tot = sum(by.values()); tots = sum(bys.values())
print('sass', len(seq), len(data), 'inst/frame', tot / frames)
for key, v in sorted(by.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if v / tot < 0.004 and bys[key] / tots < 0.004: continue
    top = ', '.join(f"{k[6:]}:{c}" for k, c in bystall[key].most_common(3))
    print(f"{key[0]:22s}:{key[1]:4d} inst/frame {v/frames:7.1f} ({100*v/tot:4.1f}%) samples {100*bys[key]/tots:5.1f}%  smem wf/frame {byw[key]/frames:6.1f} excess {byc[key]/frames:5.1f}  {top}")
