#!/usr/bin/env python3
"""Join an `ncu --page source --print-source sass --csv` dump with nvdisasm line info and aggregate
executed instructions / stall samples per CUDA source line (and per file).

usage: prof_by_line.py <src_sass.csv> <nvdisasm -g -c output> <kernel substring> [top]
"""
import csv, re, sys
from collections import Counter, defaultdict

src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')

# parse nvdisasm: find the function section, then sequence of (line info, instruction)
lines = open(dis).read().split('\n')
in_fn = False; cur = ('?', 0); seq = []
stack_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for ln in lines:
    if ln.startswith('.text.') or ln.strip().startswith('.section'):
        in_fn = (kname in ln)
        continue
    if not in_fn:
        continue
    m = stack_re.search(ln)
    if m:
        f = m.group(1).split('/')[-1]
        cur = (f, int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        seq.append((cur, m.group(2)))
print('sass instr in dis:', len(seq), ' in ncu:', len(data))
n = min(len(seq), len(data))
by_line = Counter(); by_line_s = Counter(); by_file = Counter()
for i in range(n):
    cur, _ = seq[i]
    ex = int(data[i][ie]); sm = int(data[i][isamp])
    # attribute inlined codelet code to the outermost frontend.cu line if available
    key = (cur[0], cur[1]) if cur[0] != 'codelets.h' or not cur[2] else (cur[2], cur[3])
    by_line[key] += ex; by_line_s[key] += sm; by_file[cur[0]] += ex
tot = sum(by_line.values())
print('by file:', {k: round(100 * v / tot, 1) for k, v in by_file.items()})
for key, v in by_line.most_common(top):
    print(f"{key[0]:16s}:{key[1]:5d}  {100*v/tot:5.1f}%  samples {by_line_s[key]}")

# ---- opcode mix per source-line range (frontend.cu outermost line) ----
if len(sys.argv) > 5:
    ranges = [tuple(map(int, r.split('-'))) for r in sys.argv[5].split(',')]
    for lo, hi in ranges:
        c = Counter(); t = 0
        for i in range(n):
            cur, ins = seq[i]
            key = (cur[0], cur[1]) if cur[0] != 'codelets.h' or not cur[2] else (cur[2], cur[3])
            if key[0] == 'frontend.cu' and lo <= key[1] <= hi:
                tk = ins.split()
                op = tk[1] if tk[0].startswith('@') else tk[0]
                c[op.split('.')[0]] += int(data[i][ie]); t += int(data[i][ie])
        print(f"lines {lo}-{hi}: {100*t/tot:5.1f}% of all; per tile {t/24000:.0f}: " + ', '.join(f"{k}:{v/24000:.0f}" for k, v in c.most_common(14)))
    # time share (warp-stall samples) and stall reasons per range, codelets separately
    stall_idx = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_')]
    tot_s = sum(int(r[isamp]) for r in data[:n])
    def region(i):
        cur, _ = seq[i]
        if cur[0] == 'codelets.h':
            return 'codelets(' + str(cur[3]) + ')'
        if cur[0] == 'mel_baked.h':
            return 'mel_baked'
        key = (cur[0], cur[1])
        if key[0] == 'frontend.cu':
            for lo, hi in ranges:
                if lo <= key[1] <= hi:
                    return f'{lo}-{hi}'
        return 'other'
    agg = defaultdict(lambda: Counter())
    for i in range(n):
        rg = region(i)
        agg[rg]['samples'] += int(data[i][isamp]); agg[rg]['inst'] += int(data[i][ie])
        for j, h in stall_idx:
            try: agg[rg][h] += int(data[i][j])
            except ValueError: pass
    print('--- time share by region (samples), issue efficiency proxy = inst/sample, top stall reasons')
    for rg, c in sorted(agg.items(), key=lambda kv: -kv[1]['samples']):
        st = sorted(((v, k) for k, v in c.items() if k.startswith('stall_')), reverse=True)[:4]
        print(f"{rg:16s} time {100*c['samples']/tot_s:5.1f}%  inst {100*c['inst']/tot:5.1f}%  " + ', '.join(f"{k[6:]}:{100*v/max(c['samples'],1):.0f}%" for v, k in st))
