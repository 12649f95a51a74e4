#!/usr/bin/env python
"""Join an `ncu --page source --print-source sass --csv` export with nvdisasm line info and rank source lines by
stall samples / executed instructions.  usage: ncu_lines.py prof_sass.csv dis.txt kernel_substring [topN]"""
import csv, re, sys
sass_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: per function, list of (offset, file, line)
cur_line = None; in_fn = False; lines = {}
for l in open(dis, errors='replace'):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        in_fn = kern in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
out = {}; base = None; k_ok = False
for r in rows:
    if r and r[0] == 'Kernel Name': k_ok = kern in r[1] or True; continue
    if r and r[0] == 'Address': hdr = r; continue
    if not r or not r[0].startswith('0x'): continue
    a = int(r[0], 16)
    if base is None: base = a
    d = dict(zip(hdr, r))
    key = lines.get(a - base, (None, ''))[0]
    s = int(d['# Samples'] or 0); i = int(d['Instructions Executed'] or 0)
    st = {k: int(d[k] or 0) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
    o = out.setdefault(key, [0, 0, {}])
    o[0] += s; o[1] += i
    for k, v in st.items(): o[2][k] = o[2].get(k, 0) + v
ts = sum(o[0] for o in out.values()) or 1; ti = sum(o[1] for o in out.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for key, (s, i, st) in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ','.join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst  {key}  [{tops}]")
