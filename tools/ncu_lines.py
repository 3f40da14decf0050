#!/usr/bin/env python
"""Aggregate an ncu report's source page by CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != '-': continue
    try: ln = int(r[0])
    except ValueError: continue
    g = lambda name: int(r[hdr.index(name)])
    a = agg.setdefault((cur, ln), [0, 0, 0, r[1], {}])
    a[0] += g('# Samples'); a[1] += g('Instructions Executed'); a[2] += g('Thread Instructions Executed')
    for nm in ('stall_barrier','stall_long_sb','stall_short_sb','stall_wait','stall_math','stall_mio','stall_lg','stall_branch_resolving','stall_no_inst','stall_not_selected','stall_selected','stall_dispatch'):
        if nm in hdr: a[4][nm] = a[4].get(nm, 0) + g(nm)
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print('total samples', ts, 'warp-instructions', ti)
byfile = {}
for (f, ln), a in agg.items():
    b = byfile.setdefault(f, [0, 0, 0]); b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for f, b in byfile.items(): print('  %-24s samples %5.1f%%  inst %5.1f%%  thr/inst %4.1f' % (f, 100*b[0]/ts, 100*b[1]/ti, b[2]/max(b[1],1)))
stall = {}
for a in agg.values():
    for k, v in a[4].items(): stall[k] = stall.get(k, 0) + v
print('  stalls:', ', '.join('%s %.1f%%' % (k[6:], 100*v/ts) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = max(a[4].items(), key=lambda kv: kv[1])[0][6:] if a[4] else ''
    print('%-20s %4d samp %5.1f%% inst %5.1f%% thr %4.1f %-10s %s' % (f[:20], ln, 100*a[0]/ts, 100*a[1]/ti, a[2]/max(a[1],1), top, a[3].strip()[:80]))
