#!/usr/bin/env python
"""Aggregate an ncu report per CUDA source line: python scripts/ncu_lines.py rep.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if hdr and r[0] != '':
        d = dict(zip(hdr, r))
        try:
            key = (cur, int(r[0]), r[1].strip()[:90])
            a = agg.setdefault(key, [0, 0, {}])
            a[0] += int(d['Instructions Executed']); a[1] += int(d['Warp Stall Sampling (All Samples)'])
            for k, v in d.items():
                if k.startswith('stall_') and 'Not Issued' not in k:
                    try: a[2][k] = a[2].get(k, 0) + int(v)
                    except Exception: pass
        except Exception: pass
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print("total warp-inst", tot, "stall samples", tots)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    t3 = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{a[1]/tots*100:5.1f}% stall {a[0]/tot*100:5.1f}% inst {key[0]}:{key[1]} {key[2][:64]} | {[(k[6:],v) for k,v in t3]}")
