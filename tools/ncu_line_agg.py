"""per-source-line samples / instructions of one kernel: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--kernel-name regex:K] > f.csv
usage: ncu_line_agg.py f.csv [bucket]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr): continue
    try: ln = int(r[0]); s = int(r[i_s]); e = int(r[i_e])
    except ValueError: continue
    a = agg.setdefault((cur, ln), [0, 0]); a[0] += s; a[1] += e
tot = sum(v[0] for v in agg.values()); tote = sum(v[1] for v in agg.values())
print("samples", tot, "instructions", tote)
b = {}
for (f, l), (s, e) in agg.items():
    k = (f, l // bucket * bucket); a = b.setdefault(k, [0, 0]); a[0] += s; a[1] += e
for k, (s, e) in sorted(b.items()):
    if s / tot > 0.004: print("%-16s %5d  samples %5.1f%%  inst %5.1f%%" % (k[0], k[1], 100 * s / tot, 100 * e / tote))
