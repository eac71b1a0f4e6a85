"""aggregates the ncu source page (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > f.csv): stall reasons and the
dynamic opcode mix of one kernel"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = execd = 0; st = {}; opc = {}
for r in data:
    s = int(r[ix["# Samples"]]); tot += s
    e = int(r[ix["Instructions Executed"]]); execd += e
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    d = opc.setdefault(op, [0, 0]); d[0] += e; d[1] += s
    for n in names:
        st[n] = st.get(n, 0) + int(r[ix[n]])
print("samples", tot, "warp instructions executed", execd)
for n, v in sorted(st.items(), key=lambda x: -x[1])[:12]:
    print("  %-24s %5.1f%%" % (n, 100 * v / tot))
for op, (e, s) in sorted(opc.items(), key=lambda x: -x[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  %-8s exec %5.1f%%  samples %5.1f%%" % (op, 100 * e / execd, 100 * s / tot))
