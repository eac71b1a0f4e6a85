"""per-function breakdown (samples, instructions, top stall reasons) of the fused kernel from an .ncu-rep (needs -lineinfo)"""
import csv, io, subprocess, collections, re, sys, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; shdr = None; lines = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == 'Line No': shdr = r; continue
    if shdr and len(r) > 8 and r[0].isdigit(): lines.append((cur, r))
si = shdr.index('# Samples'); ii = shdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(shdr) if h.startswith('stall_') and 'Not Issued' not in h]
def I(x):
    try: return int(x)
    except Exception: return 0
srcl = open(os.path.join(root, 'plugin_navierstokes_b200/csrc/ns_fused.cuh')).read().split('\n')
marks = []
for i, l in enumerate(srcl, 1):
    m = re.search(r'(fused_\w+|fv1_fused_kernel)\s*\(', l)
    if m and (l.startswith('NSB_HD') or l.startswith('template') or l.startswith('__global__') or (i >= 2 and srcl[i - 2].startswith('template') and not l.startswith(' '))):
        marks.append((i, m.group(1)))
def func_of(line):
    f = 'top'
    for ln, name in marks:
        if line >= ln: f = name
    return f
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for f, r in lines:
    key = func_of(int(r[0])) if f == 'ns_fused.cuh' else f
    agg[key][0] += I(r[si]); agg[key][1] += I(r[ii])
    for c in stall_cols: agg[key][2][shdr[c][6:]] += I(r[c])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print('total warp instructions %.3f G' % (ti / 1e9))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 200 < ts and v[1] * 200 < ti: continue
    print(f"{k:26s} samples {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}% ({v[1]/1e9:.3f} G) ", ' '.join(f"{a}:{100*b/max(1,v[0]):.0f}%" for a, b in v[2].most_common(4)))
