"""summarise an .ncu-rep: key metrics, instruction mix, stall reasons, hottest source lines"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor']
for h, v in zip(hdr, vals):
    if h in want: print(f"{h:70s} {v}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); stall = collections.Counter()
for r in rows[2:]:
    try: n = int(r[ix['Instructions Executed']])
    except Exception: continue
    tot += n
    toks = r[ix['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    byop[op.split('.')[0]] += n
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h: stall[h] += int(r[ix[h]])
print('total warp instr', tot)
print('  '.join(f'{k}:{100*v/tot:.1f}%' for k, v in byop.most_common(16)))
st = sum(stall.values())
print('  '.join(f'{k[6:]}:{100*v/st:.0f}%' for k, v in stall.most_common(8)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; lines = []; shdr = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == 'Line No': shdr = r; continue
    if shdr and len(r) > 8 and r[0].isdigit(): lines.append((cur, r))
si = shdr.index('# Samples'); ii = shdr.index('Instructions Executed')
def I(x):
    try: return int(x)
    except Exception: return 0
tots = sum(I(t[1][si]) for t in lines); toti = sum(I(t[1][ii]) for t in lines)
byfile = collections.Counter()
for f, r in lines: byfile[f] += I(r[ii])
print('instr by file:', {k: f"{100*v/toti:.1f}%" for k, v in byfile.most_common()})
top = sorted(lines, key=lambda t: -I(t[1][ii]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for f, r in top:
    print(f"{100*I(r[ii])/toti:5.1f}% inst {100*I(r[si])/tots:5.1f}% samp  {f}:{r[0]}: {r[1].strip()[:100]}")
