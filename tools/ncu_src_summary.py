"""Summarise an `ncu --page source --csv` dump: instructions by opcode, stall reasons, hot spots.
usage: python tools/ncu_src_summary.py src.csv [n_hot]"""
import csv, sys, collections
path = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# a dump may hold several sections (one per launch, each starting with a 'Kernel Name' row): keep the first
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
if len(ends) > 1:
    rows = rows[:ends[1]]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot_inst = sum(int(r[col['Instructions Executed']]) for r in data)
tot_samp = sum(int(r[col['# Samples']]) for r in data)
print('SASS lines %d  warp-instructions %d  samples %d' % (len(data), tot_inst, tot_samp))
byop = collections.Counter(); sampop = collections.Counter()
for r in data:
    src = r[col['Source']].strip()
    toks = src.split()
    op = toks[0] if not toks[0].startswith('@') else toks[1]
    op = op.split('.')[0]
    byop[op] += int(r[col['Instructions Executed']]); sampop[op] += int(r[col['# Samples']])
print('--- by opcode (warp-instr %, samples %)')
for op, c in byop.most_common(22):
    print('%-10s %6.2f%%  %6.2f%%' % (op, 100.0 * c / tot_inst, 100.0 * sampop[op] / tot_samp))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('--- stall reasons (all samples)')
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += int(r[col[s]])
for s, c in tot.most_common():
    if c: print('%-24s %6.2f%%' % (s, 100.0 * c / tot_samp))
print('--- hottest SASS lines by samples')
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col['# Samples']]))[:nhot]
for i in sorted(idx):
    r = data[i]
    top = sorted(((int(r[col[s]]), s) for s in stalls), reverse=True)[:2]
    print('%5d %6s samp %8s exec  %-60s %s' % (i, r[col['# Samples']], r[col['Instructions Executed']], r[col['Source']].strip()[:60], top))
exc = sum(int(r[col['L1 Wavefronts Shared Excessive']]) for r in data); wf = sum(int(r[col['L1 Wavefronts Shared']]) for r in data)
print('shared wavefronts %d excessive %d' % (wf, exc))
