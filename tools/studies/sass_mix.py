"""Instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv` output (SASS view)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isrc = hdr.index('Source'); ismp = hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data); tt = sum(int(r[it]) for r in data); ts = sum(int(r[ismp]) for r in data)
print('warp instr', tot, 'thread instr', tt, 'avg lanes', tt / tot, 'samples', ts)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    f = r[isrc].split()
    o = f[1] if f[0].startswith('@') else f[0]
    o = o.split('.')[0]
    op[o] += int(r[ia]); ops[o] += int(r[ismp])
for o, c in op.most_common(28):
    print('%-10s %10d %5.1f%%  samples %5.1f%%' % (o, c, 100 * c / tot, 100 * ops[o] / ts))
