"""Print the hottest SASS instructions and the stall mix per region (before / inside / after the DMMA loop) of an ncu source-page CSV."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:100])
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
keys = ['stall_long_sb', 'stall_math', 'stall_wait', 'stall_short_sb', 'stall_barrier', 'stall_branch_resolving', 'stall_not_selected', 'stall_lg', 'stall_mio', 'stall_selected']
def op(s):
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s); return m.group(2) if m else s
top = sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in top:
    print(data.index(r), r[ix['Source']][:60].ljust(60), '%5.1f%%' % (100 * int(r[ix['# Samples']]) / tot), r[ix['Instructions Executed']].rjust(10),
          ' '.join('%s=%s' % (k[6:], r[ix[k]]) for k in keys[:-1] if r[ix[k]] not in ('0', '')))
first = next(i for i, r in enumerate(data) if op(r[ix['Source']]).startswith('DMMA'))
last = max(i for i, r in enumerate(data) if op(r[ix['Source']]).startswith('DMMA'))
for name, (a, b) in {'pre': (0, first - 80), 'loop': (first - 80, last + 10), 'post': (last + 10, len(data))}.items():
    share = 100 * sum(int(r[ix['# Samples']] or 0) for r in data[a:b]) / tot
    print('%-5s %5.1f%%' % (name, share), {k[6:]: '%.1f' % (100 * sum(int(r[ix[k]] or 0) for r in data[a:b]) / tot) for k in keys})
