"""Per-opcode and per-region executed-instruction totals from an ncu source page.
  python tools/ncu_hot.py rep.ncu-rep <launch-skip> [topN]"""
import csv, io, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', skip, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = 0; byop = collections.Counter(); samp = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit(): continue
    src = r[iS].strip(); n = int(r[iE] or 0); s = int(r[iSm] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    op = op.split('.')[0]
    byop[op] += n; samp[op] += s; tot += n
    lines.append((n, s, src))
print('total warp instructions', tot)
for op, n in byop.most_common(topn):
    print(f'{op:12s} {n:14d} {100*n/tot:6.2f}%   samples {samp[op]}')
if '--lines' in sys.argv:
    for i, (n, s, src) in enumerate(lines):
        print(f'{i:5d} {n:12d} {s:7d}  {src}')
