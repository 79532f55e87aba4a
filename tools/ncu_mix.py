"""Dynamic instruction mix / hottest SASS lines from `ncu --page source --csv` output."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
body = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body)
print('total warp-inst', tot)
c = Counter(); st = Counter()
for r in body:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    c[op.split('.')[0]] += int(r[ie]); st[op.split('.')[0]] += int(r[iss] or 0)
stt = sum(st.values())
for k, v in c.most_common(28):
    print(f'{k:10s} {v:11d} {100*v/tot:5.1f}%   stall-samples {100*st[k]/max(stt,1):5.1f}%')
if len(sys.argv) > 2:
    print('--- hottest lines by stall samples')
    for r in sorted(body, key=lambda r: -int(r[iss] or 0))[:int(sys.argv[2])]:
        print(r[iss], r[ie], r[isrc][:100])
