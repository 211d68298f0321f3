"""Summarise an `ncu --page source --csv` export: stall reasons, per-opcode samples and
the hottest instructions of the main loop."""
import csv, sys
def I(x):
    try: return int(float(x))
    except Exception: return 0
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(fn)))
# several kernels may be concatenated: take the first block
hdr = rows[1]
isrc, ist, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == 'Address': break
    data.append(r)
tot = sum(I(r[ist]) for r in data)
print(rows[0][1][:90]); print('samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stall: agg[hdr[i]] = agg.get(hdr[i], 0) + I(r[i])
print(' '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
cat = {}
for r in data:
    t = r[isrc].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    c = cat.setdefault(op, [0, 0]); c[0] += I(r[ist]); c[1] += I(r[iex])
for op, (s_, e_) in sorted(cat.items(), key=lambda kv: -kv[1][0])[:12]:
    print('  %-22s samples %6d (%4.1f%%) executed %9d' % (op, s_, 100.0 * s_ / max(tot, 1), e_))
mx = max(I(r[iex]) for r in data)
print('hottest loop instructions:')
for k, r in enumerate(data):
    if I(r[ist]) * 120 > tot and I(r[iex]) > mx // 3:
        rs = sorted([(I(r[i]), hdr[i][6:]) for i in stall], reverse=True)[:2]
        print('  %5d %5d %-60s %s' % (k, I(r[ist]), r[isrc][:60], [x for x in rs if x[0]]))
