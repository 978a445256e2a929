"""List the hottest SASS instructions (warp-stall samples) per kernel from an `ncu --page source --csv --print-source sass` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for bi, b in enumerate(blocks):
    if only is not None and bi != only:
        continue
    hdr, data = b['hdr'], b['data']
    iS, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iS]) for r in data)
    print('== kernel %d %s: total samples %d, %d instrs, %d warp instr executed' % (bi, b['name'][:60], tot, len(data), sum(int(r[iex]) for r in data)))
    for k, r in enumerate(data):
        s = int(r[iS])
        if s > tot * thr:
            st = sorted([(int(r[i]), h) for i, h in stall_cols if r[i] not in ('', '0')], reverse=True)[:3]
            print(k, r[isrc].strip()[:64].ljust(64), s, r[iex], st)
