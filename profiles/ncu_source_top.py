#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output: stall mix + hottest SASS lines."""
import csv
import sys


def main(path, n=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != 'Address']
    ix = {h: i for i, h in enumerate(hdr)}
    iv = lambda r, h: int(float(r[ix[h]] or 0))
    tot = sum(iv(r, '# Samples') for r in data)
    inst = sum(iv(r, 'Instructions Executed') for r in data)
    print('kernel:', rows[0][1][:100] if rows[0] else '?')
    print('total samples', tot, 'warp instructions', inst)
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(iv(r, h) for r in data) for h in reasons}
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f'  {h:28s} {v:8d} {v / max(tot, 1):7.2%}')
    top = sorted(data, key=lambda r: -iv(r, '# Samples'))[:n]
    for r in top:
        why = max(reasons, key=lambda h: iv(r, h))
        print(f"{iv(r, '# Samples'):7d} {iv(r, 'Instructions Executed'):9d}  {r[ix['Source']].strip()[:64]:64s} {why}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
