#!/usr/bin/env python3
"""Per-SASS-instruction view of an `ncu --page source --csv` export: share of executed warp instructions, average active
threads and stall-sample share.  usage: ncu_hot.py source.csv [min_inst_pct]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix['stall_wait'] and r[0].startswith('0x')]
    # several launches of one kernel are concatenated: keep the first
    first = data[0][0]
    for k in range(1, len(data)):
        if data[k][0] == first:
            data = data[:k]
            break
    tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
    tot_s = sum(int(r[ix['# Samples']]) for r in data)
    print('sass instructions', len(data), 'executed', tot_i, 'samples', tot_s)
    for k, r in enumerate(data):
        i = int(r[ix['Instructions Executed']])
        s = int(r[ix['# Samples']])
        if 100 * i / tot_i > thresh or 100 * s / tot_s > 0.5:
            print(k, r[ix['Source']].strip()[:64].ljust(64), 'inst%5.2f' % (100 * i / tot_i), 'thr', r[ix['Avg. Threads Executed']].rjust(3),
                  'smp%5.2f' % (100 * s / tot_s), 'lsb', r[ix['stall_long_sb']], 'ssb', r[ix['stall_short_sb']])


if __name__ == '__main__':
    main()
