"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def main(path, note=''):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('oake::<unnamed>::', '')
        us = float(r[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}[r[iu]]
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    print(f'# {note}')
    print('# (gpu__time_duration.sum, --clock-control none; cold-cache serialised launches: compare SHARES)')
    print(f'# total {total / 1e3:.1f} ms over {sum(cnt.values())} launches')
    print('kernel,launches,total_us,share')
    for k, v in tot.most_common():
        print(f'{k},{cnt[k]},{v:.1f},{v / total:.4f}')


if __name__ == '__main__':
    main(sys.argv[1], ' '.join(sys.argv[2:]))
