"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export: top source lines by stall samples."""
import collections
import csv
import sys


def main(path, top=16):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = {'file': r[1], 'rows': [], 'hdr': None}
            sections.append(cur)
        elif len(r) == 2 and r[0] == 'Function Name':
            cur['func'] = r[1]
        elif r and r[0] == 'Line No':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    # group sections into launches: a new launch starts when a file name repeats
    launches, seen = [], set()
    for s in sections:
        if s['file'] in seen or not launches:
            launches.append([])
            seen = set()
        seen.add(s['file'])
        launches[-1].append(s)
    for k, secs in enumerate(launches):
        per_line, stall, tot = collections.Counter(), collections.Counter(), 0
        for s in secs:
            hdr = s['hdr']
            i_samp = hdr.index('# Samples')
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
            fname = s['file'].split('/')[-1]
            for r in s['rows']:
                if r[0] == '' or r[2] != '-':
                    continue
                try:
                    n = int(r[i_samp])
                except ValueError:
                    continue
                if n == 0:
                    continue
                key = (fname, int(r[0]), r[1].strip()[:80])
                per_line[key] += n
                tot += n
                for i, h in stall_cols:
                    try:
                        stall[(key, h)] += int(r[i])
                    except ValueError:
                        pass
        print(f'===== launch {k}: {secs[0].get("func", "")[:60]}  samples {tot}')
        for key, n in per_line.most_common(top):
            st = sorted([(v, h) for (kk, h), v in stall.items() if kk == key], reverse=True)[:3]
            print(f'{100 * n / max(tot, 1):5.1f}% {key[0]}:{key[1]} {key[2][:64]} | ' +
                  ', '.join(f'{h[6:]}={v}' for v, h in st))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
