"""profiles/ncu_traffic.json from an `ncu --set full --nvtx --nvtx-include "<class>/"` capture of ONE step of bench.py:
per kernel class (the NVTX range encoder.cu pushes around every launch), the mean dram__bytes_read.sum +
dram__bytes_write.sum per launch and the mean algorithmic FLOPs per launch of the same launches (from the grid's
M: recorded by bench.py --dump-launch-flops).  bench.py prints `roofline.traffic` from this file only when its own
run reports the same FLOPs per launch (same step mix)."""
import csv
import json
import pathlib
import subprocess
import sys


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main(rep, kernel_class, flops_per_launch, out_path):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    tot, n, dur = 0.0, 0, 0.0
    for r in rows[2:]:
        tot += to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
        dur += float(r[it].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[it]]
        n += 1
    path = pathlib.Path(out_path)
    table = json.loads(path.read_text()) if path.exists() else {}
    table[kernel_class] = dict(dram_bytes_per_launch=tot / max(n, 1), launches=n, flops_per_launch=float(flops_per_launch),
                               us_per_launch_under_ncu=dur / max(n, 1), capture=pathlib.Path(rep).name)
    path.write_text(json.dumps(table, indent=1) + '\n')
    print(json.dumps(table[kernel_class]))


if __name__ == '__main__':
    main(*sys.argv[1:5])
