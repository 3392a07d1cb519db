"""Key metrics of every launch in an .ncu-rep (raw page) as a compact table."""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'dur_us', 1e-3), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 1),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1),
        ('dram__bytes_read.sum', 'dram_rd_MB', None), ('dram__bytes_write.sum', 'dram_wr_MB', None),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%', 1),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 1), ('launch__registers_per_thread', 'regs', 1),
        ('launch__grid_size', 'grid', 1), ('launch__block_size', 'block', 1)]


def to_mb(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[unit]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print('kernel | ' + ' | '.join(n for _, n, _ in WANT))
    for r in rows[2:]:
        name = r[ki].split('(')[0].replace('void oake::<unnamed>::', '').replace('oake::<unnamed>::', '')
        vals = []
        for metric, label, scale in WANT:
            if metric not in hdr:
                vals.append('-')
                continue
            i = hdr.index(metric)
            if scale is None:
                vals.append(f'{to_mb(r[i], units[i]):.1f}')
            else:
                v = float(r[i].replace(',', '')) * scale
                if label == 'dur_us' and units[i] == 'us':
                    v = float(r[i].replace(',', ''))
                elif label == 'dur_us' and units[i] == 'ms':
                    v = float(r[i].replace(',', '')) * 1e3
                vals.append(f'{v:.1f}')
        print(name + ' | ' + ' | '.join(vals))


if __name__ == '__main__':
    main(sys.argv[1])
