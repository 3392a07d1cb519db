"""Per-kernel SASS mnemonic counts of liboake_b200.so (`cuobjdump -sass`): which kernels carry tcgen05 MMAs
(UTCHMMA), TMEM loads / stores (LDTM / STTM), TMA loads (UTMALDG), mbarrier / tcgen05.commit traffic (SYNCS /
UTCBAR), legacy mma.sync (HMMA), register re-allocation (USETMAXREG) -- so that the Blackwell-native claims of
DESIGN.md can be audited from the tree without a rebuild.  Writes profiles/r2_sass_mnemonics.txt."""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
COLS = ('UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'USETMAXREG', 'HMMA', 'IMMA', 'MUFU', 'LDGSTS',
        'BAR', 'STL', 'LDL')


def main():
    so = ROOT / 'oadp_b200' / 'liboake_b200.so'
    sass = subprocess.run(['cuobjdump', '-sass', str(so)], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    rows = []
    for block in re.split(r'\n\s*Function : ', sass)[1:]:
        name = block.split('\n', 1)[0].strip()
        ins = [re.sub(r'/\*[0-9a-f]+\*/', '', ln).strip() for ln in block.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/\s+\S', ln)]
        ops = collections.Counter()
        for i in ins:
            tok = i.split()
            if not tok:
                continue
            op = tok[1] if tok[0].startswith('@') and len(tok) > 1 else tok[0]
            ops[op.split('.')[0]] += 1
        short = re.sub(r'oake::\(anonymous namespace\)::|\(anonymous namespace\)::|oake::|^void ', '', demangle(name))
        short = re.sub(r'\((?!bool|int).*', '', short)  # drop the parameter list, keep the template arguments
        rows.append((short, len(ins), [ops.get(c, 0) for c in COLS]))
    rows.sort(key=lambda r: (-r[2][0], -r[1]))
    w = max(len(r[0]) for r in rows)
    out = ['# liboake_b200.so, sm_100a: SASS mnemonic counts per kernel (tools/sass_table.py; cuobjdump -sass)',
           '# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor,',
           '# SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, HMMA = mma.sync (legacy pipe), STL / LDL = spills',
           f'{"kernel":{w}s} {"instr":>6s} ' + ' '.join(f'{c:>8s}' for c in COLS)]
    for name, n, counts in rows:
        out.append(f'{name:{w}s} {n:6d} ' + ' '.join(f'{c:8d}' for c in counts))
    text = '\n'.join(out) + '\n'
    (ROOT / 'profiles' / 'r2_sass_mnemonics.txt').write_text(text)
    sys.stdout.write(text)


if __name__ == '__main__':
    main()
