"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`) into a
markdown table: python tools/launch_list.py gpurun_out/X.csv profiles/rNN_launches.md "<command that was profiled>" """
import csv
import sys
from collections import defaultdict


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
    rows = [r for r in csv.reader(open(src, errors='replace')) if r]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1e-3)          # -> microseconds
        a = agg[r[ki]]
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    out = [f'# ncu launch list\n\nCommand: `{cmd}`\n\nPer-launch times under ncu are cold-cache and serialised: compare shares, not '
           f'absolutes. {sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms in total.\n',
           '| kernel | launches | total ms | share | avg us |', '|---|---|---|---|---|']
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| `{k[:110]}` | {n} | {us / 1e3:.2f} | {100 * us / total:.1f}% | {us / n:.1f} |')
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:14]))


if __name__ == '__main__':
    main()
