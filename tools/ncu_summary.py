"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a markdown table for profiles/.
    python tools/ncu_summary.py gpurun_out/prof_r01_final.ncu-rep profiles/r01_ncu_final.md"""
import csv
import io
import json
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor inst'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('launch__cluster_size', 'cluster'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %')]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in WANT if m in idx]
    lines = ['| # | kernel | ' + ' | '.join(n for _, n in cols) + ' |', '|---|---|' + '---|' * len(cols)]
    traffic = {}
    for k, r in enumerate(rows[2:]):
        name = r[idx['Kernel Name']]
        vals = []
        for m, n in cols:
            v = r[idx[m]]
            try:
                v = f'{float(v.replace(",", "")):.4g}'
            except ValueError:
                pass
            vals.append(f'{v} {units[idx[m]]}'.strip())
        lines.append(f'| {k} | `{name[:70]}` | ' + ' | '.join(vals) + ' |')
        try:
            def b(m):
                v, u = float(r[idx[m]].replace(',', '')), units[idx[m]]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            key = name.split('(')[0].replace('void ', '').split('<')[0]
            traffic.setdefault(key, []).append(b('dram__bytes_read.sum') + b('dram__bytes_write.sum'))
        except Exception:
            pass
    open(out, 'w').write(f'# ncu --set full summary of `{rep}`\n\n' + '\n'.join(lines) + '\n')
    json.dump({k + '_bytes_per_launch': round(sum(v) / len(v)) for k, v in traffic.items()},
              open(out.replace('.md', '_traffic.json'), 'w'), indent=1)
    print('\n'.join(lines[:20]))


if __name__ == '__main__':
    main()
