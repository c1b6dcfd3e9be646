"""Summaries of ncu exports kept under profiles/.

  python scripts/ncu_summarise.py launches <ncu --csv launch list> <out.csv>
      per-kernel launches / average / total duration / share of the run (gpu__time_duration.sum)
  python scripts/ncu_summarise.py full <ncu -i rep --page raw --csv> <out.csv>
      one row per captured kernel with the metrics bench.py and profiles/README.md quote
"""
import collections
import csv
import re
import sys

FULL = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'^pnp::', '', name)
    return re.sub(r'\(.*$', '', name).replace('(int)', '').replace('(bool)', '')


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    acc = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != 'gpu__time_duration.sum':
            continue
        acc.setdefault(short(r[ki]), []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in acc.values())
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'avg_ns', 'total_ns', 'share'])
        for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k, len(v), int(sum(v) / len(v)), int(sum(v)), '%.4f' % (sum(v) / tot)])


def full(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]        # ncu's "==PROF==" lines and blank lines drop out
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in FULL if m in hdr]
    ki = hdr.index('Kernel Name')
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['Kernel Name'] + [hdr[c] for c in cols])
        w.writerow([''] + [units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[ki])] + [r[c] for c in cols])


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
