"""Where a kernel's instructions and stall samples go, by source file / line, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (argv[1]); argv[2] = number of per-file sections that make
up the first captured launch (the export repeats one section per source file per launch), argv[3] = lines to list."""
import collections
import csv
import sys


def main():
    nsec = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    agg, sec, cur, fn = collections.OrderedDict(), -1, None, None
    for r in csv.reader(open(sys.argv[1])):
        if not r:
            continue
        if r[0] == 'Function Name':
            sec += 1
            fn = fn or r[1]
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            ii, si = r.index('Instructions Executed'), r.index('# Samples')
            continue
        if sec >= nsec or not r[0].isdigit():
            continue
        key = (cur, int(r[0]))
        a = agg.get(key, (0, 0, ''))
        num = lambda v: int(v) if v not in ('', '-') else 0
        agg[key] = (a[0] + num(r[ii]), a[1] + num(r[si]), r[1].strip()[:100])
    tot = sum(v[0] for v in agg.values())
    tots = max(sum(v[1] for v in agg.values()), 1)
    print('kernel: %s' % fn)
    print('warp instructions executed: %d, stall samples: %d' % (tot, tots))
    by, bys = collections.Counter(), collections.Counter()
    for (f, l), v in agg.items():
        by[f] += v[0]
        bys[f] += v[1]
    print('%-34s %8s %9s' % ('source file', 'inst %', 'samples %'))
    for f, v in by.most_common():
        print('%-34s %8.1f %9.1f' % (f, 100.0 * v / tot, 100.0 * bys[f] / tots))
    print('\ntop lines by instructions executed')
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * v[0] / tot, 100.0 * v[1] / tots, f, l, v[2]))
    print('\ntop lines by stall samples')
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * v[1] / tots, 100.0 * v[0] / tot, f, l, v[2]))


if __name__ == '__main__':
    main()
