"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv` (reads the CSV on stdin or argv[1]).
Prints, per kernel, total samples and the N instructions with the most stall samples plus their dominant stall reason."""
import csv
import sys

def main():
    src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kern, hdr, rows = None, None, []
    out = []
    def flush():
        if kern is None or not rows:
            return
        si = hdr.index('# Samples')
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(r[si] or 0) for r in rows)
        print('== %s  (samples %d, sass instructions %d)' % (kern[:70], tot, len(rows)))
        agg = {}
        for r in rows:
            for i in stall_cols:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
        print('   by reason:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        order = sorted(range(len(rows)), key=lambda j: -int(rows[j][si] or 0))[:topn]
        for j in sorted(order):
            r = rows[j]
            st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
            print('   #%5d %5.1f%%  %-60s %s' % (j, 100.0 * int(r[si] or 0) / max(tot, 1), r[1].strip()[:60], ' '.join('%s=%d' % (n, v) for v, n in st if v)))
    for r in csv.reader(src):
        if not r:
            continue
        if r[0] == 'Kernel Name':
            flush()
            kern, hdr, rows = r[1], None, []
        elif r[0] == 'Address':
            hdr = r
        elif hdr is not None:
            rows.append(r)
    flush()

main()
