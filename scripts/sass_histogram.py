"""SASS opcode histogram per kernel of libpnp_b200.so (cuobjdump -sass, no GPU needed):
    python scripts/sass_histogram.py [lib.so] > profiles/rNN_sass_opcode_histogram.csv
One row per kernel: instruction count, the Blackwell-specific opcodes that prove the hand-written paths (UTCHMMA =
tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier, UCGABAR = cluster barrier, FADD2 / FMUL2 / FFMA2 = packed fp32), then the ten most
frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPECIAL = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'UCGABAR', 'FADD2', 'FMUL2', 'FFMA2', 'HMMA', 'MEMBAR', 'ATOMS',
           'ATOMG', 'RED', 'BAR', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'pnp_svrg_b200', 'lib', 'libpnp_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r'\s+Function : (\S+)', ln)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('kernel,instructions,' + ','.join(SPECIAL) + ',top10')
    for (mangled, c), name in zip(kernels.items(), names):
        name = re.sub(r'\(.*$', '', name.replace('void ', '').replace('pnp::', '')).replace(',', ';').replace('(int)', '').replace('(bool)', '')
        top = ' '.join('%s:%d' % kv for kv in c.most_common(10))
        print('%s,%d,%s,%s' % (name, sum(c.values()), ','.join(str(sum(v for k, v in c.items() if k == s or k.startswith(s + '_'))) for s in SPECIAL), top))


if __name__ == '__main__':
    main()
