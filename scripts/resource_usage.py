"""Static resource usage of every kernel in libpnp_b200.so (cuobjdump --dump-resource-usage, no GPU needed):
registers, static shared memory, stack / local bytes (spills).  python scripts/resource_usage.py > profiles/...txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'pnp_svrg_b200', 'lib', 'libpnp_b200.so')
out = subprocess.run(['cuobjdump', '--dump-resource-usage', so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function (\S+?):', out)), capture_output=True, text=True).stdout.split('\n')
rows = []
for name, res in zip(names, re.findall(r'Function \S+?:\n\s*(REG:.*)', out)):
    f = dict(kv.split(':') for kv in res.split() if ':' in kv and not kv.startswith('CONSTANT'))
    short = re.sub(r'\(.*', '', name).replace('void ', '').replace('pnp::', '')
    rows.append((short, int(f['REG']), int(f['SHARED']), int(f['STACK']), int(f['LOCAL'])))
print('%-44s %5s %8s %6s %6s' % ('kernel (sm_100a)', 'regs', 'smem(B)', 'stack', 'local'))
for r in sorted(rows):
    print('%-44s %5d %8d %6d %6d' % r)
print('\n%d kernels; kernels with local-memory spills: %d' % (len(rows), sum(1 for r in rows if r[4] > 0)))
