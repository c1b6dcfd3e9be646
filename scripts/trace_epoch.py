"""Timeline of one whole-epoch graph from a -DPNP_TRACE build (see scripts/trace_iter.py for the build line): every tail
launch (k_update_prox, with the fused forward line pass of the next iteration) and the column passes between them.
Prints, per inner iteration of the epoch, the medians over CTAs of the phase stamps relative to the tail's first CTA start.
    PNP_LIB=pnp_svrg_b200/lib/libpnp_b200_trace.so python scripts/trace_epoch.py [out.json]"""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench

ba = argparse.Namespace(size=int(os.environ.get('SIZE', 2048)), batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
cfg = bench.workload(ba)
prob, run = bench.make_run(cfg, seed=0)
eng = run.eng
eng.psnr_log.append(0.0)
buf = np.zeros((1 << 20, 2), dtype=np.uint64)


def read():
    n = eng.lib.pnp_debug_read(1, buf.ctypes.data, buf.nbytes)
    if n < 0:
        raise SystemExit('trace read failed (library built without -DPNP_TRACE?)')
    raw = buf[:n].copy()
    return raw[:, 0].astype(np.int64), (raw[:, 1] & 0xffffffff).astype(np.int64), ((raw[:, 1] >> 32) & 0xffff).astype(np.int64)


for _ in range(3):
    run.epoch()
eng.stream.synchronize()
read()
run.epoch()
eng.stream.synchronize()
t, tag, cta = read()
t0 = t.min()
rel = (t - t0) * 1e-3
names = {300: 'start', 310: 'inv round0', 311: 'inv round1', 301: 'inv done', 302: 'sigma(w0)', 305: 'haar fwd', 303: 'barrier', 322: 'stored(w0)',
         304: 'shrink done', 330: 'fwd round0', 331: 'fwd round1'}
starts = np.sort(rel[tag == 300])
# launches of the tail: gaps of more than 15 us between CTA starts
cuts = [starts[0]] + [b for a, b in zip(starts[:-1], starts[1:]) if b - a > 15.0]
out = {'epoch_us': float(rel.max()), 'tail_launches': len(cuts), 'iterations': []}
for k, c0 in enumerate(cuts):
    c1 = cuts[k + 1] if k + 1 < len(cuts) else rel.max() + 1
    m = (rel >= c0) & (rel < c1)
    row = {'tail_start_us': float(c0)}
    for tg, nm in names.items():
        v = rel[m & (tag == tg)]
        if len(v):
            row[nm] = [round(float(np.median(v) - c0), 2), round(float(v.max() - c0), 2)]
    for nm, a, b in (('cols', 200, 209), ('r2c', 100, 109)):
        va, vb = rel[m & (tag == a)], rel[m & (tag == b)]
        if len(va):
            row[nm] = {'first_start': round(float(va.min() - c0), 2), 'last_end': round(float(vb.max() - c0), 2)}
    out['iterations'].append(row)
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
