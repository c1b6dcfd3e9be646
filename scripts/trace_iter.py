"""Timeline of the inner iteration from a -DPNP_TRACE build (events stamped with %globaltimer by thread 0 of every CTA):

    PNP_NVCC_EXTRA=-DPNP_TRACE PNP_LIB_OUT=pnp_svrg_b200/lib/libpnp_b200_trace.so python -m pnp_svrg_b200.build --force
    PNP_LIB=pnp_svrg_b200/lib/libpnp_b200_trace.so python scripts/trace_iter.py [out.json]

Runs the bench workload (2048^2) through the whole-epoch graph (SvrgRun.epoch), reads the trace of one epoch and
prints, per kernel: when the first / last CTA started and ended (relative to the iteration start), the mean time a
CTA waited for its first data, per-item times, and the gaps between kernels."""
import os, sys, json, argparse, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
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
    t = raw[:, 0].astype(np.int64)
    tag = (raw[:, 1] & 0xffffffff).astype(np.int64)
    cta = ((raw[:, 1] >> 32) & 0xffff).astype(np.int64)
    sm = ((raw[:, 1] >> 48) & 0xffff).astype(np.int64)
    return t, tag, cta, sm


if os.environ.get('DBG'):
    eng.lib.pnp_debug_set(3, int(os.environ['DBG']))
for _ in range(3):
    run.epoch()
eng.stream.synchronize()
read()                                   # drop the warm-up
run.epoch()                              # one whole-epoch graph: snapshot + T2 inner iterations
eng.stream.synchronize()
t, tag, cta, sm = read()
# split into iterations at the selection / r2c starts: use the last full iteration
starts = np.sort(t[tag == 100])
n_r2c = int((tag == 100).sum() // (cfg['T2'] + 1))       # r2c CTAs per launch (the snapshot runs the pass too)
k = 6                                                    # look at the 6th inner iteration of the epoch
it0, it1 = starts[k * n_r2c], starts[(k + 1) * n_r2c]
keep = (t >= it0) & (t < it1)
t, tag, cta, sm = t[keep], tag[keep], cta[keep], sm[keep]
t0 = it0
rel = (t - t0) * 1e-3
out_period = (it1 - it0) * 1e-3
out = {'iteration_period_us': float(out_period)}


def span(name, a, b):
    ta, tb = rel[tag == a], rel[tag == b]
    if len(ta) == 0 or len(tb) == 0:
        return
    out[name] = {'ctas': int(len(ta)), 'first_start_us': float(ta.min()), 'last_start_us': float(ta.max()),
                 'first_end_us': float(tb.min()), 'last_end_us': float(tb.max()), 'median_end_us': float(np.median(tb))}


span('sel', 400, 409)
span('r2c', 100, 109)
span('cols', 200, 209)
span('cols_col0', 209, 210)
span('tail', 300, 304)
for k, (a, b) in {'r2c_first_data_wait': (100, 101), 'cols_first_data_wait': (200, 201)}.items():
    d = []
    for c in np.unique(cta[tag == a]):
        ta = rel[(tag == a) & (cta == c)]
        tb = rel[(tag == b) & (cta == c)]
        if len(ta) and len(tb):
            d.append(tb.min() - ta.min())
    out[k + '_us'] = {'mean': float(np.mean(d)), 'max': float(np.max(d))}
for k, (a, b) in {'r2c_item': (101, 102), 'cols_item': (201, 202)}.items():
    d = []
    for c in np.unique(cta[tag == a]):
        ta = np.sort(rel[(tag == a) & (cta == c)])
        tb = np.sort(rel[(tag == b) & (cta == c)])
        d.extend(list(tb[:len(ta)] - ta[:len(tb)]))
    out[k + '_us'] = {'mean': float(np.mean(d)), 'max': float(np.max(d)), 'n': len(d)}
    first, later = [], []
    for c in np.unique(cta[tag == a]):
        ta = np.sort(rel[(tag == a) & (cta == c)])
        tb = np.sort(rel[(tag == b) & (cta == c)])
        dd = list(tb[:len(ta)] - ta[:len(tb)])
        first.extend(dd[:1])
        later.extend(dd[1:])
    out[k + '_us']['first_item_mean'] = float(np.mean(first)) if first else None
    out[k + '_us']['later_items_mean'] = float(np.mean(later)) if later else None
ph = {}
names = {300: 'start', 310: 'fft_round0_done', 311: 'fft_round1_done', 301: 'fft_done', 302: 'sigma_done(warp0)', 305: 'haar_forward_done', 303: 'barrier_done', 306: 'xrec_landed(warp0)', 320: 'thresholds(warp0)', 321: 'inverse_top(warp0)', 323: 'chunk_loop_first_pass(warp0)', 322: 'stored(warp0)', 304: 'shrink_done'}
for tg, nm in names.items():
    v = rel[tag == tg]
    if len(v):
        ph[nm] = {'min': float(v.min()), 'median': float(np.median(v)), 'max': float(v.max())}
out['tail_phases_us'] = ph
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
