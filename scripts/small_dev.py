"""Development harness of the cluster kernel (csrc/small.cuh): builds scripts/small_dev.cu (seconds), runs it on the
buffers of a BatchedSVRG, checks the result against the three-pass path of the full library, times it with CUDA
events for several batch sizes and, with a -DPNP_TRACE build, prints the phase timeline of one cluster.
usage: python scripts/small_dev.py [--trace] [--L 256] [--out file.json]   (build: --build-only on the CPU box)"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
LIBDIR = os.path.join(ROOT, 'pnp_svrg_b200', 'lib')


def build(trace, extra=(), tag=''):
    out = os.path.join(LIBDIR, 'libsmall_dev%s%s.so' % (tag, '_trace' if trace else ''))
    cmd = ['nvcc', '-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-shared', '-Xcompiler', '-fPIC',
           '-Xptxas', '-v'] + (['-DPNP_TRACE', '-DPNP_TRACE_CTA=96'] if trace else []) + list(extra) + ['-o', out, os.path.join(ROOT, 'scripts', 'small_dev.cu')]
    print(out)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit(1)
    for ln in r.stderr.splitlines():
        if 'spill' in ln or 'Used' in ln:
            print(ln.strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--trace', action='store_true')
    ap.add_argument('--build-only', action='store_true')
    ap.add_argument('--L', type=int, default=256)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--out', default='')
    ap.add_argument('--variant', default='', help="'' or '_seq'")
    ap.add_argument('--cluster', type=int, default=8)
    ap.add_argument('--threads', type=int, default=512)
    a = ap.parse_args()
    if a.build_only:
        build(False)
        build(True)
        return
    import numpy as np
    import torch
    from conftest import rel_l2, synth_image
    from pnp_svrg_b200 import _lib, device as D
    from pnp_svrg_b200.batched import BatchedSVRG, csmri_host_spec
    path = os.path.join(LIBDIR, 'libsmall_dev%s%s.so' % (a.variant, '_trace' if a.trace else ''))
    dev = C.CDLL(path)
    dev.dev_run.argtypes = [C.POINTER(_lib.SvrgSmallArgs), C.c_void_p, C.c_int, C.c_int]
    dev.dev_trace_read.argtypes = [C.c_void_p, C.c_longlong]
    L, B, T2 = a.L, (1000 if a.L == 256 else 300), 10
    os.environ['PNP_SMALL'] = '0'
    res = {'L': L, 'variant': a.variant, 'cluster': a.cluster, 'threads': a.threads}

    def specs_of(nb):
        return [csmri_host_spec(synth_image(L, L, s % 12), L, L, [0.3, 0.5, 0.7][s % 3], 20., rng=np.random.RandomState(s)) for s in range(nb)]

    def make(nb):
        sp = specs_of(nb)
        return BatchedSVRG(sp, T2=T2, mini_batch_size=B, etas=[min(0.15 * s['M0'], 3.0 * B) for s in sp], seed=3)

    def run_dev(b, n):
        args = _lib.SvrgSmallArgs(
            H=L, W=L, batch=b.nb, z=D.ptr(b.z), xrec=D.ptr(b.xrec), Y1=D.ptr(b.Y1), Y2=D.ptr(b.Y2), Y1n=D.ptr(b.Y1n), Y2n=D.ptr(b.Y2n),
            bits_full=D.ptr(b.bits_full), support=D.ptr(b.support), m0=D.ptr(b.m0), support_img_stride=b.sup_stride,
            idx=None, idx_img_stride=0, idx_iter_stride=0, snap_scale_ptr=D.ptr(b.inv_m0), snap_scale=0.0,
            step=D.ptr(b.step), step_img_stride=1, sig_log=D.ptr(b.sig_log), mse_log=D.ptr(b.mse_log),
            slot=D.ptr(b.counters[0:1]), draw_counter=D.ptr(b.counters[2:3]), n_inner=n, T2=T2, mini_batch_size=B,
            seed=b.seed & 0xffffffff, lr_decay=1.0, sigma_modifier=1.0, fallback_sigma=0.0, fallback_decay=1.0)
        rc = dev.dev_run(C.byref(args), b.sptr, a.cluster, a.threads)
        assert rc == 0, rc

    torch.cuda.set_device(0)
    _lib.init_device()
    assert dev.dev_init() == 0
    res['max_active_clusters'] = dev.dev_max_clusters(L, a.cluster, a.threads)
    # ---- parity against the three-pass path
    ref, dut = make(3), make(3)
    ref.run(25)
    run_dev(dut, 25)
    dut.slots_used = 25
    r, d = ref.results(), dut.results()
    res['rel_l2_vs_three_pass'] = [rel_l2(d['z'][i], r['z'][i]) for i in range(3)]
    res['psnr_max_diff'] = float(np.max(np.abs(r['psnr'] - d['psnr'])))
    ref.close(); dut.close()
    # ---- timing
    res['timing'] = {}
    for nb in (1, 15, 30, 120):
        b = make(nb)
        run_dev(b, 20)
        b.stream.synchronize()
        b.mse_log.zero_(); b.sig_log.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(b.stream)
        run_dev(b, a.iters)
        e1.record(b.stream)
        b.stream.synchronize()
        ms = e0.elapsed_time(e1)
        res['timing'][nb] = {'ms': ms, 'us_per_iteration_per_launch': 1e3 * ms / a.iters, 'image_iterations_per_s': nb * a.iters / (ms * 1e-3)}
        b.close()
    # ---- trace of one cluster
    if a.trace:
        b = make(1)
        run_dev(b, 3)
        b.stream.synchronize()
        buf = np.zeros((1 << 16, 2), dtype=np.uint64)
        dev.dev_trace_read(buf.ctypes.data, buf.nbytes)
        run_dev(b, 14)
        b.stream.synchronize()
        n = dev.dev_trace_read(buf.ctypes.data, buf.nbytes)
        raw = buf[:n]
        t = raw[:, 0].astype(np.int64)
        tag = (raw[:, 1] & 0xffffffff).astype(np.int64)
        cta = ((raw[:, 1] >> 32) & 0xffff).astype(np.int64)
        names = {510: 'start', 511: 'A done', 512: 'sync1', 513: 'B done', 514: 'sync2', 515: 'C done', 516: 'sigma+haar fwd', 517: 'cta red',
                 518: 'sync3', 519: 'shrink'}
        tl = {}
        for c in np.unique(cta):
            m = cta == c
            tc, gc = t[m], tag[m]
            o = np.argsort(tc, kind='stable')
            tc, gc = tc[o], gc[o]
            # the last full inner iteration: from the last 510 that is followed by a 519
            starts = np.flatnonzero(gc == 510)
            s0 = [s for s in starts if s + 10 <= len(gc) and gc[s + 9] == 519][-1]
            seg_t, seg_g = tc[s0:s0 + 10], gc[s0:s0 + 10]
            tl[int(c)] = {names[int(g)]: float((x - seg_t[0]) * 1e-3) for x, g in zip(seg_t, seg_g)}
            prev = [s for s in starts if s < s0]
            if prev:
                tl[int(c)]['period'] = float((tc[s0] - tc[prev[-1]]) * 1e-3)
        res['timeline_us_cta0'] = tl[min(tl)]
        b.close()
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        open(a.out, 'w').write(s)


if __name__ == '__main__':
    main()
