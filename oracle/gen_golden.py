"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under the
shims of oracle/refshim.py.  Run in the build container only:

    cd /root/repo && python -m oracle.gen_golden

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every fixture stores its inputs (the uint8
image after the reference's own PIL load/resize, seeds, keyword arguments) next to the outputs
the reference produced, so the fixtures are self-contained on the GPU box where /root/reference
does not exist.  The reference is unseeded; we seed NumPy's global RNG ourselves:
np.random.seed(seed_problem) before the problem constructor and np.random.seed(seed_run) before
the algorithm.  The denoiser calls go through oracle.skimage_port (scikit-image is not installed),
everything else is the reference's own NumPy code.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def _image_u8(path, H, W):
    from PIL import Image
    return np.array(Image.open(path).resize((H, W)))


def _kernel_u8(path, H, W):
    from PIL import Image
    return np.array(Image.open(path).resize((H, W)))


def main():
    sys.path.insert(0, ROOT)
    from oracle.refshim import Reference
    os.chdir(REF)
    R = Reference(REF)
    os.makedirs(OUT, exist_ok=True)
    img01 = 'data/Set12/01.png'
    img04 = 'data/Set12/04.png'
    k25 = 'data/kernel25.png'

    def run_case(name, pkind, pkw, algo, akw, budget, den='tv', seed_problem=0, seed_run=1, extra_inputs=None):
        H, W = pkw['H'], pkw['W']
        np.random.seed(seed_problem)
        if pkind == 'csmri':
            p = R.problems.CSMRI(pkw['img'], H=H, W=W, sample_prob=pkw['sample_prob'], snr=pkw['snr'])
        elif pkind == 'deblur':
            p = R.problems.Deblur(pkw['img'], H=H, W=W, kernel_path=pkw.get('kernel_path'), kernel=pkw.get('kernel'),
                                  scale_percent=pkw['scale_percent'], snr=pkw['snr'])
        elif pkind == 'pr':
            p = R.problems.PhaseRetrieval(pkw['img'], H=H, W=W, num_meas=pkw['num_meas'], snr=pkw['snr'])
        d = R.TV.TVDenoiser() if den == 'tv' else R.NLM.NLMDenoiser()
        if den == 'nlm':
            d.sigma = 1.0        # the attribute the reference forgot to set (denoisers/NLM.py:24): any value > 0
        # gradients at a fixed random point (before the run consumes the RNG)
        zr = np.random.default_rng(7).uniform(0, 1, p.N)
        g_full = p.grad_full(zr)
        np.random.seed(5)
        mbsz = akw.get('mini_batch_size', 16)
        mb = p.select_mb(mbsz)
        g_st = p.grad_stoch(zr, mb)
        log = R.record(p, d)
        np.random.seed(seed_run)
        out = R.run(algo, p, d, budget, converge_check=False, **akw)
        meta = dict(name=name, problem=pkind, problem_kwargs={k: v for k, v in pkw.items() if k != 'img'},
                    image_file=pkw['img'], algo=algo, algo_kwargs=akw, budget=budget, denoiser=den,
                    seed_problem=seed_problem, seed_run=seed_run, mb_size_for_grad=mbsz,
                    reference_files=['problems/*.py', 'algorithms/%s.py' % algo, 'denoisers/%s.py' % ('TV' if den == 'tv' else 'NLM')])
        arrays = dict(
            image_u8=_image_u8(pkw['img'], H, W),
            Xinit=p.Xinit, Y=np.asarray(p.Y), sigma=np.float64(p.sigma), M=np.int64(p.M),
            z_rand=zr, grad_full=g_full, mb_idx=np.flatnonzero(np.asarray(mb).ravel()).astype(np.int32),
            grad_stoch=g_st,
            z_final=out['z'], psnr=np.array(out['psnr_per_iter'], dtype=np.float64),
            sigma_est=np.array(log['sigma_est']),
            mb_stream=(np.stack(log['mb']) if log['mb'] else np.zeros((0, 0), np.int32)),
            iterates=np.stack([a.ravel() for a in (log['denoised'] if p.N <= 4096 else log['denoised'][-2:])]).astype(np.float32),
        )
        if pkind == 'csmri':
            arrays['mask'] = p.mask.astype(np.uint8)
            arrays['M0'] = np.int64(p.M0)
        if pkind == 'deblur' and pkw.get('kernel_path'):
            arrays['kernel_u8'] = _kernel_u8(pkw['kernel_path'], H, W)
        if pkind == 'pr':        # A itself is 4 MiB: it is reproduced from seed_problem (np.random.randn)
            arrays['A_checksum'] = np.array([p.A.sum(), np.abs(p.A).sum(), p.A[0, 0], p.A[-1, -1]])
        if extra_inputs:
            arrays.update(extra_inputs)
        np.savez_compressed(os.path.join(OUT, 'ref_%s.npz' % name), meta=json.dumps(meta), **arrays)
        print('%-28s psnr %6.2f -> %6.2f  (%d prox calls, %d KiB)' % (
            name, out['psnr_per_iter'][0], out['psnr_per_iter'][-1], len(log['denoised']),
            os.path.getsize(os.path.join(OUT, 'ref_%s.npz' % name)) // 1024))

    cs = dict(img=img01, H=64, W=64, sample_prob=0.3, snr=20.)
    run_case('csmri64_gd', 'csmri', cs, 'pnp_gd', dict(eta=400.0), 8)
    run_case('csmri64_sgd', 'csmri', cs, 'pnp_sgd', dict(eta=150.0, mini_batch_size=200), 8)
    run_case('csmri64_svrg', 'csmri', cs, 'pnp_svrg', dict(eta=400.0, T2=3, mini_batch_size=100), 9)
    run_case('csmri64_saga', 'csmri', cs, 'pnp_saga', dict(eta=100.0, mini_batch_size=200, hist_size=5), 8)
    run_case('csmri64_sarah', 'csmri', cs, 'pnp_sarah', dict(eta=100.0, T2=4, mini_batch_size=200), 10)
    run_case('csmri256_svrg', 'csmri', dict(img=img01, H=256, W=256, sample_prob=0.3, snr=20.), 'pnp_svrg',
             dict(eta=6000.0, T2=10, mini_batch_size=1000), 12)
    db = dict(img=img01, H=64, W=64, kernel_path=k25, scale_percent=50, snr=20.)
    run_case('deblur64_s50_saga', 'deblur', db, 'pnp_saga', dict(eta=0.3, mini_batch_size=100, hist_size=5), 8)
    run_case('deblur64_s100_gd', 'deblur', dict(db, scale_percent=100), 'pnp_gd', dict(eta=1.0), 6)
    run_case('deblur64_min_sgd', 'deblur', dict(img=img01, H=64, W=64, kernel='Minimal', scale_percent=100, snr=20.),
             'pnp_sgd', dict(eta=300.0, mini_batch_size=400), 6)
    run_case('deblur32_nlm_saga', 'deblur', dict(img=img01, H=32, W=32, kernel_path=k25, scale_percent=50, snr=20.),
             'pnp_saga', dict(eta=0.3, mini_batch_size=50, hist_size=4), 4, den='nlm')
    pr = dict(img=img04, H=32, W=32, num_meas=512, snr=20.)
    run_case('pr32_svrg', 'pr', pr, 'pnp_svrg', dict(eta=0.05, T2=4, mini_batch_size=64), 8)
    run_case('pr32_sarah', 'pr', pr, 'pnp_sarah', dict(eta=0.03, T2=3, mini_batch_size=64), 8)
    run_case('pr32_sgd', 'pr', pr, 'pnp_sgd', dict(eta=0.05, mini_batch_size=64), 6)

    # G1: the one deterministic known answer the reference ships (create_paper_figures_deblur.ipynb
    # cell 4 prints sigma = 0.0015155036596592854, M = 65536)
    p = R.problems.Deblur('./data/Set12/01.png', kernel='Minimal', H=256, W=256, snr=5., scale_percent=100)
    np.savez_compressed(os.path.join(OUT, 'ref_G1_deblur_sigma.npz'),
                        meta=json.dumps(dict(name='G1', notebook_value=0.0015155036596592854, M=65536)),
                        image_u8=_image_u8('./data/Set12/01.png', 256, 256), sigma=np.float64(p.sigma),
                        M=np.int64(p.M))
    print('G1 sigma', repr(float(p.sigma)), 'M', p.M)


if __name__ == '__main__':
    main()
