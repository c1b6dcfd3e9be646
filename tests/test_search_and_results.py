"""Host logic either side of the hot path: the hyperopt-shaped search that drives tune_pnp_* in the sweep scripts
(script_diff_sampratio_set12.py:63-134) and display_results' CSV (Utilities.py:5-64).  No GPU needed except for the
one test marked so."""
import csv
import math
from functools import partial

import numpy as np
import pytest

from pnp_svrg_b200 import search, sweep
from pnp_svrg_b200.Utilities import CSV_HEADER, display_results, metrics_row
from pnp_svrg_b200.search import STATUS_OK, Trials, fmin, hp, quniform, rand, scope, space_eval, tpe


def test_space_draws_respect_ranges_and_types():
    space = (hp.uniform('eta', 0, 100), scope.int(quniform('mini_batch_size', 1, 100, q=1)),
             scope.int(quniform('T2', 1, 100, q=1)), hp.uniform('dstrength', 0, 2))
    seen = []

    def fn(args):
        eta, mb, T2, ds = args
        assert isinstance(mb, int) and isinstance(T2, int)
        assert 0 <= eta <= 100 and 1 <= mb <= 100 and 1 <= T2 <= 100 and 0 <= ds <= 2
        seen.append(args)
        return {'loss': abs(eta - 30) + abs(mb - 10), 'status': STATUS_OK}
    tr = Trials()
    best = fmin(fn, space, algo=rand.suggest, max_evals=60, trials=tr, rstate=3)
    assert len(tr) == 60 and len(seen) == 60
    assert set(best) == {'eta', 'mini_batch_size', 'T2', 'dstrength'}
    assert tr.best_trial['result']['loss'] == min(tr.losses())
    # fmin returns the stored values of the best trial (floats for quniform, as hyperopt does)
    assert best == {k: v[0] for k, v in tr.best_trial['misc']['vals'].items()}
    assert space_eval(space, best)[1] == int(best['mini_batch_size'])


def test_seed_reproducible_and_resumable():
    space = {'x': hp.uniform('x', -5, 5), 'k': hp.choice('k', ['a', 'b', 'c'])}
    fn = lambda d: (d['x'] - 1.0) ** 2 + {'a': 1.0, 'b': 0.0, 'c': 2.0}[d['k']]
    t1, t2 = Trials(), Trials()
    b1 = fmin(fn, space, algo=tpe.suggest, max_evals=40, trials=t1, rstate=7)
    b2 = fmin(fn, space, algo=tpe.suggest, max_evals=40, trials=t2, rstate=7)
    assert b1 == b2 and t1.losses() == t2.losses()
    fmin(fn, space, algo=tpe.suggest, max_evals=50, trials=t1, rstate=8)      # continues: ten more trials
    assert len(t1) == 50 and t1.losses()[:40] == t2.losses()
    assert space_eval(space, b1)['k'] in ('a', 'b', 'c') and isinstance(b1['k'], int)


def test_tpe_beats_random_on_a_smooth_objective():
    space = (hp.uniform('a', 0, 100), hp.uniform('b', 0, 2), hp.loguniform('c', math.log(1e-3), math.log(1e3)))
    fn = lambda v: (v[0] - 71.0) ** 2 / 100.0 + (v[1] - 0.4) ** 2 * 25.0 + (math.log10(v[2]) - 1.0) ** 2
    wins = 0
    for seed in range(6):
        tt, tr = Trials(), Trials()
        fmin(fn, space, algo=tpe.suggest, max_evals=120, trials=tt, rstate=seed)
        fmin(fn, space, algo=rand.suggest, max_evals=120, trials=tr, rstate=seed)
        wins += tt.best_trial['result']['loss'] < tr.best_trial['result']['loss']
    assert wins >= 4, wins


def test_failed_trials_are_skipped_and_errors_surface():
    space = (hp.uniform('x', 0, 1),)

    def fn(v):
        if v[0] > 0.5:
            raise RuntimeError('diverged')
        return v[0]
    with pytest.raises(RuntimeError):
        fmin(fn, space, algo=rand.suggest, max_evals=30, rstate=0)
    tr = Trials()
    best = fmin(fn, space, algo=rand.suggest, max_evals=30, trials=tr, rstate=0, catch=True)
    assert best['x'] <= 0.5 and any(l is None for l in tr.losses())
    with pytest.raises(ValueError):
        fmin(lambda v: 1 / 0, space, algo=rand.suggest, max_evals=3, rstate=0, catch=True)     # nothing finished
    with pytest.raises(ValueError):
        fmin(fn, (hp.uniform('x', 0, 1), hp.uniform('x', 0, 1)), max_evals=1)                  # duplicate label
    with pytest.raises(ValueError):
        hp.uniform('x', 1, 1)
    with pytest.raises(TypeError):
        scope.int(hp.choice('c', [1, 2]))


def test_pspace_matches_the_objectives_argument_order():
    """get_pspace yields the tuple tune_pnp_* unpacks (algorithms/pnp_*.py) with the reference's ranges (:37-40)."""
    for algo, labels in sweep.TUNE_ARGS.items():
        sp = sweep.get_pspace(algo)
        assert tuple(n.label for n in sp) == labels
        for n in sp:
            if n.label in ('mini_batch_size', 'T2', 'hist_size'):      # hist_size: the 4th value tune_pnp_saga unpacks
                assert n.as_int and n.kind == 'quniform' and (n.lo, n.hi, n.q) == (1, 100, 1)
            elif n.label == 'eta':
                assert (n.lo, n.hi) == (0, 100)
            else:
                assert (n.lo, n.hi) == (0, 2)
    with pytest.raises(Exception, match='not found'):
        sweep.get_pspace('pnp_adam')
    with pytest.raises(Exception, match='not found'):
        sweep.get_problem('Tomography', None, 0.5, 20.)
    with pytest.raises(Exception, match='not found'):
        sweep.get_denoiser('CNN')


def test_tuning_csv_layout(tmp_path):
    job = dict(id=0, image='01.png', problem='CSMRI', denoiser='TV', algo='pnp_svrg', alpha=0.3, snr=20.)
    space = sweep.get_pspace('pnp_svrg')
    tr = Trials()
    best = fmin(lambda a: {'loss': -a[0] / 10.0, 'status': STATUS_OK}, space, algo=rand.suggest, max_evals=5, trials=tr, rstate=1)
    row = sweep.tuning_row(job, tr, best)
    assert row[:5] == ['CSMRI', 'TV', 'pnp_svrg', 0.3, 20.] and row[6] == 'PARAMETERS:'
    assert row[5] == tr.best_trial['result']['loss'] and row[7::2] == list(best) and len(row) == 7 + 2 * 4
    out = tmp_path / 'tuning.csv'
    sweep.write_tuning_csv(str(out), [dict(row=row), dict(error='x')])
    lines = list(csv.reader(open(out)))
    assert lines[0] == ['Results:'] and len(lines) == 2 and lines[1][0] == 'CSMRI' and lines[1][6] == 'PARAMETERS:'


class _P:
    H = W = 4
    prob_dir = None
    color_map = 'gray'


def test_display_results_csv_and_print(tmp_path, capsys):
    out = dict(z=np.linspace(0, 1, 16), time_per_iter=[0.1, 0.2, 0.3], psnr_per_iter=[10.0, 15.55, 21.349],
               gradient_time=1.234, denoise_time=5.678, algo_name='PnP SVRG')
    p = _P()
    display_results(p, out, save_results=True, save_dir=str(tmp_path) + '/')
    rows = list(csv.reader(open(tmp_path / 'output.csv')))
    assert rows[0] == CSV_HEADER == ['Output PSNR', 'Change in PSNR', 'Gradient Time', 'Denoising Time']
    assert [float(v) for v in rows[1]] == [21.3, 11.35, 1.23, 5.68]
    assert [float(v) for v in metrics_row(out)] == [21.3, 11.35, 1.23, 5.68]
    text = capsys.readouterr().out
    assert 'Output PSNR: 21.3' in text and 'Gradient Time: 1.23' in text and 'Denoising Time: 5.68' in text
    p.prob_dir = str(tmp_path) + '/prob/'                   # set by problem.display(save_results=True): <dir>/<algo>/
    display_results(p, out, save_results=True)
    assert (tmp_path / 'prob' / 'PnP SVRG' / 'output.csv').exists()


@pytest.mark.gpu
def test_tune_job_runs_a_search_on_the_gpu(tmp_path):
    from conftest import synth_image
    job = dict(id=2, image=0, problem='CSMRI', denoiser='TV', algo='pnp_svrg', alpha=0.5, snr=20.)
    rec = sweep.tune_job(job, max_evals=4, tt=1e9, seed=0, H=64, W=64, images=[synth_image(64, 64, 0)], algo='rand',
                         spaces=dict(eta=(0, 50), T2=(1, 4)), max_iters=6)
    assert rec['trials'] == 4 and math.isfinite(rec['loss']) and rec['row'][:3] == ['CSMRI', 'TV', 'pnp_svrg']
    assert set(rec['best']) == {'eta', 'mini_batch_size', 'T2', 'dstrength'}
    job2 = dict(job, id=3, algo='pnp_gd', problem='DeblurSR', alpha=1.0)
    rec2 = sweep.tune_job(job2, max_evals=3, tt=1e9, seed=0, H=64, W=64, images=[synth_image(64, 64, 0)], algo='tpe',
                          max_iters=4)
    assert rec2['trials'] == 3 and set(rec2['best']) == {'eta', 'dstrength'}
    sweep.write_tuning_csv(str(tmp_path / 't.csv'), [rec, rec2])
    assert len(list(csv.reader(open(tmp_path / 't.csv')))) == 3
