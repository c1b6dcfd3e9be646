"""Hyper-parameter search with the call shape the reference's sweep scripts use from hyperopt
(script_diff_sampratio_set12.py:1-5,63-107,121-134 / script_diff_snr_set12.py: ``fmin(proxy, space=pspace,
algo=tpe.suggest, trials=Trials(), max_evals=MAX_EVALS)``, ``hp.uniform``, ``scope.int(quniform(...))``,
``trials.best_trial['result']['loss']``).  hyperopt is a third-party dependency that is absent from this
image; this module is the host-side replacement for the subset those scripts touch, so the
``tune_pnp_*`` objectives (algorithms/pnp_*.py) can be driven without it:

    from pnp_svrg_b200.search import fmin, tpe, hp, Trials, scope, quniform
    pspace = (hp.uniform('eta', 0, 100), scope.int(quniform('mini_batch_size', 1, 100, q=1)),
              scope.int(quniform('T2', 1, 100, q=1)), hp.uniform('dstrength', 0, 2))
    trials = Trials()
    best = fmin(partial(tune_pnp_svrg, problem=p, denoiser=d, tt=30), space=pspace, algo=tpe.suggest,
                trials=trials, max_evals=100, rstate=np.random.default_rng(0))

``rand.suggest`` draws every trial from the prior.  ``tpe.suggest`` is a tree-structured Parzen estimator over
independent dimensions (Bergstra et al. 2011: split the finished trials at the gamma quantile of the loss, model
the good and the bad group by Gaussian mixtures centred on their values plus the prior, propose the candidate
that maximises l(x)/g(x)); it is the published algorithm, not hyperopt's code, and the sequences it draws are
its own -- seeds do not reproduce hyperopt runs.  All of it is host logic: no GPU work happens here.
"""
import math

import numpy as np

STATUS_OK = 'ok'
STATUS_FAIL = 'fail'


# ------------------------------------------------------------------------------------------ search space
class Node:
    """One labelled dimension.  kind: 'uniform' | 'quniform' | 'loguniform' | 'randint' | 'choice'."""

    def __init__(self, kind, label, lo=None, hi=None, q=None, options=None, as_int=False):
        if not isinstance(label, str):
            raise TypeError('the first argument of an hp.* node is its label (a string)')
        if kind != 'choice' and not (hi > lo):
            raise ValueError('%s(%r): empty range [%r, %r]' % (kind, label, lo, hi))
        if kind == 'quniform' and not (q and q > 0):
            raise ValueError('quniform(%r): q must be positive' % label)
        if kind == 'choice' and not options:
            raise ValueError('choice(%r): no options' % label)
        self.kind, self.label, self.lo, self.hi, self.q, self.options, self.as_int = kind, label, lo, hi, q, options, as_int

    # the stored value ("vals") is what hyperopt reports: the number itself, or the option index of a choice
    def prior(self, rng):
        if self.kind == 'choice':
            return int(rng.integers(len(self.options)))
        if self.kind == 'randint':
            return int(rng.integers(self.lo, self.hi))
        return self._finish(rng.uniform(self.lo, self.hi))

    def _finish(self, u):
        """u lives on the sampling axis (the log axis for loguniform)."""
        if self.kind == 'loguniform':
            return float(math.exp(u))
        if self.kind == 'quniform':
            return float(np.round(u / self.q) * self.q)
        if self.kind == 'randint':
            return int(min(max(int(np.floor(u)), self.lo), self.hi - 1))
        return float(u)

    def axis(self, v):
        return math.log(v) if self.kind == 'loguniform' else float(v)

    def value(self, stored):
        """What the objective receives."""
        if self.kind == 'choice':
            return self.options[stored]
        return int(stored) if self.as_int else stored


class _Hp:
    @staticmethod
    def uniform(label, low, high):
        return Node('uniform', label, low, high)

    @staticmethod
    def quniform(label, low, high, q):
        return Node('quniform', label, low, high, q=q)

    @staticmethod
    def loguniform(label, low, high):
        return Node('loguniform', label, low, high)

    @staticmethod
    def randint(label, low, high=None):
        return Node('randint', label, 0 if high is None else low, low if high is None else high)

    @staticmethod
    def choice(label, options):
        return Node('choice', label, options=list(options))


class _Scope:
    @staticmethod
    def int(node):
        if not isinstance(node, Node) or node.kind == 'choice':
            raise TypeError('scope.int wraps a numeric hp.* node')
        return Node(node.kind, node.label, node.lo, node.hi, q=node.q, as_int=True)


hp = _Hp()
scope = _Scope()
quniform = _Hp.quniform          # ``from hyperopt.hp import quniform`` in the reference scripts


def _nodes(space):
    """Flatten a tuple / list / dict space into its nodes (depth first, stable order)."""
    if isinstance(space, Node):
        return [space]
    if isinstance(space, dict):
        return [n for k in space for n in _nodes(space[k])]
    if isinstance(space, (tuple, list)):
        return [n for s in space for n in _nodes(s)]
    return []                     # constants are allowed inside a space


def _instantiate(space, vals):
    if isinstance(space, Node):
        return space.value(vals[space.label])
    if isinstance(space, dict):
        return {k: _instantiate(v, vals) for k, v in space.items()}
    if isinstance(space, tuple):
        return tuple(_instantiate(s, vals) for s in space)
    if isinstance(space, list):
        return [_instantiate(s, vals) for s in space]
    return space


# ------------------------------------------------------------------------------------------ trials
class Trials:
    """The slice of hyperopt.Trials the reference reads: ``trials``, ``results``, ``losses()``, ``best_trial``."""

    def __init__(self):
        self.trials = []

    def _add(self, vals, result):
        self.trials.append({'tid': len(self.trials), 'result': result,
                            'misc': {'vals': {k: [v] for k, v in vals.items()}}})

    @property
    def results(self):
        return [t['result'] for t in self.trials]

    def losses(self):
        return [t['result'].get('loss') if t['result'].get('status') == STATUS_OK else None for t in self.trials]

    def _ok(self):
        return [t for t in self.trials if t['result'].get('status') == STATUS_OK and t['result'].get('loss') is not None
                and not math.isnan(t['result']['loss'])]

    @property
    def best_trial(self):
        ok = self._ok()
        if not ok:
            raise ValueError('no trial finished with status %r' % STATUS_OK)     # hyperopt raises AllTrialsFailed
        return min(ok, key=lambda t: t['result']['loss'])

    @property
    def argmin(self):
        return {k: v[0] for k, v in self.best_trial['misc']['vals'].items()}

    def __len__(self):
        return len(self.trials)


# ------------------------------------------------------------------------------------------ suggesters
class _Rand:
    @staticmethod
    def suggest(nodes, trials, rng):
        return {n.label: n.prior(rng) for n in nodes}


class _Tpe:
    gamma = 0.25
    n_startup = 20
    n_candidates = 24

    @classmethod
    def suggest(cls, nodes, trials, rng):
        done = trials._ok()
        if len(done) < cls.n_startup:
            return _Rand.suggest(nodes, trials, rng)
        done = sorted(done, key=lambda t: t['result']['loss'])
        n_good = max(1, min(int(math.ceil(cls.gamma * math.sqrt(len(done)))), 25))
        good, bad = done[:n_good], done[n_good:]
        return {n.label: cls._dimension(n, [t['misc']['vals'][n.label][0] for t in good],
                                        [t['misc']['vals'][n.label][0] for t in bad], rng) for n in nodes}

    @classmethod
    def _dimension(cls, n, good, bad, rng):
        if n.kind == 'choice' or n.kind == 'randint':
            lo, k = (0, len(n.options)) if n.kind == 'choice' else (n.lo, n.hi - n.lo)
            # categorical posterior with a uniform pseudo-count of one per option
            pg = np.bincount(np.asarray(good, dtype=np.int64) - lo, minlength=k) + 1.0
            pb = np.bincount(np.asarray(bad, dtype=np.int64) - lo, minlength=k) + 1.0
            pg, pb = pg / pg.sum(), pb / pb.sum()
            cand = rng.choice(k, size=cls.n_candidates, p=pg)
            return int(cand[np.argmax(np.log(pg[cand]) - np.log(pb[cand]))]) + lo
        lo, hi = float(n.lo), float(n.hi)
        g = _Parzen([n.axis(v) for v in good], lo, hi)
        b = _Parzen([n.axis(v) for v in bad], lo, hi)
        cand = g.sample(rng, cls.n_candidates)
        best = cand[np.argmax(g.logpdf(cand) - b.logpdf(cand))]
        return n._finish(best)


class _Parzen:
    """Mixture of truncated Gaussians on [lo, hi]: one per observation plus the prior (centre of the range, width
    of the range); each observation's width is the larger distance to its sorted neighbours, clipped to
    [range / min(100, 1 + n), range]."""

    def __init__(self, obs, lo, hi):
        span = hi - lo
        mus = np.asarray(list(obs) + [0.5 * (lo + hi)], dtype=np.float64)
        order = np.argsort(mus)
        s = mus[order]
        left = np.diff(s, prepend=lo)
        right = np.diff(s, append=hi)
        sig = np.maximum(left, right)
        sig = np.clip(sig, span / min(100.0, 1.0 + len(s)), span)
        sig[np.searchsorted(s, 0.5 * (lo + hi))] = span          # the prior component keeps the full width
        self.mu, self.sig, self.lo, self.hi = s, sig, lo, hi
        self.w = np.full(len(s), 1.0 / len(s))

    def sample(self, rng, k):
        out = np.empty(k)
        for i in range(k):
            j = rng.integers(len(self.mu))
            for _ in range(64):                                   # rejection onto [lo, hi]
                x = rng.normal(self.mu[j], self.sig[j])
                if self.lo <= x <= self.hi:
                    break
            else:
                x = min(max(self.mu[j], self.lo), self.hi)
            out[i] = x
        return out

    def logpdf(self, x):
        x = np.asarray(x, dtype=np.float64)[:, None]
        z = (x - self.mu[None, :]) / self.sig[None, :]
        # mass of each component inside [lo, hi] (truncation normaliser)
        cdf = lambda t: 0.5 * (1.0 + np.vectorize(math.erf)(t / math.sqrt(2.0)))
        mass = np.maximum(cdf((self.hi - self.mu) / self.sig) - cdf((self.lo - self.mu) / self.sig), 1e-12)
        dens = self.w[None, :] * np.exp(-0.5 * z * z) / (self.sig[None, :] * math.sqrt(2.0 * math.pi) * mass[None, :])
        return np.log(np.maximum(dens.sum(axis=1), 1e-300))


rand = _Rand()
tpe = _Tpe()


# ------------------------------------------------------------------------------------------ driver
def fmin(fn, space, algo=None, max_evals=100, trials=None, rstate=None, catch=False, verbose=False):
    """Minimise ``fn`` over ``space`` for ``max_evals`` trials (including the ones ``trials`` already holds) and return
    {label: best stored value} like hyperopt.fmin.  ``fn`` gets the space with every node replaced by its draw and
    returns a float or a dict with 'loss' and 'status' (the tune_pnp_* wrappers return the latter).  ``algo`` is
    ``tpe.suggest`` (default) or ``rand.suggest``; ``rstate`` a numpy Generator or an int seed.  With ``catch`` an
    exception raised by ``fn`` becomes a failed trial instead of propagating."""
    nodes = _nodes(space)
    labels = [n.label for n in nodes]
    if len(set(labels)) != len(labels):
        raise ValueError('duplicate label in the search space: %r' % labels)
    if not nodes:
        raise ValueError('the search space holds no hp.* node')
    suggest = algo if algo is not None else tpe.suggest
    rng = rstate if isinstance(rstate, np.random.Generator) else np.random.default_rng(rstate)
    trials = trials if trials is not None else Trials()
    while len(trials) < max_evals:
        vals = suggest(nodes, trials, rng)
        try:
            res = fn(_instantiate(space, vals))
        except Exception as e:
            if not catch:
                raise
            res = {'status': STATUS_FAIL, 'error': repr(e)}
        if not isinstance(res, dict):
            res = {'loss': float(res), 'status': STATUS_OK}
        trials._add(vals, res)
        if verbose:
            print('trial %d: %s -> %s' % (len(trials) - 1, vals, res.get('loss')))
    return trials.argmin


def space_eval(space, vals):
    """The space with the stored values of ``vals`` (fmin's return) put in place -- choices resolved to options."""
    return _instantiate(space, vals)
