"""CPU: the C-ABI library loads and exports every symbol include/pnp_b200.h declares (no compute
calls without a GPU); host-side logic (minibatch container, sampler twin, budgets)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'pnp_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pnp_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as g
    g.build()
    from pnp_svrg_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'library does not export %s' % n
        assert n in _lib.PROTOTYPES, 'ctypes prototype missing for %s' % n
    assert set(_lib.PROTOTYPES) == set(names)
    assert _lib.load().pnp_version() >= 100
    # the args struct mirrors the header field for field
    hdr = open(os.path.join(ROOT, 'include', 'pnp_b200.h')).read()
    body = re.sub(r'/\*.*?\*/', '', hdr[hdr.index('typedef struct {'):hdr.index('} pnp_csmri_grad_args;')], flags=re.S)
    fields = [f for stmt in body.split(';') for f in re.findall(r'[\*\s,]([A-Za-z_][A-Za-z0-9_]*)\s*(?=,|$)', stmt.strip())]
    assert fields == [f[0] for f in _lib.CsmriGradArgs._fields_], fields


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    from pnp_svrg_b200 import _lib
    from pnp_svrg_b200.problems import CSMRI
    with pytest.raises(_lib.PnpError):
        CSMRI(image=np.zeros((32, 32), np.uint8) + np.arange(32, dtype=np.uint8), H=32, W=32)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pnp_svrg_b200')):
        for f in files:
            if f.endswith('.py') and re.search(r'^\s*(from|import)\s+oracle\b', open(os.path.join(dirpath, f)).read(), re.M):
                bad.append(f)
    assert not bad, bad


def test_minibatch_container():
    from pnp_svrg_b200.problems.problem import MiniBatch
    dense = np.zeros(16, dtype=int)
    dense[[3, 7]] = 1
    mb = MiniBatch(dense.reshape(4, 4), [3, 7])
    assert mb.shape == (4, 4) and list(mb.indices) == [3, 7]
    assert list(mb.ravel().indices) == [3, 7]
    assert list((np.ones((4, 4), dtype=int) * mb).indices) == [3, 7]
    assert mb[:2].indices is None
    assert int(np.asarray(mb).sum()) == 2


def test_feistel_sampler_host():
    from pnp_svrg_b200.engine import feistel_sample
    for n, c in [(1, 1), (5, 5), (1234, 200), (19575, 1000), (70000, 70000)]:
        a = feistel_sample(n, c, seed=9, counter=3)
        assert a.shape == (c,) and a.min() >= 0 and a.max() < n and len(np.unique(a)) == c
    a = feistel_sample(19575, 1000, 9, 3)
    assert not np.array_equal(a, feistel_sample(19575, 1000, 9, 4))
    assert np.array_equal(a, feistel_sample(19575, 1000, 9, 3))
    # roughly uniform: mean position near n/2
    m = np.mean([feistel_sample(10000, 500, 1, k).mean() for k in range(40)])
    assert abs(m - 5000) < 150


def test_budget_and_stop_rule():
    from pnp_svrg_b200.engine import Budget, stop_rule
    b = Budget(1e9, 3)
    assert b.alive() and b.left() == 3
    b.calls = 3
    assert not b.alive()
    assert not Budget(0.0, None).alive()
    assert stop_rule(12.34, 12.34, True, False) and not stop_rule(12.34, 12.35, True, False)
    assert stop_rule(1.0, -0.01, False, True) and not stop_rule(1.0, -0.01, False, False)


def test_reference_names_alias():
    import sys
    import pnp_svrg_b200
    names = ('problems', 'denoisers', 'algorithms', 'Utilities', 'hyperopt', 'hyperopt.hp', 'hyperopt.pyll')
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pnp_svrg_b200.install_as_reference(hyperopt=True)
        # the import block of the sweep scripts (script_diff_sampratio_set12.py:1-5) and of the notebooks
        from hyperopt import fmin, tpe, hp, Trials          # noqa: F401
        from hyperopt.hp import quniform
        from hyperopt.pyll import scope
        from Utilities import display_results               # noqa: F401
        assert scope.int(quniform('T2', 1, 100, q=1)).as_int and hp.uniform('eta', 0, 100).label == 'eta'
        from algorithms import pnp_svrg, tune_pnp_svrg      # noqa: F401
        from denoisers import TVDenoiser                    # noqa: F401
        from problems import CSMRI, Deblur, PhaseRetrieval  # noqa: F401
        from problems.CSMRI import CSMRI as C2
        assert C2 is CSMRI
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_integration_stub_mirrors_the_struct():
    """INTEGRATION.md shows the ctypes stub a reference maintainer would add: its GradArgs must list the
    fields of pnp_csmri_grad_args in order (a missing trailing field would make the library read garbage)."""
    import os
    import re
    from pnp_svrg_b200 import _lib
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'INTEGRATION.md')).read()
    block = text[text.index('class GradArgs(C.Structure)'):text.index('lib.pnp_init.restype')]
    names = re.findall(r"\('(\w+)',", block)
    assert names == [f[0] for f in _lib.CsmriGradArgs._fields_]


def test_feistel_is_a_permutation_and_the_c_twin_agrees():
    """The unbalanced Feistel network (domain 2^hb x ceil(n / 2^hb) + cycle walking) must be a bijection of [0, n)
    for every n -- powers of two, one above / below them, primes, tiny -- and the vectorised C sampler of the host
    path (pnp_sample_indices_host) must reproduce the NumPy twin bit for bit, threads or not, gather or not."""
    from pnp_svrg_b200 import _lib
    from pnp_svrg_b200.engine import feistel_sample
    lib = _lib.load()
    for n in (1, 2, 3, 4, 5, 15, 16, 17, 63, 64, 65, 97, 1000, 1023, 1024, 1025, 4093, 19661, 65536, 65537):
        ref = feistel_sample(n, n, seed=5, counter=n)
        assert np.array_equal(np.sort(ref), np.arange(n)), n
        out = np.empty(n, dtype=np.int32)
        assert lib.pnp_sample_indices_host(out.ctypes.data, n, n, 5, n, 0, 1, None) == 0
        assert np.array_equal(out, ref), n
    n, c = 1258000, 100000
    sup = (np.arange(n, dtype=np.int64) * 3 + 1).astype(np.int32)
    ref = sup[feistel_sample(n, c, 123, 7, 0)]
    for threads in (1, 4):
        out = np.empty(c, dtype=np.int32)
        assert lib.pnp_sample_indices_host(out.ctypes.data, n, c, 123, 7, 0, threads, sup.ctypes.data) == 0
        assert np.array_equal(out, ref)
    assert len(np.unique(ref)) == c
    # odd per-thread ranges (block starts that are not multiples of the 32-index vector group) and a large domain
    for n, c, threads in ((16777219, 70001, 3), (4194304, 100000, 7), (33, 33, 1)):
        ref = feistel_sample(n, c, 77, 11, 2)
        out = np.empty(c, dtype=np.int32)
        assert lib.pnp_sample_indices_host(out.ctypes.data, n, c, 77, 11, 2, threads, None) == 0
        assert np.array_equal(out, ref), (n, c, threads)


def test_host_draw_ring_sequence_and_buffer_reuse():
    """HostDrawRing (the native look-ahead draws of mb_source='host', pnp_host_draws_*): draw c lands in buffer c % R
    and equals the NumPy twin gathered through the support, whatever the number of draws in flight; entries behind
    the indices belong to the consumer; the handle can be closed with draws in flight."""
    from pnp_svrg_b200 import _lib
    from pnp_svrg_b200.engine import HostDrawRing, feistel_sample
    lib = _lib.load()
    n, B = 19661, 1000
    sup = (np.arange(n, dtype=np.int64) * 2 + 5).astype(np.int32)
    for ahead, R in ((1, 3), (3, 7), (8, 12)):
        bufs = [np.full(B + 1, -7, dtype=np.int32) for _ in range(R)]
        ring = HostDrawRing(lib, n, B, 9, sup, bufs, ahead)
        assert ring.n_extra == 1
        for c in range(3 * R + 2):
            slot = ring.next()
            assert slot == c % R and ring.drawn == c + 1
            assert np.array_equal(bufs[slot][:B], sup[feistel_sample(n, B, 9, c)]), (ahead, c)
            assert bufs[slot][B] == -7
        ring.close()
        ring.close()                                         # idempotent
    # no support: positions in [0, n); a large draw; closing right after creation (draws in flight)
    n, B = 1258000, 100000
    bufs = [np.empty(B, dtype=np.int32) for _ in range(6)]
    ring = HostDrawRing(lib, n, B, 1, None, bufs, 4)
    for c in range(9):
        assert np.array_equal(bufs[ring.next()], feistel_sample(n, B, 1, c))
    ring.close()
    HostDrawRing(lib, n, B, 1, None, bufs, 4).close()
    del ring
    # too few buffers / wrong dtype are refused
    with pytest.raises(_lib.PnpError):
        HostDrawRing(lib, n, B, 1, None, bufs[:3], 2)
    with pytest.raises(ValueError):
        HostDrawRing(lib, n, B, 1, None, [np.empty(B, dtype=np.int64) for _ in range(4)], 2)
    with pytest.raises(ValueError):
        HostDrawRing(lib, n, B, 1, np.zeros(n, dtype=np.int64), bufs, 2)
    with pytest.raises(ValueError):                          # support on the host AND on the device
        HostDrawRing(lib, 100, 10, 1, np.arange(100, dtype=np.int32), [np.empty(10, dtype=np.int32) for _ in range(4)], 2,
                     support_dev_ptr=4096)


def test_line_layout_conversions_are_value_exact():
    """device.to_lines / from_lines (host float64 raveled image <-> float32 [W][H] line layout): the torch cast +
    transpose must give exactly what the plain NumPy transpose + cast gives, for vectors, images, tensors and
    read-only inputs, square or not."""
    import torch
    from pnp_svrg_b200 import device as D
    for H, W in ((64, 64), (32, 128), (256, 64)):
        z = np.random.default_rng(H).uniform(-1, 1, H * W)
        want = torch.from_numpy(np.ascontiguousarray(z.reshape(H, W).T, dtype=np.float32))
        got = D.to_lines(z, H, W, device='cpu')
        assert got.shape == (W, H) and got.is_contiguous() and got.dtype == torch.float32 and torch.equal(got, want)
        assert torch.equal(D.to_lines(z.reshape(H, W), H, W, device='cpu'), want)
        assert torch.equal(D.to_lines(torch.from_numpy(z), H, W, device='cpu'), want)
        ro = z.copy()
        ro.flags.writeable = False
        assert torch.equal(D.to_lines(ro, H, W, device='cpu'), want)
        back = D.from_lines(got, H, W)
        assert back.dtype == np.float64 and back.shape == (H * W,)
        assert np.array_equal(back, z.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize('simd', ['scalar', 'avx2', 'avx512'])
def test_host_sampler_simd_paths_agree(simd):
    """The scalar, AVX2 and AVX-512 bodies of the host sampler are picked at run time (the best the CPU has, capped by
    PNP_HOST_SIMD): each must reproduce the NumPy twin bit for bit.  Fresh interpreter per level: the choice is
    made once per process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import numpy as np
from pnp_svrg_b200 import _lib
from pnp_svrg_b200.engine import feistel_sample
lib = _lib.load()
for n, c, thr in ((1, 1, 1), (5, 5, 1), (97, 97, 1), (1025, 1025, 1), (65537, 40000, 3), (1258000, 100000, 1), (16777219, 70001, 2)):
    ref = feistel_sample(n, c, 5, n)
    out = np.empty(c, dtype=np.int32)
    assert lib.pnp_sample_indices_host(out.ctypes.data, n, c, 5, n, 0, thr, None) == 0
    assert np.array_equal(out, ref), n
    sup = ((np.arange(n, dtype=np.int64) * 7 + 3) % 2000000011).astype(np.int32)
    assert lib.pnp_sample_indices_host(out.ctypes.data, n, c, 5, n, 0, thr, sup.ctypes.data) == 0
    assert np.array_equal(out, sup[ref]), n
print('SIMD-OK')
'''
    r = subprocess.run([sys.executable, '-c', code], cwd=root, env=dict(os.environ, PNP_HOST_SIMD=simd, PYTHONPATH=root),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'SIMD-OK' in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_pin_to_local_rank_gives_disjoint_core_slices():
    """one process per GPU on one host: the local ranks' sampler threads get disjoint slices of the allowed cores"""
    import os
    from pnp_svrg_b200.device import pin_to_local_rank
    if not hasattr(os, 'sched_setaffinity'):
        pytest.skip('no sched_setaffinity on this platform')
    before = sorted(os.sched_getaffinity(0))
    try:
        if len(before) < 4:
            assert pin_to_local_rank(0, 2) is None and sorted(os.sched_getaffinity(0)) == before
            return
        seen = []
        for r in range(2):
            os.sched_setaffinity(0, before)
            mine = pin_to_local_rank(r, 2)
            assert mine == sorted(os.sched_getaffinity(0)) and len(mine) == len(before) // 2
            seen.append(set(mine))
        assert not (seen[0] & seen[1])
        os.sched_setaffinity(0, before)
        assert pin_to_local_rank(0, 1) is None and sorted(os.sched_getaffinity(0)) == before       # single process: untouched
    finally:
        os.sched_setaffinity(0, before)
