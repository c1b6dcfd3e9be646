"""CPU oracle for the PnP iteration hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a float64 NumPy restatement of the reference
(vmonardo/pnp-svrg) and of the third-party numerics it calls.  It exists so the
CUDA path can be checked; it is never the thing that is shipped or measured.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``pnp_svrg_b200``) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * problems / algorithms restatements: PINNED bit-exactly against the reference's own
    NumPy code executed in the build container (``oracle/gen_golden.py`` imports
    ``/root/reference`` through ``oracle/refshim.py``; outputs in ``tests/golden``),
    plus the one deterministic known answer the reference ships (Deblur sigma, notebook
    cell 4 of create_paper_figures_deblur.ipynb).
  * scikit-image 0.18.2 / PyWavelets 1.1.1 / pylops 1.14.0 restatements
    (``skimage_port``, ``pylops_port``): PARITY UNPINNED -- those libraries are not
    installed in the build image and the reference holds no tests or golden vectors for
    them.  They are restated from the published algorithms and checked against analytic
    identities only (Haar perfect reconstruction, Parseval, the documented
    ``pywt.dwt([1,2,3,4],'db1')`` value, adjointness of the bilinear operator).
"""
