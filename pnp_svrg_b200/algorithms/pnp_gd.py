"""pnp_gd / tune_pnp_gd -- same signatures and result dict as the reference's algorithms/pnp_gd.py:8-84, tune at :86-110;
the loop body runs on the GPU (see _loops.py and engine.py)."""
from ._loops import pnp_gd
from ._tune import status_ok, tune_result

tol = 1e-5


def tune_pnp_gd(args, problem, denoiser, tt, lr_decay=1, verbose=False, converge_check=True, diverge_check=True,
             **extra):
    """hyperopt objective wrapper: loss = PSNR(Xinit) - PSNR(z)."""
    eta, dstrength = args
    denoiser.sigma_est = dstrength        # assigned but read by no denoiser, as in the reference
    result = pnp_gd(problem=problem, denoiser=denoiser, tt=tt, eta=eta, verbose=verbose, lr_decay=lr_decay,
                 converge_check=converge_check, diverge_check=diverge_check, **extra)
    return tune_result(problem, result)
