"""pnp_saga / tune_pnp_saga -- same signatures and result dict as the reference's algorithms/pnp_saga.py:8-102, tune at :104-129;
the loop body runs on the GPU (see _loops.py and engine.py)."""
from ._loops import pnp_saga
from ._tune import status_ok, tune_result

tol = 1e-5


def tune_pnp_saga(args, problem, denoiser, tt, lr_decay=1, verbose=False, converge_check=True, diverge_check=True,
             **extra):
    """hyperopt objective wrapper: loss = PSNR(Xinit) - PSNR(z)."""
    eta, mini_batch_size, dstrength, hist_size = args
    denoiser.sigma_est = dstrength        # assigned but read by no denoiser, as in the reference
    result = pnp_saga(problem=problem, denoiser=denoiser, tt=tt, eta=eta, mini_batch_size=mini_batch_size, hist_size=hist_size, verbose=verbose, lr_decay=lr_decay,
                 converge_check=converge_check, diverge_check=diverge_check, **extra)
    return tune_result(problem, result)
