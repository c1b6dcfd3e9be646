"""pnp_sarah / tune_pnp_sarah -- same signatures and result dict as the reference's algorithms/pnp_sarah.py:8-129, tune at :130-155;
the loop body runs on the GPU (see _loops.py and engine.py)."""
from ._loops import pnp_sarah
from ._tune import status_ok, tune_result

tol = 1e-5


def tune_pnp_sarah(args, problem, denoiser, tt, lr_decay=1, verbose=False, converge_check=True, diverge_check=True,
             **extra):
    """hyperopt objective wrapper: loss = PSNR(Xinit) - PSNR(z)."""
    eta, mini_batch_size, T2, dstrength = args
    denoiser.sigma_est = dstrength        # assigned but read by no denoiser, as in the reference
    result = pnp_sarah(problem=problem, denoiser=denoiser, tt=tt, eta=eta, mini_batch_size=mini_batch_size, T2=T2, verbose=verbose, lr_decay=lr_decay,
                 converge_check=converge_check, diverge_check=diverge_check, **extra)
    return tune_result(problem, result)
