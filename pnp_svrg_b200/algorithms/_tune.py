"""Shared tail of the tune_pnp_* wrappers (reference algorithms/pnp_svrg.py:107-132 and siblings)."""


def status_ok():
    try:
        from hyperopt import STATUS_OK
        return STATUS_OK
    except ImportError:          # hyperopt is optional; its constant is the string 'ok'
        return 'ok'


def tune_result(problem, result):
    return {
        'loss': (problem.PSNR(problem.Xinit) - problem.PSNR(result['z'])),
        'status': status_ok(),
        'algo_name': result['algo_name'],
        'z': result['z'],
        'time_per_iter': result['time_per_iter'],
        'psnr_per_iter': result['psnr_per_iter'],
        'gradient_time': result['gradient_time'],
        'denoise_time': result['denoise_time'],
    }
