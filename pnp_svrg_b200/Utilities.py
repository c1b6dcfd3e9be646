"""display_results -- the step after a reconstruction in the reference's drivers and notebooks (Utilities.py:5-64):
output image, PSNR-over-time plot, the one-line metrics print and ``output.csv``.

Same signature, same file names (``output.eps``, ``psnr_over_time.eps``, ``output.csv``), same CSV header and
rounding (Utilities.py:55-63).  Two deliberate differences, both on the host and neither touching a result:
  * matplotlib is optional here (absent from this image): without it the figures are skipped, the print and the
    CSV are still produced and the function returns None instead of the PSNR axes;
  * the reference's print formats ``gradient_time`` into the "Denoising Time" field as well (``{3}`` twice,
    Utilities.py:51-53); the denoising time is printed here.  The CSV was already correct in the reference.
"""
import csv
import os

import numpy as np


def _pyplot():
    try:
        import matplotlib
        if not os.environ.get('DISPLAY') and not os.environ.get('MPLBACKEND'):
            matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        return plt
    except ImportError:
        return None


def metrics_row(output_dict):
    """[Output PSNR, Change in PSNR, Gradient Time, Denoising Time] rounded as Utilities.py:59-62."""
    psnr = output_dict['psnr_per_iter']
    return [np.around(psnr[-1], decimals=1), np.around(psnr[-1] - psnr[0], decimals=2),
            np.around(output_dict['gradient_time'], decimals=2), np.around(output_dict['denoise_time'], decimals=2)]


CSV_HEADER = ['Output PSNR', 'Change in PSNR', 'Gradient Time', 'Denoising Time']


def display_results(problem, output_dict, save_results=False, save_dir='figures/', show_figs=False):
    base = None
    if save_results:
        base = (problem.prob_dir + output_dict['algo_name'] + '/') if getattr(problem, 'prob_dir', None) else save_dir
        os.makedirs(base, exist_ok=True)

    t_arr = np.asarray(output_dict['time_per_iter'], dtype=np.float64)
    psnr = np.asarray(output_dict['psnr_per_iter'], dtype=np.float64)
    plt = _pyplot()
    psnr_ax = None
    if plt is not None:
        img = np.asarray(output_dict['z']).reshape(problem.H, problem.W)
        out_fig = plt.figure(figsize=(6, 6))
        plt.imshow(img, cmap=getattr(problem, 'color_map', 'gray'), vmin=0, vmax=1)
        plt.title('Output Image')
        plt.xticks([])
        plt.yticks([])
        if save_results:
            out_fig.savefig(base + 'output.eps', transparent=True, bbox_inches='tight', pad_inches=0)
        if show_figs:
            plt.show()

        psnr_fig = plt.figure(figsize=(6, 6))
        psnr_ax = psnr_fig.add_subplot(1, 1, 1)
        psnr_ax.plot(np.cumsum(t_arr), psnr, "b", linewidth=3, label=str(output_dict['algo_name']))
        psnr_ax.plot(np.cumsum(t_arr)[::30], psnr[::30], "b*", markersize=10)
        psnr_ax.set(xlabel='time (s)', ylabel='PSNR (dB)')
        psnr_ax.legend()
        psnr_ax.grid()
        psnr_fig.tight_layout()
        if show_figs:
            plt.show()
        if save_results:
            psnr_fig.savefig(base + 'psnr_over_time.eps', transparent=True, bbox_inches='tight', pad_inches=0)

    print('Output PSNR: {0:3.1f}\tChange in PSNR: {1:3.2f}\tGradient Time: {2:3.2f}\tDenoising Time: {3:3.2f}'.format(
        psnr[-1], psnr[-1] - psnr[0], output_dict['gradient_time'], output_dict['denoise_time']))
    if save_results:
        with open(base + 'output.csv', 'w') as f:
            w = csv.writer(f, delimiter=',')
            w.writerow(CSV_HEADER)
            w.writerow(metrics_row(output_dict))
    return psnr_ax
