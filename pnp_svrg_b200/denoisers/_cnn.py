"""Weight packing and device forward shared by the CNN denoisers (csrc/cnn_fp32.cuh, cnn_tc.cuh).

The checkpoints are the reference's own files: state dicts with keys ``[module.]dncnn.N.*``
(denoisers/DeepDenoisers/utils/utils.py:10-33) or a pickled DataParallel(simple_CNN)
(denoisers/MMODenoise.py:42-71).  RealSN checkpoints carry ``weight_orig`` / ``weight_u`` next to
``weight``; at eval only the stored ``weight`` buffer is used (Spectral_Normalize_chen.py:82-89).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, device as D


def _strip(sd):
    return {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}


def layers_from_dncnn_state_dict(sd, eps=1e-5):
    """-> list of dict(w=(co,ci,3,3) float64, scale, shift, slope) for a DnCNN / SimpleCNN Sequential."""
    sd = {k: (v.detach().cpu().double().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64))
          for k, v in _strip(sd).items()}
    conv_ids = sorted(int(k.split('.')[1]) for k in sd if k.startswith('dncnn.') and k.endswith('.weight')
                      and sd[k].ndim == 4)
    layers = []
    for n, i in enumerate(conv_ids):
        lay = dict(w=sd['dncnn.%d.weight' % i], scale=None, shift=None, slope=0.0)
        bn = 'dncnn.%d.running_var' % (i + 1)
        if bn in sd:       # eval-mode BatchNorm folded into a per-channel affine map
            g, b = sd['dncnn.%d.weight' % (i + 1)], sd['dncnn.%d.bias' % (i + 1)]
            m, v = sd['dncnn.%d.running_mean' % (i + 1)], sd[bn]
            lay['scale'] = g / np.sqrt(v + eps)
            lay['shift'] = b - m * lay['scale']
        if 'dncnn.%d.bias' % i in sd:
            lay['shift'] = sd['dncnn.%d.bias' % i] if lay['shift'] is None else lay['shift'] + sd['dncnn.%d.bias' % i] * lay['scale']
        layers.append(lay)
    return layers


def layers_from_simple_cnn(module):
    """MMO simple_CNN (denoisers/MMODenoise.py:73-103): conv + bias + LeakyReLU(0.01), no BN."""
    sd = {k: v.detach().cpu().double().numpy() for k, v in module.state_dict().items()}
    names = ['in_conv'] + ['conv_list.%d' % i for i in range(module.depth - 2)] + ['out_conv']
    slope = float(getattr(module.nl_list[0], 'negative_slope', 0.0))
    return [dict(w=sd[n + '.weight'], scale=None, shift=sd[n + '.bias'], slope=slope) for n in names]


class PackedNet:
    """Device copy of a 3x3 conv stack in the kernels' layout."""

    def __init__(self, layers, mode, swap_spatial, device, range_=1.0, shift_in=0.0):
        if len(layers) < 2 or len(layers) > _lib.CNN_MAX_LAYERS:
            raise ValueError('unsupported depth %d' % len(layers))
        self.device = device
        self.keep = []
        net = _lib.CnnNet()
        net.n_layers = len(layers)
        for i, lay in enumerate(layers):
            w = np.asarray(lay['w'], dtype=np.float64)
            co, ci = w.shape[:2]
            ok = (ci == 1 and co == 64) if i == 0 else ((ci == 64 and co == 1) if i == len(layers) - 1 else (ci == 64 and co == 64))
            if not ok or w.shape[2:] != (3, 3):
                raise NotImplementedError('only 1->64->...->64->1 stacks of 3x3 convolutions are built (layer %d is %s)'
                                          % (i, w.shape))
            if swap_spatial:       # the device image is the transpose: swap the two kernel axes
                w = w.transpose(0, 1, 3, 2)
            # [co][ci][dl][dp] -> [tap][ci][co]
            packed = np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(9, ci, co), dtype=np.float32)
            net.w[i] = self._up(packed)
            if 0 < i < len(layers) - 1:
                # tensor-core operand: rows n = (dp, co), columns k = (dl, ci), bf16, BatchNorm scale folded in
                ws = w if lay['scale'] is None else w * np.asarray(lay['scale'], dtype=np.float64)[:, None, None, None]
                w2 = ws.transpose(3, 0, 2, 1).reshape(3 * 64, 3 * 64)       # [dp][co][dl][ci]
            elif i == len(layers) - 1:
                w2 = np.zeros((16, 3 * 64))                                 # rows dp, padded to a legal UMMA N
                w2[:3] = w[0].transpose(2, 1, 0).reshape(3, 3 * 64)         # [dp][dl][ci]
            if i > 0:
                w32 = torch.from_numpy(np.ascontiguousarray(w2, dtype=np.float32)).to(self.device)
                t = w32.to(torch.bfloat16).contiguous()
                # error-compensated mode (precision 2): the part of the fp32 weight that bf16 dropped, as a second bf16 operand
                t_lo = (w32 - t.to(torch.float32)).to(torch.bfloat16).contiguous()
                self.keep += [t, t_lo]
                net.w_tc[i] = t.data_ptr()
                net.w_tc_lo[i] = t_lo.data_ptr()
            last = i == len(layers) - 1
            if not last:
                net.scale[i] = self._up(lay['scale']) if lay['scale'] is not None else None
                net.shift[i] = self._up(lay['shift']) if lay['shift'] is not None else None
                net.slope[i] = float(lay['slope'])
            else:
                net.last_bias = float(np.ravel(lay['shift'])[0]) if lay['shift'] is not None else 0.0
        net.mode = int(mode)
        net.range = float(range_)
        net.shift_in = float(shift_in)
        self.net = net
        self.stats = torch.zeros(2, dtype=torch.int32, device=device)
        self._act = None

    def _up(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        self.keep.append(t)
        return t.data_ptr()

    def forward(self, img, out, PH, PW, xrec=None, mse_log=None, slot=None, precision=0):
        if precision >= 1:
            # one zero pad pixel per line, zeroed once; precision 2 keeps a hi and a lo plane per buffer
            n, dt = PH * (PW + 1) * 64 * (2 if precision == 2 else 1), torch.bfloat16
        else:
            n, dt = PH * PW * 64, torch.float32
        if self._act is None or self._act[0].numel() != n or self._act[0].dtype != dt:
            self._act = (torch.zeros(n, dtype=dt, device=self.device), torch.zeros(n, dtype=dt, device=self.device))
        _lib.check(_lib.load().pnp_cnn_forward(C.byref(self.net), D.ptr(img), D.ptr(out), PH, PW, D.ptr(self._act[0]),
                                               D.ptr(self._act[1]), D.ptr(self.stats), D.ptr(xrec), D.ptr(mse_log),
                                               D.ptr(slot), int(precision), D.stream()))
