"""Golden vectors for the CNN denoisers from the REFERENCE's own model classes and checkpoints
(torch CPU fp32), run in the build container only:

    python -m oracle.gen_golden_cnn

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference wrappers hard-code .cuda()
(denoisers/DeepDenoisers/utils/utils.py:15-29) and import matplotlib (MMODenoise.py:10), neither of
which is available here, so the wrapper arithmetic around ``model(x)`` is restated below line by line
(RealSN_DnCNN.py:16-40, MMODenoise.py:18-40,124-128); the networks themselves are the reference's
classes with the reference's weights.  The weights are stored in the fixture (float32) because the
checkpoints do not travel to the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, 'denoisers'))
    from PIL import Image
    img = np.array(Image.open(os.path.join(REF, 'data/Set12/01.png')).resize((64, 64))).astype(np.float64) / 255.0
    rng = np.random.default_rng(0)
    noisy = img + (15 / 255.0) * rng.standard_normal(img.shape)
    noisy[:, :8] *= 1.3                      # leave [0, 1] so the clamps / min-max paths matter

    # ---- DnCNN-17 through the RealSN_DnCNNDenoiser wrapper arithmetic --------------------------
    # import the model file directly: denoisers/__init__.py pulls in bm3d / skimage
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_dncnn_models', os.path.join(REF, 'denoisers/DeepDenoisers/model/models.py'))
    ref_models = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_models)
    DnCNN = ref_models.DnCNN
    sigma = 15
    sd = torch.load(os.path.join(REF, 'denoisers/DeepDenoisers/Pretrained_models/DnCNN_noise15.pth'), map_location='cpu')
    net = torch.nn.DataParallel(DnCNN(channels=1, num_of_layers=17))
    net.load_state_dict(sd)
    net.eval()
    m, n = noisy.shape
    xtilde = np.copy(noisy)
    mintmp, maxtmp = np.min(xtilde), np.max(xtilde)
    xtilde = (xtilde - mintmp) / (maxtmp - mintmp)
    scale_range = 1.0 + sigma / 255.0 / 2.0
    scale_shift = (1 - scale_range) / 2.0
    xtilde = xtilde * scale_range + scale_shift
    with torch.no_grad():
        r = net.module(torch.from_numpy(np.reshape(xtilde, (1, 1, m, n))).type(torch.FloatTensor)).numpy()
    x = xtilde - np.reshape(r, (m, n))
    x = (x - scale_shift) / scale_range
    x = x * (maxtmp - mintmp) + mintmp
    arrays = {'sd__' + k: v.numpy().astype(np.float32) for k, v in sd.items() if v.ndim > 0}
    np.savez_compressed(os.path.join(OUT, 'ref_cnn_dncnn15.npz'),
                        meta=json.dumps(dict(name='cnn_dncnn15', model_type='DnCNN', sigma=sigma,
                                             reference_files=['denoisers/RealSN_DnCNN.py:16-40',
                                                              'denoisers/DeepDenoisers/model/models.py:5-22'])),
                        noisy=noisy, denoised=x, **arrays)
    print('dncnn15: in [%.3f, %.3f] out [%.3f, %.3f]' % (noisy.min(), noisy.max(), x.min(), x.max()))

    # ---- MMO DnCNN_nobn-20 through apply_model / MMODenoiser.denoise ----------------------------
    model = torch.load(os.path.join(REF, 'denoisers/checkpoints/pretrained/DnCNN_nobn_nch_1_nlev_0.01.pth'),
                       map_location='cpu', weights_only=False)
    mod = model.module.eval()
    x_cur = np.moveaxis(noisy, -1, 0)                               # MMODenoise.py:126 (a transpose for 2-D)
    imgn = torch.from_numpy(np.ascontiguousarray(x_cur)).unsqueeze(0).unsqueeze(0).type(torch.FloatTensor)
    with torch.no_grad():
        imgn.clamp_(0, 1)
        out_net = mod(imgn)
        out_net.clamp_(0, 1)
    tmp = out_net[0, 0].numpy()
    y = np.clip(np.moveaxis(tmp, 0, -1), 0., 1.)
    arrays = {'sd__' + k: v.detach().numpy().astype(np.float32) for k, v in mod.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, 'ref_cnn_mmo_nobn.npz'),
                        meta=json.dumps(dict(name='cnn_mmo_nobn', depth=int(mod.depth), slope=0.01,
                                             reference_files=['denoisers/MMODenoise.py:18-40,73-103,124-128'])),
                        noisy=noisy, denoised=y.astype(np.float64), **arrays)
    print('mmo: out [%.3f, %.3f]' % (y.min(), y.max()))
    for f in ('ref_cnn_dncnn15.npz', 'ref_cnn_mmo_nobn.npz'):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
