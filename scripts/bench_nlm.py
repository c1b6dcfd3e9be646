"""NLM prox (config 2: patch 4 -> 5, distance 5) at 256^2 and 512^2: time per launch (CUDA events) and the difference
between the specialised kernel (k_nlm5) and the generic one (PNP_NLM_GENERIC=1 in a second process: the switch is read
once).  Usage: python scripts/bench_nlm.py [out.json]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def run():
    import numpy as np
    import torch
    from conftest import synth_image
    from pnp_svrg_b200 import device as D
    from pnp_svrg_b200.denoisers import NLMDenoiser
    from pnp_svrg_b200.engine import ProxCtx
    dev = D.require_cuda()
    res = {}
    for H in (256, 512):
        rng = np.random.default_rng(H)
        img = synth_image(H, H, 0).astype(np.float64) / 255 + 0.05 * rng.standard_normal((H, H))
        z = D.to_lines(img, H, H, dev)
        o = torch.empty_like(z)
        nlm = NLMDenoiser()
        for _ in range(3):
            nlm._dev_denoise(ProxCtx(z, o, H, H, sigma_est=0.05))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            nlm._dev_denoise(ProxCtx(z, o, H, H, sigma_est=0.05))
        e1.record()
        torch.cuda.synchronize()
        res[str(H)] = {'us': e0.elapsed_time(e1) * 1e3 / n, 'sum': float(o.double().sum()), 'out': o.cpu().numpy()}
    return res


if __name__ == '__main__':
    if os.environ.get('NLM_CHILD'):
        import numpy as np
        r = run()
        np.savez(os.environ['NLM_CHILD'], **{k: v['out'] for k, v in r.items()})
        print(json.dumps({k: {'us': v['us'], 'sum': v['sum']} for k, v in r.items()}))
        sys.exit(0)
    import numpy as np
    out = {}
    for name, env in (('k_nlm5', {}), ('k_nlm_generic', {'PNP_NLM_GENERIC': '1'})):
        e = dict(os.environ, NLM_CHILD='/tmp/nlm_%s.npz' % name, **env)
        txt = subprocess.run([sys.executable, __file__], env=e, capture_output=True, text=True, check=True).stdout
        out[name] = json.loads(txt.strip().splitlines()[-1])
    a, b = np.load('/tmp/nlm_k_nlm5.npz'), np.load('/tmp/nlm_k_nlm_generic.npz')
    out['max_abs_diff'] = {k: float(np.abs(a[k] - b[k]).max()) for k in a.files}
    out['rel_l2_diff'] = {k: float(np.linalg.norm(a[k] - b[k]) / np.linalg.norm(b[k])) for k in a.files}
    s = json.dumps(out)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(s + '\n')
