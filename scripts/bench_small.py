"""256x256: the one-cluster-per-image kernel (csrc/small.cuh) against the three-pass path it replaces (PNP_SMALL=0):
single-image it/s (device-resident epochs + the public call) and the 840-job sweep, same process, same box.
usage: python scripts/bench_small.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import bench_sections as BS
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    out = {}
    for tag, flag in (('three_pass', '0'), ('cluster', '1')):
        os.environ['PNP_SMALL'] = flag
        out[tag] = {'small': BS.small(0, 1, dev), 'sweep': BS.sweep(0, 1, dev, with_cpu=False),
                    'sweep_serial_batches': BS.sweep(0, 1, dev, with_cpu=False, pipelined=False),
                    'sweep_batch120': BS.sweep(0, 1, dev, with_cpu=False, batch=120)}
    os.environ.pop('PNP_SMALL', None)
    a, b = out['three_pass'], out['cluster']
    out['speedup'] = {'single_image': b['small']['value'] / a['small']['value'],
                      'single_image_e2e': b['small']['e2e_value'] / a['small']['e2e_value'],
                      'sweep': b['sweep']['value'] / a['sweep']['value']}
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(s)


if __name__ == '__main__':
    main()
