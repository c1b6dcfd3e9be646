"""Profile target: a few eager PnP-SVRG inner iterations of the bench workload (for ncu): the launches of
SvrgRun.fast_ops() -- line pass (+ in-pass minibatch selection), column pass, single-launch tail."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse

import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--iters', type=int, default=3)
    a = ap.parse_args()
    ba = argparse.Namespace(size=a.size, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
    cfg = bench.workload(ba)
    prob, run = bench.make_run(cfg, seed=0)
    eng = run.eng
    with torch.cuda.stream(eng.stream):
        eng.set_step(cfg['eta'])
        run.snapshot()
        for _ in range(a.iters):
            run.fast_ops()
    eng.stream.synchronize()
    print('ok')


if __name__ == '__main__':
    main()
