"""Profile target: a few eager PnP-SVRG inner iterations of the bench workload (for ncu)."""
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
    ep = bench.Epoch(cfg, seed=0)
    eng = ep.eng
    with torch.cuda.stream(eng.stream):
        ep.snapshot_ops()
        for _ in range(a.iters):
            ep.inner_ops()
    eng.stream.synchronize()
    print('ok', eng.flush_fast() if False else '')


if __name__ == '__main__':
    main()
