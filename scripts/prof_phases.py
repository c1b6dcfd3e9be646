"""Phase times inside k_update_prox (CTA 0, %globaltimer).  Needs a build with -DPNP_PHASE_TIMING:
    PNP_NVCC_EXTRA=-DPNP_PHASE_TIMING python -m pnp_svrg_b200.build --force && python scripts/prof_phases.py"""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
ba = argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1)
ep = bench.Epoch(bench.workload(ba), seed=0)
eng = ep.eng
with torch.cuda.stream(eng.stream):
    ep.snapshot_ops()
    for _ in range(5):
        ep.inner_ops()
        eng.stream.synchronize()
        eng.lib.pnp_debug_set(2, 0)
