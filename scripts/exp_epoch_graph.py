"""EXPERIMENT (not yet run on a GPU when committed): one CUDA graph per SVRG epoch (snapshot gradient + T2 inner
iterations) against one graph per inner iteration (what bench.py's Epoch.step and the API loops replay today), on the
bench workload with the device sampler.  Prints us per inner iteration for both and checks that the iterates agree
bit for bit (same kernels, same order, same draws).  Run: python scripts/exp_epoch_graph.py [epochs]"""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = bench.workload(argparse.Namespace(size=2048, batch_size=0, sample_prob=0.3, eta=0.0, T2=10, gpus=1))
T2 = cfg['T2']
ep = bench.Epoch(cfg, seed=0)
eng = ep.eng
with torch.cuda.stream(eng.stream):
    ep.snapshot_ops()
    ep.inner_ops()                                   # eager warm iteration before any capture
eng.stream.synchronize()
ep.capture()                                         # graph of one inner iteration


def epoch_ops():
    ep.snapshot_ops()
    for _ in range(T2):
        ep.inner_ops()


t_before = ep.den.t
g_epoch = eng.capture(epoch_ops)                     # graph of a whole epoch
ep.den.t = t_before                                  # the capture ran inner_ops T2 times on the host side only


def run(step_fn, n):
    ep.reset()
    with torch.cuda.stream(eng.stream):
        eng.counters.zero_()
    eng.stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for _ in range(n):
        step_fn()
    e1.record(eng.stream)
    eng.stream.synchronize()
    eng.slot_host = 0
    with torch.cuda.stream(eng.stream):
        eng._reset_logs()
    return e0.elapsed_time(e1) * 1e3 / (n * T2), eng.z.clone()


def step_iter_graphs():
    with torch.cuda.stream(eng.stream):
        ep.snapshot_ops()
        for _ in range(T2):
            eng.replay(ep.graph)


def step_epoch_graph():
    eng.replay(g_epoch)


n = min(epochs, 4096 // T2 - 1)                      # the PSNR / sigma log holds 4096 slots between flushes
run(step_iter_graphs, 3), run(step_epoch_graph, 3)   # warm-up
us_a, z_a = run(step_iter_graphs, n)
us_b, z_b = run(step_epoch_graph, n)
print(json.dumps({'epochs': n, 'T2': T2, 'us_per_inner_iteration': {'graph_per_iteration': us_a, 'graph_per_epoch': us_b},
                  'iterates_identical': bool(torch.equal(z_a, z_b))}))
