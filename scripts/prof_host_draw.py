"""Host sampler throughput on this box: one C call per draw (single thread) and N concurrent Python threads."""
import os, sys, time, json
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pnp_svrg_b200 import _lib
lib = _lib.load()
n, c = 1258000, 100000
sup = np.sort(np.random.default_rng(0).choice(4194304, n, replace=False)).astype(np.int32)
res = {'cpu': [l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0],
       'cpus': os.cpu_count(), 'affinity': len(os.sched_getaffinity(0))}
out = np.empty(c, dtype=np.int32)
t0 = time.time()
for k in range(50):
    lib.pnp_sample_indices_host(out.ctypes.data, n, c, 123, k, 0, 1, sup.ctypes.data)
res['ms_per_draw_1thread'] = (time.time() - t0) / 50 * 1e3
for workers in (2, 4, 8, 12):
    bufs = [np.empty(c, dtype=np.int32) for _ in range(workers)]
    def job(k):
        lib.pnp_sample_indices_host(bufs[k % workers].ctypes.data, n, c, 123, k, 0, 1, sup.ctypes.data)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(job, range(workers * 4)))
        t0 = time.time()
        list(ex.map(job, range(workers * 40)))
        res['draws_per_s_%dworkers' % workers] = workers * 40 / (time.time() - t0)
print(json.dumps(res))
