"""Build libpnp_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'pnp_b200.cu')
HOST_SRC = os.path.join(HERE, 'csrc', 'host_sampler.cpp')     # plain C++: nvcc hands it to the host compiler
OUT = os.environ.get('PNP_LIB_OUT') or os.path.join(HERE, 'lib', 'libpnp_b200.so')
NVCC_FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-shared', '-Xcompiler', '-fPIC']


def _sources():
    d = os.path.join(HERE, 'csrc')
    inc = os.path.join(os.path.dirname(HERE), 'include')
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(inc, f) for f in os.listdir(inc)]


def build(force=False, verbose=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT):
        newest = max(os.path.getmtime(f) for f in _sources())
        if os.path.getmtime(OUT) >= newest:
            return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('PNP_NVCC_EXTRA', '').split()          # e.g. -DPNP_PHASE_TIMING for scripts/prof_phases.py
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC, HOST_SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
