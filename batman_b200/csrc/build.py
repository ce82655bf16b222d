"""Build libbatman_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python batman_b200/csrc/build.py [--force]

Objects are compiled in parallel (one nvcc per .cu) and linked into
batman_b200/libbatman_b200.so, which travels to the GPU box with the snapshot.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, 'libbatman_b200.so')
BUILD = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']
# the Saltelli kernels are 50-60 template instantiations per file (10 and 5 minutes of ptxas on one core): their
# device code is compiled on four threads; every other file keeps the single-threaded code generation it was tuned with
SPLIT_COMPILE = {'sobol.cu': ['-split-compile', '4'], 'sobol_static.cu': ['-split-compile', '4']}


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith('.cu'))


def newest_dep():
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh', '.h'))]
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'batman_b200.h'))
    return max(os.path.getmtime(d) for d in deps)


def header_deps(src, seen=None):
    """Quoted includes of `src`, followed recursively (the kernels only include files of this tree)."""
    seen = set() if seen is None else seen
    with open(src) as f:
        text = f.read()
    for inc in re.findall(r'#include\s+"([^"]+)"', text):
        path = os.path.normpath(os.path.join(os.path.dirname(src), inc))
        if os.path.exists(path) and path not in seen:
            seen.add(path)
            header_deps(path, seen)
    return seen


def compile_one(src, force=False):
    obj = os.path.join(BUILD, src[:-3] + '.o')
    log = os.path.join(BUILD, src[:-3] + '.ptxas.log')
    path = os.path.join(HERE, src)
    newest = max(os.path.getmtime(d) for d in [path, os.path.abspath(__file__)] + sorted(header_deps(path)))
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    r = subprocess.run([NVCC] + FLAGS + SPLIT_COMPILE.get(src, []) + ['-c', os.path.join(HERE, src), '-o', obj],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, 'w') as f:
        f.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stdout[-6000:]))
    return obj


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest_dep():
        return OUT
    os.makedirs(BUILD, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda f: compile_one(f, force), sources()))
    r = subprocess.run([NVCC, '-shared', '-o', OUT] + objs + ['-lcudart'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout[-4000:])
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
