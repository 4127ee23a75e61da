"""Build recipe for the C-ABI CUDA library (sm_100a only) - ``python -m htd_b200.build``.

Compiles every ``csrc/*.cu`` with nvcc into ``htd_b200/_lib/libhtd_b200.so`` (in-tree, so it
travels to the GPU box; ``*.so`` is git-ignored).  nvcc cross-compiles without a GPU.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_lib')
LIB = os.path.join(OUT_DIR, 'libhtd_b200.so')
# Same sources compiled with -DHTD_DEBUG_HOOKS: kernel-variant selection (htd_debug_set_bwd_variant,
# HTD_FWD_KERNEL / HTD_BWD_KERNEL / HTD_DENSE_* environment switches), the backward trace and the
# in-kernel experiment switches.  The product library has none of them; tests and tools that
# compare kernel variants load this one (htd_b200._lib.hooks_library()).
LIB_HOOKS = os.path.join(OUT_DIR, 'libhtd_b200_hooks.so')
HOOK_SOURCES = ('roi_align.cu', 'dense_gemm.cu')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.isfile(LIB) or not os.path.isfile(LIB_HOOKS):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_HOOKS))
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) \
        + [os.path.join(os.path.dirname(HERE), 'include', 'htd_b200.h'), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Returns the path of the built product library; rebuilds only when a source is newer.  All
    translation units (and the hook variants of two of them) compile in parallel."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(OUT_DIR, exist_ok=True)
    objs, hook_objs = [], []
    procs = []

    def compile_(src, obj, extra):
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))

    for src in sources():
        base = os.path.basename(src)
        obj = os.path.join(OUT_DIR, base[:-3] + '.o')
        compile_(src, obj, [])
        objs.append(obj)
        if base in HOOK_SOURCES:
            hobj = os.path.join(OUT_DIR, base[:-3] + '.hooks.o')
            compile_(src, hobj, ['-DHTD_DEBUG_HOOKS=1'])
            hook_objs.append(hobj)
        else:
            hook_objs.append(obj)
    for cmd, pr in procs:
        out = pr.communicate()[0].decode()
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + out)
        if verbose:
            print(out)
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-lcudart', '-lcuda'])
    subprocess.check_call([nvcc, '-shared', '-o', LIB_HOOKS] + hook_objs + ['-lcudart', '-lcuda'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
