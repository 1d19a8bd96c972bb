"""Build libbmc_b200.so in-tree with nvcc for sm_100a (one explicit recipe, no JIT cache).

    python -m bmcnet_esr_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the
working tree to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libbmc_b200.so')
SOURCES = ['util.cu', 'encode.cu', 'gemm_tc.cu', 'gemm_slab.cu', 'gemm_slabt.cu', 'gemm_slab2.cu', 'bie_fused.cu', 'gemm_simt.cu', 'pointwise.cu', 'eval_tail.cu', 'redistribute.cu', 'train.cu', 'model.cu']
HEADERS = ['common.cuh', 'gemm.cuh', 'gemm_epi.cuh', os.path.join('..', '..', 'include', 'bmc_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


MEASURE_SOURCES = ['gemm_pair.cu']      # experiment kernels: only in --measure libraries


def build(force=False, verbose=True, bf16=False, measure=False):
    """Compile every CUDA source to an object (in parallel) and link the shared library.
    bf16=True builds the bf16-operand variant as libbmc_b200_bf16.so; measure=True builds
    libbmc_b200_measure.so with -DBMC_MEASURE: the environment switches of DESIGN.md section 6.2, the per-role
    cycle counters and the non-computing measurement variants of the kernels exist ONLY there (select a variant
    library at run time with BMC_B200_LIB=<path>).  The default library uses fp16 operands (Precision)
    and reads no environment variable."""
    tag = ('bf16' if bf16 else 'f16') + ('_measure' if measure else '')
    objdir = os.path.join(HERE, 'build', tag)
    flags = NVCC_FLAGS + (['-DBMC_ACT_BF16'] if bf16 else []) + (['-DBMC_MEASURE'] if measure else [])
    lib = LIB.replace('.so', ('_bf16' if bf16 else '') + ('_measure' if measure else '') + '.so')
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    procs, objs = [], []
    for src in SOURCES + (MEASURE_SOURCES if measure else []):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [_nvcc()] + flags + ['-c', s, '-o', o]
            if verbose:
                print(' '.join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, out))
        if verbose and out.strip():
            print(out)
    if force or procs or _stale(lib, objs):
        cmd = [_nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, bf16='--bf16' in sys.argv, measure='--measure' in sys.argv))
