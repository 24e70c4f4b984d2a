"""TEST INFRASTRUCTURE ONLY -- builds libgpslim_b200 FOR THE CPU out of the unmodified .cu sources:
each file is transformed textually (list below), compiled with g++ against the stand-in
<cuda_runtime.h> / <cuda.h> of this directory and linked into one shared library that exports the
same C ABI.  tests/test_library_on_cpu.py loads it through the package's own ctypes binding.

Transformations (all mechanical):
  1. k<<<grid, block, smem, stream>>>(args);  ->  EMU_LAUNCH(grid, block, smem, k(args));
  2. extern __shared__ [__align__(16)] double NAME[];  ->  double* NAME = emu_smem;
  3. asm volatile("prefetch.global.L2 ...");   ->  (void)0;      (a cache hint)
  4. gemm.cu: the six inline-PTX helper functions of the TMA + mbarrier kernel (mbarrier init / arrive /
     expect_tx / wait, cp.async.bulk.tensor.2d, ld.shared) are replaced by the functional stand-ins of
     tma_emulation.inc and the tensor-map encoder by a plain description of the operand view: the
     kernel's own main loop runs; gemm_impl 2 still selects the cp.async tensor-core kernel
  5. handle.cu: DLPack device_type 2 (kDLCUDA) -> 1 (kDLCPU) in the argument checks
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, 'gpflow-slim_b200', 'csrc')
FILES = ['handle.cu', 'gemm.cu', 'potrf.cu', 'gram.cu', 'gpr.cu', 'adjoint.cu']


def _split_top(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip())
            cur = ''
        else:
            cur += ch
    out.append(cur.strip())
    return out


_LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;(]*>)?)\s*<<<(.*?)>>>\s*\(', re.S)


def _rewrite_launches(src):
    out, pos, n = '', 0, 0
    while True:
        m = _LAUNCH.search(src, pos)
        if not m:
            break
        # find the matching ')' of the argument list
        i, depth = m.end(), 1
        while depth:
            depth += {'(': 1, ')': -1}.get(src[i], 0)
            i += 1
        args = src[m.end():i - 1]
        assert src[i:].lstrip().startswith(';'), src[m.start():i + 20]
        cfg = _split_top(m.group(2))
        assert len(cfg) == 4, cfg
        out += src[pos:m.start()] + 'EMU_LAUNCH(%s, %s, %s, %s(%s))' % (cfg[0], cfg[1], cfg[2], m.group(1), args)
        pos = i
        n += 1
    return out + src[pos:], n


def transform(name, src):
    src, n_launch = _rewrite_launches(src)
    assert '<<<' not in src
    src = re.sub(r'extern __shared__ (?:__align__\(16\) )?double (\w+)\[\];', r'double* \1 = emu_smem;', src)
    src = re.sub(r'asm volatile\("prefetch\.global\.L2.*?\)\);', '(void)0;', src, flags=re.S)
    if name == 'gemm.cu':
        # the six PTX helpers of the TMA kernel -> functional stand-ins; the kernel body stays
        a = src.index('__device__ __forceinline__ void mbar_init(')
        b = src.index('__global__ void __launch_bounds__(TMA_THREADS, 1)')
        assert src[a:b].count('asm volatile') == 6
        src = src[:a] + open(os.path.join(HERE, 'tma_emulation.inc')).read() + '\n' + src[b:]
        src, k = re.subn(r'asm volatile\("fence\.mbarrier_init\.release\.cluster;\\n" ::: "memory"\);', '(void)0;', src)
        assert k == 1, k
        a = src.index('typedef CUresult (*EncodeTiledFn)')
        b = src.index('double gemm_flops(const GemmArgs& g) {', a)
        src = src[:a] + ('bool make_tensor_map(CUtensorMap* tm, const Mat& X) {\n'
                         '  tm->p = X.p; tm->cols = (uint64_t)X.cols; tm->rows = (uint64_t)X.rows; tm->ld = (uint64_t)X.ld;\n'
                         '  return true;\n}\n\n') + src[b:]
    if name == 'handle.cu':
        src, k = re.subn(r'device_type != 2', 'device_type != 1', src)
        assert k == 2, k
    assert 'asm' not in re.sub(r'//.*', '', src), name
    return src, n_launch


def build(outdir):
    objs, launches = [], 0
    for f in FILES:
        src, n = transform(f, open(os.path.join(CSRC, f)).read())
        launches += n
        cpp = os.path.join(outdir, f.replace('.cu', '_cpu.cpp'))
        open(cpp, 'w').write(src)
        obj = cpp.replace('.cpp', '.o')
        cmd = ['g++', '-std=c++17', '-O1', '-fPIC', '-pthread', '-Wno-attributes', '-Wno-unused-function',
               '-Wno-unused-variable', '-I', HERE, '-I', os.path.dirname(HERE), '-I', CSRC, '-I', os.path.join(ROOT, 'include'), '-c', cpp, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError('%s:\n%s' % (f, res.stderr[-5000:]))
        objs.append(obj)
    so = os.path.join(outdir, 'libgpslim_b200_cpu.so')
    res = subprocess.run(['g++', '-shared', '-pthread', '-o', so] + objs, capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError(res.stderr[-5000:])
    assert launches >= 40, launches     # every kernel launch of the library was rewritten
    return so


if __name__ == '__main__':
    import sys
    import tempfile
    print(build(sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp()))
