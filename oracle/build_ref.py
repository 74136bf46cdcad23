"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by dynavsr_b200/).

Builds the reference's own native code and stages its own Python modules, UNMODIFIED, for the GPU box:

  oracle/_ref/deform_conv_cuda*.so   the pybind11 extension `deform_conv_cuda` compiled for sm_100a from
                                      /root/reference/codes/models/archs/dcn/src/{deform_conv_cuda.cpp,
                                      deform_conv_cuda_kernel.cu} WHERE THEY LIE (nothing is copied into the repo);
                                      recipe = dcn/setup.py:5-22 + SURVEY.md Appendix A (-DAT_CHECK=TORCH_CHECK for
                                      torch >= 1.5, arch 10.0a).  nvcc cross-compiles here without a GPU.
  baseline/_ref/codes/               a verbatim copy of /root/reference/codes (git-ignored, NOT gpurun-ignored):
                                      /root/reference does not exist on the GPU box, and the reference-CUDA arm of
                                      bench.py / tools/ref_cuda_bench.py imports models.archs.EDVR_arch etc. from it.

Both directories are git-ignored (outputs only); this script is the committed recipe.  It is a no-op when
/root/reference is absent (GPU box: the prebuilt files travel with the snapshot).

    python oracle/build_ref.py [--force]
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
SRC = os.path.join(REF, 'codes', 'models', 'archs', 'dcn', 'src')
OUT = os.path.join(HERE, '_ref')
PY_COPY = os.path.join(ROOT, 'baseline', '_ref', 'codes')


def ext_path():
    hits = sorted(glob.glob(os.path.join(OUT, 'deform_conv_cuda*.so')))
    return hits[0] if hits else None


def build(force=False, verbose=False):
    if not os.path.isdir(SRC):
        return ext_path()
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(SRC, 'deform_conv_cuda.cpp'), os.path.join(SRC, 'deform_conv_cuda_kernel.cu')]
    so = ext_path()
    if force or so is None or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
        os.environ.setdefault('MAX_JOBS', '4')
        from torch.utils import cpp_extension
        flags = ['-DAT_CHECK=TORCH_CHECK']
        cpp_extension.load(name='deform_conv_cuda', sources=srcs, build_directory=OUT, extra_cflags=flags + ['-O2'],
                           extra_cuda_cflags=flags + ['-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__',
                                                      '-D__CUDA_NO_HALF2_OPERATORS__', '-lineinfo'],
                           is_python_module=False, verbose=verbose)
        so = ext_path()
    # the reference's Python tree for the GPU box (verbatim; pretrained_models / figures are not needed)
    stamp = os.path.join(PY_COPY, '.copied_from_reference')
    if force or not os.path.exists(stamp):
        if os.path.isdir(PY_COPY):
            for d, _, _f in os.walk(PY_COPY):
                os.chmod(d, 0o755)
            shutil.rmtree(PY_COPY)
        shutil.copytree(os.path.join(REF, 'codes'), PY_COPY, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        for d, _, files in os.walk(PY_COPY):          # /root/reference is read-only; the copy must stay removable
            os.chmod(d, 0o755)
            for f in files:
                os.chmod(os.path.join(d, f), 0o644)
        open(stamp, 'w').write('verbatim copy of /root/reference/codes made by oracle/build_ref.py\n')
    return so


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
