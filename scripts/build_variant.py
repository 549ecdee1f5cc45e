"""Build an experimental variant of libtikeb200.so next to the in-tree one:

    python scripts/build_variant.py NAME file.cu -DFLAG=1 [...]

recompiles `file.cu` with the extra flags and links it with the objects of the
regular build into tike_b200/lib/libtikeb200_NAME.so; select it at run time
with TB_LIB_PATH.  Development aid for A/B timing on the GPU box."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tike_b200 import build as B  # noqa: E402


def main():
    name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    B.build()
    nvcc = B._nvcc()
    obj = os.path.join(B.OBJDIR, f'{src[:-3]}_{name}.o')
    subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, '-c', os.path.join(B.CSRC, src), '-o', obj],
                   check=True, capture_output=True)
    objs = [os.path.join(B.OBJDIR, s[:-3] + '.o') for s in B.sources() if s != src] + [obj]
    lib = os.path.join(B.LIBDIR, f'libtikeb200_{name}.so')
    subprocess.run([nvcc, '-shared', '-o', lib, *objs, '-lcudart', '-gencode',
                    'arch=compute_100a,code=sm_100a'], check=True)
    print(lib)


if __name__ == '__main__':
    main()
