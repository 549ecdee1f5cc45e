"""A/B of compile-time variants of the fused rPIE kernel (development aid).

    python scripts/ab_variants.py build      # here (no GPU): nvcc the variants
    python scripts/ab_variants.py run        # on the GPU box: time + parity per variant

`build` makes tike_b200/lib/libtikeb200_<name>.so for every entry of VARIANTS
(scripts/build_variant.py); `run` loads each library through TB_LIB_PATH in a
fresh process, times the fused kernel at the bench batch size
(scripts/batch_order_experiment.py) and runs the per-batch parity tests on it.
One gpurun call covers all of them, e.g.
    gpurun --timeout 300 -- 'python scripts/ab_variants.py run > gpurun_out/ab.log 2>&1'
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {  # name: (source file, extra nvcc flags)
    'noclobber': ('rpie_fast.cu', ['-DTB_EXP_NO_CLOBBER=1']),
    'notma': ('rpie_fast.cu', ['-DTB_EXP_TMA_WINDOW=0']),
    # round 2: threads per CTA of the 128^2 kernel (512 = 16 warps, 128 registers)
    # probe-numerator replicas taking the REDs (host side of the fused launch)
}
# measured in round 2 and not adopted (profiles/r02a_variants.log,
# profiles/r02e_variants_threads_replicas.log): nt1024 / nt256 (-DTB_EXP_NT128=...),
# rep32 / rep64 (rpie.cu -DTB_MAX_REPLICAS=...), hp / pv
# (-DTB_EXP_HOIST_PROBE / _PV), am (-DTB_EXP_APPROX_MODULUS), gp
# (-DTB_EXP_GROUP_PIPE), swapinv (-DTB_EXP_SWAP_INVERSE)
PARITY = 'rpie_batch_golden or rpie_batch_vs_oracle or colliding or lstsq_batch'


def build():
    for name, (src, flags) in VARIANTS.items():
        subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'build_variant.py'),
                        name, src, *flags], check=True, cwd=ROOT)


def run():
    libs = [os.path.join(ROOT, 'tike_b200', 'lib', 'libtikeb200.so')] + sorted(
        glob.glob(os.path.join(ROOT, 'tike_b200', 'lib', 'libtikeb200_*.so')))
    for lib in libs:
        env = dict(os.environ, TB_LIB_PATH=lib)
        print(f'== {os.path.basename(lib)}', flush=True)
        t = subprocess.run([sys.executable, 'scripts/batch_order_experiment.py'], cwd=ROOT,
                           env=env, capture_output=True, text=True, timeout=120)
        print('\n'.join((t.stdout + t.stderr).strip().splitlines()[-2:]), flush=True)
        p = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_kernels.py', '-q',
                            '-p', 'no:cacheprovider', '-k', PARITY], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=300)
        print((p.stdout.strip().splitlines() or ['(no output)'])[-1], flush=True)


if __name__ == '__main__':
    {'build': build, 'run': run}[sys.argv[1] if len(sys.argv) > 1 else 'run']()
