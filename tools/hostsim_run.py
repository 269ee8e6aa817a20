"""Run any of the repository's GPU scripts on the host simulation (tests/hostsim) instead of a GPU:

    python tools/hostsim_run.py [--execute] bench.py --steps 12 --warmup 3 --no-cpu-baseline
    BM_NCCL_LIB=tests/hostsim/_build/libfakenccl.so python -m torch.distributed.run --nproc-per-node 2 \
        tools/hostsim_run.py bench.py --gpus 2 --steps 12 --warmup 3

The library's own objects run on a stand-in CUDA runtime: launches, copies and tensor maps are checked (and counted), kernels are
skipped -- or, with --execute, interpreted on the CPU (slow at benchmark sizes).  Timings printed by the scripts are meaningless
there; what the run shows is whether the script's control flow holds together and whether any runtime rule is broken."""
import ctypes as C
import os
import runpy
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    args = sys.argv[1:]
    execute = bool(args and args[0] == '--execute')
    if execute:
        args = args[1:]
    if not args:
        raise SystemExit(__doc__)
    if int(os.environ.get('LOCAL_RANK', '0')) == 0:
        subprocess.check_call(['bash', os.path.join(ROOT, 'tests', 'hostsim', 'build.sh')], stdout=subprocess.DEVNULL)
    from boltzmann_machines import _native
    sim = _native.load_library(os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libbm_hostsim.so'))
    sim.fakecuda_violation.restype = C.c_char_p
    sim.fakecuda_skipped.restype = C.c_char_p
    sim.fakecuda_set_execute(1 if execute else 0)
    _native._lib = sim
    sys.argv = args
    try:
        runpy.run_path(args[0], run_name='__main__')
    finally:
        sys.stderr.write('[hostsim] launches {0}, host syncs {1}, H2D {2} B, D2H {3} B; violations: {4}; not interpreted: {5}\n'.format(
            sim.fakecuda_launches(b''), sim.fakecuda_syncs(), sim.fakecuda_h2d_bytes(), sim.fakecuda_d2h_bytes(),
            sim.fakecuda_violation().decode() or 'none', (sim.fakecuda_skipped().decode() or 'none') if execute else 'n/a'))


if __name__ == '__main__':
    main()
