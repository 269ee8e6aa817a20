#!/usr/bin/env python
"""Headline benchmark: Gibbs updates/s (batch x k) of CD-5 training, BernoulliRBM 784-1024,
batch 4096 per GPU (BASELINE.json configs[1]), on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle restatement)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
A "step" = one full CD-5 mini-batch step: h0, 5 Gibbs sweeps, dW/dvb/dhb, sparsity, momentum update.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'boltzmann-machines_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

V, H, B, K_GIBBS = 784, 1024, 4096, 5
N_BATCHES = int(os.environ.get('BM_BENCH_BATCHES', '40'))    # resident dataset: 40 x 4096 rows = 257 MB of bf16 > 126 MB L2
#                                      (BM_BENCH_BATCHES: dry runs on the host simulation only -- `config.l2_policy` states the size)
LR, MOMENTUM, L2 = 0.05, 0.5, 1e-5
FLOP_PER_STEP = 2.0 * B * V * H * (2 * K_GIBBS + 3)      # (2k+3) GEMMs of 2BVH


def synth_mnist(n_rows, seed=1337, gibbs=30):
    """Binary 'MNIST-shaped' data: samples of a random teacher RBM 784-64 with a 13% on-rate bias
    (SURVEY.md §8d), generated once for `base` rows and tiled."""
    rng = np.random.RandomState(seed)
    base = min(n_rows, 8192)
    Wt = (0.5 * rng.randn(V, 64)).astype(np.float32)
    bt = np.float32(np.log(0.13 / 0.87))
    v = (rng.rand(base, V) < 0.5).astype(np.float32)
    for _ in range(gibbs):
        h = (rng.rand(base, 64) < 1.0 / (1.0 + np.exp(-(v @ Wt)))).astype(np.float32)
        v = (rng.rand(base, V) < 1.0 / (1.0 + np.exp(-(h @ Wt.T + bt)))).astype(np.float32)
    reps = (n_rows + base - 1) // base
    return np.ascontiguousarray(np.tile(v, (reps, 1))[:n_rows])


def model_cfg(compute='bf16'):
    return dict(n_visible=V, n_hidden=H, v_kind='bernoulli', h_kind='bernoulli', dtype='float32',
                compute=compute, l2=L2, sample_v=False, sample_h=True, max_batch=B,
                sparsity_target=0.1, sparsity_cost=0.0, sparsity_damping=0.9)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons, one sample every 100 ms from before the warm-up to the end
    of the run; every sample is stamped on arrival and only those that fall inside a timed region
    (`mark()` ... `unmark()`) are reported, i.e. the median is a median UNDER LOAD."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.proc, self.samples, self.windows, self._t0 = device, None, [], [], None

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), line))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def mark(self):
        self._t0 = time.perf_counter()

    def unmark(self):
        self.windows.append((self._t0, time.perf_counter()))

    def n_loaded(self):
        return sum(1 for t, _ in self.samples if any(a + 0.05 <= t <= b for a, b in self.windows))

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for t, line in self.samples:
            # a sample describes the 100 ms before it: require it to lie well inside a loaded window
            if not any(a + 0.05 <= t <= b for a, b in self.windows):
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0))), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
    return 1400.0, 'fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)'


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, 'profiles', 'tc_program_kernel_traffic.json')
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh).get('dram_bytes_per_launch')
    return None


# --------------------------------------------------------------------------------------------
# CPU side: the oracle (numpy + C Philox) timed on the host cores
# --------------------------------------------------------------------------------------------
def time_oracle(steps, warmup):
    """Times the oracle's CD-5 step.  OpenBLAS with every hardware thread of a 128-core host is several times
    SLOWER on these 4096 x 784 x 1024 GEMMs than with a few dozen threads, so the BLAS pool size is calibrated
    first (one step per candidate) and the fastest setting is the one reported -- the CPU arm at its best."""
    from oracle.rbm import OracleRBM
    X = synth_mnist(B * 2)
    ora = OracleRBM(model_cfg('fp32'))
    rng = np.random.RandomState(0)
    ora.set_params({'W': (0.01 * rng.randn(V, H)).astype(np.float32)})
    n_cpu = os.cpu_count() or 1
    threads = n_cpu
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        threadpool_limits = None
    tick = [0]

    def run(n, limit):
        t0 = time.perf_counter()
        for i in range(n):
            if threadpool_limits is not None:
                with threadpool_limits(limits=limit, user_api='blas'):
                    ora.train_step(X[(i % 2) * B:(i % 2 + 1) * B], LR, MOMENTUM, K_GIBBS, 1, tick[0])
            else:
                ora.train_step(X[(i % 2) * B:(i % 2 + 1) * B], LR, MOMENTUM, K_GIBBS, 1, tick[0])
            tick[0] += 1
        return time.perf_counter() - t0

    run(max(1, warmup), n_cpu)
    if threadpool_limits is not None and n_cpu > 8:
        cands = sorted({c for c in (8, 16, 32, 64, n_cpu) if c <= n_cpu})
        times = {c: run(1, c) for c in cands}
        threads = min(times, key=times.get)
    dt = run(steps, threads)
    return steps * B * K_GIBBS / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))          # ~1.5 s of CPU work per step: bounded sample
    warm = max(1, min(args.warmup, 2))
    val, dt, cores = time_oracle(steps, warm)
    sample = ('{0} CD-5 steps of batch 4096 (784-1024), numpy/OpenBLAS + C Philox; BLAS pool calibrated to its fastest size '
              '({1} of {2} hardware threads)').format(steps, cores, os.cpu_count())
    print(json.dumps({
        'impl': 'reference', 'metric': 'gibbs_updates_per_sec', 'value': val, 'unit': 'updates/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * dt / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BernoulliRBM 784-1024, batch 4096, CD-5 (BASELINE.json configs[1])',
                   'note': 'TF1/py2 reference cannot run in this image; timed: oracle/ restatement of its CPU path'},
        'cpu_baseline': {'value': val, 'unit': 'updates/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--compute', default='bf16')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist            # plumbing only: rendezvous, barrier, max-reduce
        dist.init_process_group('gloo')

    from boltzmann_machines import _native
    ctx = _native.Context(local)
    if world > 1:
        import torch
        uid = [_native.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_rows = B * N_BATCHES
    X = synth_mnist(n_rows, seed=1337 + rank)
    eng = _native.CudaRBM(model_cfg(args.compute), ctx=ctx)
    eng.init_normal_W(0.01, 1337)
    p = np.clip(X[:8192].mean(axis=0), 1e-7, 1 - 1e-7)
    eng.set_params({'vb': np.log(p / (1 - p)).astype(np.float32)})
    eng.set_data(X)

    seed = 20260922
    tick = [0]

    def step_resident(i):
        eng.train_step_at((i % N_BATCHES) * B, B, LR, MOMENTUM, K_GIBBS, seed, tick[0])
        tick[0] += 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step_resident(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    l0 = ctx.launch_count()
    barrier()
    sampler.mark()
    ctx.timer_start()
    for i in range(args.steps):
        step_resident(args.warmup + i)
    ms = ctx.timer_stop()
    barrier()
    sampler.unmark()
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    value = args.steps * B * world * K_GIBBS / (ms * 1e-3)

    # ---- region 2: same steps with per-launch CUDA events on the tensor-core kernel -------------
    ctx.profile_tc(True)
    sampler.mark()
    for i in range(args.steps):
        step_resident(i)
    flops, tc_ms, tc_launches = ctx.profile_read()
    sampler.unmark()
    ctx.profile_tc(False)
    peak, peak_src = measured_peaks()
    achieved = flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic = recorded_traffic()

    # ---- region 3: end to end from HOST buffers through bm_rbm_train_epoch[_u8] ------------------
    # (the call BaseRBM._train_epoch makes on the array BaseRBM._fit pinned with engine.pin): every step
    # uploads its own batch from pinned host memory (double-buffered against the previous step's
    # compute) and reads its MSRE back.  engine.pin keeps this binary dataset as one byte per unit
    # (lossless; bit-identical results, tests/test_rbm_gpu.py) -> `e2e`; the same loop fed float32
    # rows -- the reference's feed_dict dtype -- is reported beside it as `e2e_float32`.
    def e2e_region(Xhost, first_tick):
        def epoch_chunks(n_steps, t0):
            done = 0
            while done < n_steps:
                nb = min(N_BATCHES, n_steps - done)
                eng.train_epoch(Xhost[:nb * B], B, LR, MOMENTUM, K_GIBBS, seed, t0 + done,
                                metrics=('msre',), every=1)
                done += nb
        # warm-up: one full-size epoch call, so that every staging buffer exists before the timed passes
        epoch_chunks(max(N_BATCHES, args.warmup), first_tick)
        passes = []
        for rep in range(2):                        # two timed passes of exactly K steps; both are reported
            barrier()
            sampler.mark()
            ctx.timer_start()
            epoch_chunks(args.steps, first_tick + N_BATCHES + rep * args.steps)
            t = ctx.timer_stop()
            barrier()
            sampler.unmark()
            passes.append(max_over_ranks(t))
        return min(passes), passes

    Xpin = eng.pin(X)                               # uint8 for this dataset (asserted below)
    assert Xpin.dtype == np.uint8, 'engine.pin did not take the byte path for binary data'
    e2e_steps = args.steps
    e2e_ms, e2e_passes = e2e_region(Xpin, tick[0]); tick[0] += N_BATCHES + args.warmup + 2 * e2e_steps
    e2e_value = e2e_steps * B * world * K_GIBBS / (e2e_ms * 1e-3)
    _native.pinned_free(Xpin)
    Xpin32 = _native.pinned_copy(X)
    e2e32_ms, e2e32_passes = e2e_region(Xpin32, tick[0]); tick[0] += N_BATCHES + args.warmup + 2 * e2e_steps
    e2e32_value = e2e_steps * B * world * K_GIBBS / (e2e32_ms * 1e-3)
    _native.pinned_free(Xpin32)

    # clock probe: when K is so small that no 100 ms sample fell inside a timed region, keep the
    # same load running (untimed) until a few samples exist -- same work, same clocks
    # (with peers every step is a collective: rank 0 decides, every rank runs the same number of probe steps)
    def rank0_says(flag):
        if dist is None:
            return flag
        box = [bool(flag) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    probe = 0
    if rank == 0:
        sampler.mark()
    t_end = time.perf_counter() + 3.0
    while rank0_says(rank == 0 and sampler.n_loaded() < 5 and time.perf_counter() < t_end):
        for i in range(50):
            step_resident(i)
        ctx.sync()
        if rank == 0:
            sampler.windows.append((sampler._t0, time.perf_counter()))
        probe += 50
    if rank == 0:
        clocks = sampler.stop()
        clocks['probe_steps_after_timed_regions'] = probe
    barrier()

    if rank != 0:
        return
    out = {
        'metric': 'gibbs_updates_per_sec', 'value': value, 'unit': 'updates/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.compute == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': 'BernoulliRBM 784-1024, batch 4096 per GPU, CD-5 (BASELINE.json configs[1])',
                   'global_batch': B * world, 'k': K_GIBBS, 'parallelism': 'dp{0}'.format(world),
                   'l2_policy': 'inputs larger than L2: resident dataset {0} MB bf16, batches cycle'.format(
                       n_rows * 832 * 2 // 2 ** 20),
                   'flop_per_step': FLOP_PER_STEP},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak if peak else None, 'traffic': traffic,
                     'kernel': 'bm::tc_program_kernel<2>', 'launches': int(tc_launches), 'peak_source': peak_src,
                     'step_tflops': FLOP_PER_STEP * args.steps / (ms * 1e-3) / 1e12 / 1.0,
                     'step_frac': FLOP_PER_STEP * args.steps / (ms * 1e-3) / 1e12 / peak},
        'e2e': {'value': e2e_value, 'unit': 'updates/s', 'h2d_bytes_per_step': B * V * 1 * world,
                'd2h_bytes_per_step': 64 * world, 'ms_per_step': e2e_ms / e2e_steps,
                'timed_passes_ms_per_step': [t / e2e_steps for t in e2e_passes], 'reported': 'faster of two passes of K steps',
                'path': 'bm_rbm_train_epoch_u8(pinned host dataset as BaseRBM.fit/engine.pin stores binary data: 1 byte per '
                        'unit, widened exactly on the device; msre read back every step): what BaseRBM._train_epoch calls'},
        'e2e_float32': {'value': e2e32_value, 'unit': 'updates/s', 'h2d_bytes_per_step': B * V * 4 * world,
                        'd2h_bytes_per_step': 64 * world, 'ms_per_step': e2e32_ms / e2e_steps,
                        'timed_passes_ms_per_step': [t / e2e_steps for t in e2e32_passes],
                        'path': 'bm_rbm_train_epoch(pinned host float32 dataset, msre every step): PCIe-bound'},
    }
    if world == 1 and not args.no_cpu_baseline:
        cpu_steps = 8
        val, dt, cores = time_oracle(cpu_steps, 1)
        out['cpu_baseline'] = {'value': val, 'unit': 'updates/s', 'cores': cores, 'kind': 'port',
                               'sample': ('{0} CD-5 steps of batch 4096 on the oracle (numpy/OpenBLAS + C Philox), BLAS pool '
                                          'calibrated to its fastest size ({1} of {2} hardware threads)').format(
                                              cpu_steps, cores, os.cpu_count())}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
