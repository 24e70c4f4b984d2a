#!/usr/bin/env python
"""Headline benchmark: GPR NLML + gradient evaluations per second, FP64, ARD-RBF, synthetic data
(BASELINE.json metric; SURVEY.md section 8d inputs).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference's CPU path (oracle port)

A "step" is one pass of the hot path over one data set: Gram -> Cholesky -> alpha -> NLML and
the full gradient w.r.t. (variance, D lengthscales, noise).  `value` = steps/s with X, Y
resident in HBM; `e2e` = the same through the public gpflowSlim API from pinned HOST buffers
(H2D of X, Y and D2H of objective + gradient inside the timed region).  One JSON line on
rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'gpflow-slim_b200'))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

NOMINAL_FP64_TFLOPS = 37.0   # B200 datasheet (DGX B200: 296 TF / 8); used only if nothing measured


def synth_gpr(n, d, seed=0):
    """X ~ N(0,1)^{N x D}, Y = sin(X.1/sqrt(D)) + 0.1 eps."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(d)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def fp64_peak():
    """Measured cuBLAS DGEMM throughput on this pool's B200 (tools/measure_fp64.py ->
    profiles/fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure."""
    p = os.path.join(ROOT, 'profiles', 'fp64_peak.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'burst': d['cublas_dgemm_8192_tflops_burst'],
                'sustained': d['cublas_dgemm_8192_tflops_sustained'],
                'source': 'measured cuBLAS DGEMM 8192^3 (profiles/fp64_peak.json)'}
    return {'burst': NOMINAL_FP64_TFLOPS, 'sustained': NOMINAL_FP64_TFLOPS,
            'source': 'nominal datasheet FP64 (no measurement available)'}


def gemm_traffic():
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r01_gemm8192_tma_ncu_full.json: one 8192^3 launch), with its algorithmic bytes."""
    p = os.path.join(ROOT, 'profiles', 'r01_gemm8192_tma_ncu_full.json')
    if not os.path.exists(p):
        return None, None
    s = json.load(open(p))['_summary']
    return s['dram_traffic_bytes_per_launch'], {
        'launch': s['kernel'], 'algorithmic_bytes': s['algorithmic_bytes_per_launch'],
        'source': 'profiles/r01_gemm8192_tma_ncu_full.json (dram__bytes_read.sum + dram__bytes_write.sum)'}


class ClockSampler(threading.Thread):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def oracle_eval(n, d, reps, threads=None):
    """Reference's CPU path (oracle port): NLML + autograd gradient on a bounded sample."""
    from oracle import ref_torch as R
    if threads:
        torch.set_num_threads(threads)
    X, Y = synth_gpr(n, d)
    X, Y = torch.tensor(X), torch.tensor(Y)
    times = []
    for _ in range(reps):
        raw = [torch.tensor(R.softplus_inv(v), dtype=torch.float64, requires_grad=True)
               for v in (1.0, math.sqrt(d) * np.ones(d), 0.1)]
        t0 = time.perf_counter()
        spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
        obj = R.gpr_nlml(spec, X, Y, R.softplus_fwd(raw[2]))
        torch.autograd.grad(obj, raw)
        times.append(time.perf_counter() - t0)
    return times


def potrf_metric(gpf, model, dev, n, peak, reps=3):
    """BASELINE.json metric (ii): Cholesky TFLOP/s (N^3/3 flop) of K + noise I at the bench size,
    gps_potrf timed alone with CUDA events, against the measured FP64 peak."""
    import ctypes
    from gpflowSlim._backend import lib as L
    h = L.handle_for(dev)
    with torch.no_grad():
        K = model.kern.K(model.X)
        K.diagonal().add_(float(model.likelihood.variance))
        A = torch.empty_like(K)
        times = []
        for _ in range(reps + 1):
            A.copy_(K)
            va = L.view(A)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.check(h.lib.gps_potrf(h.ptr, va.ref, 0, None))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times[1:]))
    tf = float(n) ** 3 / 3.0 / (ms * 1e-3) / 1e12
    del K, A
    return {'n': n, 'ms': ms, 'tflops': tf, 'frac_of_fp64_peak': tf / peak['burst'], 'flops': 'N^3/3'}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (torch-CPU fp64 restatement; TensorFlow 1.x is not installable), all host threads, each
    step a bounded sample (N_s = args.cpu_n) of the workload; reported in the workload's unit
    by the N^3 cost model of the path (Cholesky + its adjoint), which is stated in `sample`."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    n, d, ns = args.n, args.d, args.cpu_n
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    times = oracle_eval(ns, d, args.warmup + args.steps)[args.warmup:]
    t = float(np.mean(times))
    scale = (float(n) / ns) ** 3
    val = 1.0 / (t * scale)
    sample = ('oracle port (torch-CPU fp64 restatement of the reference TF path), NLML+grad at N_s=%d D=%d, '
              '%.2f s/eval on %d threads; scaled to N=%d by (N/N_s)^3=%.0f' % (ns, d, t, cores, n, scale))
    line = {'impl': 'reference', 'metric': 'GPR NLML+grad evals/s', 'value': val, 'unit': 'evals/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * scale * 1e3,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': {'workload': 'GPR ARD-RBF N=%d D=%d fp64 NLML+grad' % (n, d)},
            'cpu_baseline': {'value': val, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--size', '--n', type=int, default=32768, dest='n')   # under torchrun use --size
    ap.add_argument('--d', type=int, default=8)
    ap.add_argument('--mode', default='auto', choices=['auto', 'fused', 'dist', 'independent'],
                    help='auto: 1 GPU -> fused single-GPU path; N GPUs -> ONE problem distributed over '
                         'the ranks (strong scaling); independent: one problem per GPU (weak)')
    ap.add_argument('--block', type=int, default=512)
    ap.add_argument('--cpu-n', type=int, default=4096, dest='cpu_n')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-pass', action='store_true',
                    help='take the roofline figures from a separate profiled pass (the N > 1 behaviour)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    gpf.settings.device = dev
    n, d = args.n, args.d
    mode = args.mode
    if mode == 'auto':
        mode = 'dist' if world > 1 else 'fused'
    # dist: every rank holds the same (replicated) X, Y of ONE problem and the ranks factor it
    # together; independent: every rank owns its own problem (seed = rank).  DESIGN.md (e)
    Xh, Yh = synth_gpr(n, d, seed=rank if mode == 'independent' else 0)
    if mode == 'dist':
        gpf.parallel.init(block=args.block)
    Xp = torch.from_numpy(Xh).pin_memory()
    Yp = torch.from_numpy(Yh).pin_memory()
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d))
    model = gpf.models.GPR(Xp.to(dev), Yp.to(dev), kern=kern)
    params = [p.unconstrained_tensor for p in model.parameters]
    h = handle_for(dev)

    def step():
        obj = model.objective
        return obj, torch.autograd.grad(obj, params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # GEMM launches are bracketed by CUDA events on their own stream (library option "profile").
    # One GPU: inside the timed region.  Several GPUs: the distributed path overlaps three streams
    # and the extra event records perturb that overlap, so the timed region runs clean and the
    # roofline figures come from a separate profiled pass right after it.
    profile_in_region = (world == 1) and not args.profile_pass
    if profile_in_region:
        h.set_option('profile', 1)
        h.profile_read(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    prof_steps = args.steps
    if not profile_in_region:
        prof_steps = min(2, args.steps)
        h.set_option('profile', 1)
        h.profile_read(reset=True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        p0.record()
        for _ in range(prof_steps):
            step()
        p1.record()
        barrier()
        prof_ms = p0.elapsed_time(p1) / prof_steps
    else:
        prof_ms = ms
    gemm_ms, gemm_flops, launches = h.profile_read(reset=True)
    launches = int(round(launches * args.steps / float(prof_steps)))   # launches of the timed region
    h.set_option('profile', 0)

    # end to end through the public API from pinned host buffers
    def e2e_step():
        model.X = Xp.to(dev, non_blocking=True)
        model.Y = Yp.to(dev, non_blocking=True)
        obj, grads = step()
        return obj.cpu(), [g.cpu() for g in grads]
    e2e_step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        e2e_step()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1) / args.steps

    if world > 1:
        tt = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])

    if rank == 0:
        peak = fp64_peak()
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # the GEMM launches run back to back inside a ~1 s step: sustained figure applies
        pk = peak['sustained']
        traffic, traffic_note = gemm_traffic()
        nparam = d + 2
        nprob = world if mode == 'independent' else 1     # problems evaluated per step, whole job
        par = {'fused': 'single GPU, fused path', 'independent': 'independent problem per GPU',
               'dist': 'one problem over %d GPU(s): block-row (block %d) distributed Cholesky + inverse, '
                       'NCCL broadcast / all-gather per panel, look-ahead 1' % (world, args.block)}[mode]
        line = {
            'metric': 'GPR NLML+grad evals/s', 'value': nprob * 1e3 / ms, 'unit': 'evals/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak' if mode == 'independent' else 'strong',
            'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': 'GPR ARD-RBF N=%d D=%d fp64 NLML+grad (Gram+POTRF+TRSM+backward)' % (n, d),
                       'parallelism': par,
                       'l2': 'inputs exceed L2 (K is %.1f GiB)' % (8.0 * n * n / 2 ** 30),
                       'flops_per_eval_model': float(n) ** 3},
            'e2e': {'value': nprob * 1e3 / e2e_ms, 'unit': 'evals/s',
                    'h2d_bytes_per_step': int(Xp.numel() * 8 + Yp.numel() * 8),
                    'd2h_bytes_per_step': int(8 * (1 + nparam))},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'kernel': 'gemm_nt_tma_kernel (FP64 DMMA, TMA-staged operands)', 'achieved': ach,
                         'peak': pk, 'unit': 'TFLOP/s', 'frac': ach / pk if pk else None, 'traffic': traffic,
                         'traffic_of': traffic_note,
                         'peak_source': peak['source'],
                         'gemm_share_of_step': gemm_ms / (prof_ms * prof_steps) if prof_ms > 0 else None,
                         'scope': ('rank 0, separate profiled pass of %d step(s) after the timed region%s' % (
                             prof_steps, ' (three overlapping streams: per-launch event time includes queueing '
                             'behind the other streams)' if world > 1 else ''))
                         if not profile_in_region else 'the GPU, events inside the timed region',
                         'step_tflops_vs_n3': nprob * float(n) ** 3 / (ms * 1e-3) / 1e12},
        }
        if world == 1:
            line['potrf'] = potrf_metric(gpf, model, dev, n, peak)
        if world == 1 and not args.no_cpu_baseline:
            ns = args.cpu_n
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            reps = 3
            times = oracle_eval(ns, d, reps + 1)[1:]
            t = float(np.mean(times))
            scale = (float(n) / ns) ** 3
            line['cpu_baseline'] = {
                'value': 1.0 / (t * scale), 'unit': 'evals/s', 'cores': cores, 'kind': 'port',
                'sample': 'oracle port, NLML+grad at N_s=%d D=%d: %.2f s/eval on %d threads, scaled by '
                          '(N/N_s)^3=%.0f to N=%d' % (ns, d, t, cores, scale, n)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
