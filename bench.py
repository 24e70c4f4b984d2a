#!/usr/bin/env python
"""Headline benchmark: GPR NLML + gradient evaluations per second, FP64, ARD-RBF, synthetic data
(BASELINE.json metric; SURVEY.md section 8d inputs).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference's CPU path (oracle port)

A "step" is one pass of the hot path over one data set: Gram -> Cholesky -> alpha -> NLML and
the full gradient w.r.t. (variance, D lengthscales, noise).  `value` = steps/s with X, Y
resident in HBM; `e2e` = the same through the public gpflowSlim API from pinned HOST buffers
(H2D of X, Y and D2H of objective + gradient inside the timed region).  One JSON line on
rank 0.  Outside the timed region the objective and gradient are compared with the CPU oracle's
scalars for this exact configuration (tests/golden/gpr_large_scalars.json, `parity_rel_err`,
tolerance 1e-8 relative -- a violation makes the process exit non-zero after printing the line).
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
WORKLOAD = 'GPR ARD-RBF N=%d D=%d fp64 NLML+grad (Gram+POTRF+TRSM+backward)'     # identical in both arms
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'gpr_large_scalars.json')
PARITY_RTOL = 1e-8           # north star: 1e-8 relative on NLML and its gradients


def synth_gpr(n, d, seed=0):
    """X ~ N(0,1)^{N x D}, Y = sin(X.1/sqrt(D)) + 0.1 eps."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(d)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def fp64_peak(measure_on=None):
    """FP64 roofline denominator.  MEASURED_PEAKS.json has no FP64 figure, so it is the cuBLAS DGEMM
    8192^3 rate measured IN THIS PROCESS on the same GPU right before the timed region
    (`measure_on` = device; torch.matmul = cuBLAS, used as the yardstick only); without a device:
    the committed measurement of tools/measure_fp64.py, else the datasheet."""
    if measure_on is not None:
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=measure_on)
        b = torch.randn(n, n, dtype=torch.float64, device=measure_on)
        c = torch.empty_like(a)
        for _ in range(2):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best, tot, reps = 1e30, 0.0, 8
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best, tot = min(best, ms), tot + ms
        del a, b, c
        fl = 2.0 * n ** 3
        return {'burst': fl / best / 1e9, 'sustained': fl / (tot / reps) / 1e9,
                'source': 'cuBLAS DGEMM 8192^3 measured in this process on this GPU '
                          '(MEASURED_PEAKS.json has no FP64 figure)'}
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
    (one 8192^3 launch), with its algorithmic bytes."""
    for name in ('r02_gemm8192_tma_ncu_full.json', 'r01_gemm8192_tma_ncu_full.json'):
        p = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(p):
            s = json.load(open(p))['_summary']
            return s['dram_traffic_bytes_per_launch'], {
                'launch': s['kernel'], 'algorithmic_bytes': s['algorithmic_bytes_per_launch'],
                'source': 'profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum)' % name}
    return None, None


def golden_scalars(n, d):
    """Objective and gradient of the bench configuration computed on the CPU by the oracle
    (oracle/gen_large_golden.py -> tests/golden/gpr_large_scalars.json), or None."""
    if d != 8 or not os.path.exists(GOLDEN):
        return None
    return json.load(open(GOLDEN))['cases'].get(str(n))


def parity_against_golden(model, obj, grads, n, d):
    """Max relative error of (objective, d/d variance, d/d lengthscales, d/d noise) against the CPU
    oracle's scalars.  The package differentiates w.r.t. the unconstrained tensors; the golden
    gradients are w.r.t. the constrained values: theta = softplus(raw) + 1e-6
    (transforms.py:145-146), so d theta / d raw = sigmoid(raw).  GPR.parameters is ordered
    (kernel variance, lengthscales, noise variance)."""
    gold = golden_scalars(n, d)
    if gold is None:
        return None
    cons = [(g.detach() / torch.sigmoid(p.unconstrained_tensor.detach())).reshape(-1).cpu().numpy()
            for p, g in zip(model.parameters, grads)]
    want = [np.array([gold['g_variance']]), np.array(gold['g_lengthscales']), np.array([gold['g_noise']])]
    errs = {'objective': abs(float(obj) - gold['nlml']) / abs(gold['nlml'])}
    for name, a, b in zip(('d_variance', 'd_lengthscales', 'd_noise'), cons, want):
        errs[name] = float(np.abs(a - b).max() / np.abs(b).max())
    return errs


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


def cpu_cost_exponent():
    """Exponent p of the measured cost law t ~ N^p of the oracle port between N = 8192 and 16384
    (tools/cpu_oracle_scaling.py -> profiles/r02_cpu_oracle_scaling.json; asymptotically 3, measured
    slightly below because the BLAS-3 share runs faster at larger N).  The CPU baseline is reported in
    the workload's unit with THIS law, which is the conservative choice (p < 3 makes the CPU look
    faster at full size than an N^3 extrapolation would)."""
    f = os.path.join(ROOT, 'profiles', 'r02_cpu_oracle_scaling.json')
    try:
        a = json.load(open(f))['autograd_route']
        p = math.log(a['16384']['seconds'] / a['8192']['seconds']) / math.log(2.0)
        return min(3.0, max(2.5, p))
    except Exception:
        return 3.0


def cpu_scale(n, ns):
    return (float(n) / ns) ** cpu_cost_exponent()


def cpu_sample_note(ns, d, t, cores, n):
    return ('oracle port (torch-CPU fp64 restatement of the reference TF path; TensorFlow 1.x is not '
            'installable), NLML + autograd gradient at N_s=%d D=%d: %.2f s/eval on %d threads; reported in '
            'the workload\'s unit by the measured cost law of the path, (N/N_s)^%.2f = %.1f to N=%d (law '
            'measured on the CPU at N_s = 4096 / 8192 / 16384, and for the LAPACK route at the full N=32768: '
            'profiles/r02_cpu_oracle_scaling.json)' % (ns, d, t, cores, cpu_cost_exponent(), cpu_scale(n, ns), n))


def potrf_metric(gpf, model, dev, n, peak, reps=3):
    """BASELINE.json metric (ii): Cholesky TFLOP/s (N^3/3 flop) of K + noise I at the bench size,
    gps_potrf timed alone with CUDA events, against the measured FP64 peak."""
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
    return {'n': n, 'ms': ms, 'tflops': tf, 'frac_of_fp64_peak': tf / peak['burst'], 'flops': 'N^3/3',
            'scope': 'gps_potrf alone, one GPU'}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (torch-CPU fp64 restatement; TensorFlow 1.x is not installable), all host threads, each
    step a bounded sample (N_s = args.cpu_n) of the workload; reported in the workload's unit
    by the N^3 cost model of the path (Cholesky + its adjoint), which is stated in `sample`."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    n, d, ns = args.n, args.d, min(args.cpu_n, args.n)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    times = oracle_eval(ns, d, args.warmup + args.steps)[args.warmup:]
    t = float(np.mean(times))
    scale = cpu_scale(n, ns)
    val = 1.0 / (t * scale)
    sample = cpu_sample_note(ns, d, t, cores, n)
    line = {'impl': 'reference', 'metric': 'GPR NLML+grad evals/s', 'value': val, 'unit': 'evals/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * scale * 1e3,
            'sample_ms_per_step': t * 1e3, 'steps_are_samples': True,
            'extrapolation': {'law': '(N/N_s)^%.2f (measured exponent)' % cpu_cost_exponent(), 'factor': scale, 'n_sample': ns,
                              'validated_by': 'profiles/r02_cpu_oracle_scaling.json'},
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': {'workload': WORKLOAD % (n, d)},
            'cpu_baseline': {'value': val, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- secondary configs
def nkn_c3_kernel(gpf, d):
    """BASELINE config C3: 6 primitives, Linear 6->8, Product 2, Linear 4->4,
    Product 2, Linear 2->1; Linear weights from numpy's global RNG seeded 0 (wrapper.py:100-104)."""
    k = gpf.kernels
    prims = [k.RBF(d, ARD=True, name='p0'), k.RBF(d, lengthscales=2.0, ARD=True, name='p1'),
             k.Periodic(d, period=1.0, lengthscales=1.0, name='p2'), k.Periodic(d, period=2.0, name='p3'),
             k.Linear(d, ARD=True, name='p4'), k.Linear(d, ARD=True, name='p5')]
    hparams = [dict(name='Linear', params=dict(input_dim=6, output_dim=8, name='l0')),
               dict(name='Product', params=dict(input_dim=8, step=2, name='l1')),
               dict(name='Linear', params=dict(input_dim=4, output_dim=4, name='l2')),
               dict(name='Product', params=dict(input_dim=4, step=2, name='l3')),
               dict(name='Linear', params=dict(input_dim=2, output_dim=1, name='l4'))]
    np.random.seed(0)
    return gpf.neural_kernel_network.NeuralKernelNetwork(d, prims, gpf.neural_kernel_network.NKNWrapper(hparams))


def _timed(fn, reps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def secondary_lines(gpf, dev, peak_tf, world, rank):
    """The other named BASELINE configs, time-boxed (a few seconds each), so that the driver's
    bench line carries every named shape: C2 GPR N=8192, C3 NKN GPR N=16384 (one GPU), C4 SVGP
    N=1M M=1024 B=8192 (one GPU: CUDA-graph replay; several: minibatch rows sharded + all-reduce).
    `frac` = model flops / time / the FP64 peak measured in this process."""
    out = {}
    conv = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    if world == 1:
        for tag, n, kern, reps in (('C2 GPR ARD-RBF N=8192 D=8', 8192,
                                    gpf.kernels.RBF(8, ARD=True, lengthscales=math.sqrt(8)), 5),
                                   ('C3 NKN GPR N=16384 D=8 (6 primitives, Linear/Product x5)', 16384,
                                    nkn_c3_kernel(gpf, 8), 2)):
            X, Y = synth_gpr(n, 8)
            m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
            params = [p.unconstrained_tensor for p in m.parameters]

            def step():
                obj = m.objective
                return obj, torch.autograd.grad(obj, params)
            ms, (obj, _) = _timed(step, reps, 2)
            tf = float(n) ** 3 / ms / 1e9
            out[tag.split(' ')[0]] = {'workload': tag + ' fp64 NLML+grad', 'metric': 'evals/s', 'value': 1e3 / ms,
                                      'ms': ms, 'objective': float(obj.detach()),
                                      'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': peak_tf,
                                                   'unit': 'TFLOP/s', 'frac': tf / peak_tf,
                                                   'flops_model': 'N^3 per eval'}}
            del m, X, Y
            torch.cuda.empty_cache()
    # C4
    n, d, M, B = 1000000, 16, 1024, 8192
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / 4.0) + 0.1 * rng.standard_normal((n, 1))
    Z = X[np.random.default_rng(2).permutation(n)[:M]].copy()
    Xd, Yd = conv(X), conv(Y)
    m = gpf.models.SVGP(Xd[:B], Yd[:B], gpf.kernels.RBF(d, ARD=True, lengthscales=4.0),
                        gpf.likelihoods.Gaussian(var=0.1), Z=Z, num_data=n)
    st = {'i': 0}
    if world == 1:
        gstep = gpf.training.GraphedStep(m, Xd[:B], Yd[:B], learning_rate=1e-3)

        def step4():
            i0 = (st['i'] * B) % (n - B)
            st['i'] += 1
            return gstep(Xd[i0:i0 + B], Yd[i0:i0 + B])
        how = 'one GPU, step (ELBO + all gradients + Adam) captured in a CUDA graph and replayed'
    else:
        params = m.trainable_tensors
        opt = gpf.training.AdamOptimizer(1e-3)
        bl = B // world

        def step4():
            i0 = (st['i'] * B) % (n - B)
            st['i'] += 1
            sl = slice(i0 + rank * bl, i0 + (rank + 1) * bl)
            obj, grads = gpf.parallel.svgp_objective_and_grads(m, Xd[sl], Yd[sl], params)
            opt.apply_gradients(zip(grads, params))
            return obj
        how = ('minibatch rows sharded over %d GPUs, Kuu / KL replicated, one all-reduce of the flat '
               'gradient' % world)
    ms, obj = _timed(step4, 30, 5)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    flop = 5.2e10
    out['C4'] = {'workload': 'C4 SVGP Gaussian N=1M D=16 M=1024 B=8192 whiten, full q_sqrt, fp64: ELBO + '
                             'grad + Adam step',
                 'metric': 'steps/s', 'value': 1e3 / ms, 'ms': ms, 'objective': float(obj), 'how': how,
                 'roofline': {'bound': 'tensor', 'achieved': flop / ms / 1e9, 'peak': peak_tf * world,
                              'unit': 'TFLOP/s', 'frac': flop / ms / 1e9 / (peak_tf * world),
                              'flops_model': '5.2e10 per step (SURVEY.md 8d)'}}
    return out


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
    ap.add_argument('--schedule', default='v1', choices=['v1', 'v2'],
                    help='look-ahead schedule of the distributed factorisation (v1: three streams, the faster one '
                         'on 8 x B200; v2: the five-stream pipeline)')
    ap.add_argument('--cpu-n', type=int, default=8192, dest='cpu_n')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the C2 / C3 / C4 lines')
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
    peak = fp64_peak(dev)          # same-box, same-process FP64 yardstick (on every rank: keeps them aligned)
    # dist: every rank holds the same (replicated) X, Y of ONE problem and the ranks factor it
    # together; independent: every rank owns its own problem (seed = rank).  DESIGN.md (e)
    Xh, Yh = synth_gpr(n, d, seed=rank if mode == 'independent' else 0)
    if mode == 'dist':
        gpf.parallel.init(block=args.block, lookahead=args.schedule)
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
    # One GPU: inside the timed region.  Several GPUs: the distributed path overlaps several streams
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
        obj_last, grads_last = step()
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

    # parity of the timed computation against the CPU oracle's scalars (outside the timed region)
    parity = None
    if mode != 'independent' or rank == 0:
        parity = parity_against_golden(model, obj_last.detach(), grads_last, n, d)

    # phases of the distributed path: a separate pass with CUDA-event marks on the main stream,
    # max over ranks per phase
    phases = None
    if mode == 'dist':
        from gpflowSlim._backend import dist_gpr
        dist_gpr.TIMER = dist_gpr.PhaseTimer()
        barrier()
        step()
        rep = dist_gpr.TIMER.report()
        dist_gpr.TIMER = None
        names = sorted(rep)
        tt = torch.tensor([rep[k] for k in names], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        phases = {k: float(v) for k, v in zip(names, tt)}

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

    secondary = None
    if not args.no_secondary and mode != 'independent':
        try:
            secondary = secondary_lines(gpf, dev, peak['burst'], world, rank)
        except Exception as e:      # the headline must not die with a secondary line
            secondary = {'error': repr(e)[:300]}

    rc = 0
    if rank == 0:
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # the GEMM launches run back to back inside a ~1 s step: sustained figure applies
        pk = peak['sustained']
        traffic, traffic_note = gemm_traffic()
        nparam = d + 2
        nprob = world if mode == 'independent' else 1     # problems evaluated per step, whole job
        par = {'fused': 'single GPU, fused path', 'independent': 'independent problem per GPU',
               'dist': ('one problem over %d GPU(s): block-row (block %d) distributed Cholesky + inverse; schedule %s '
                        '(v1: three-stream look-ahead; v2: five-stream pipeline); NCCL broadcasts (diagonal / top '
                        'block) + all-gather per panel on three communicators' % (world, args.block, args.schedule))}[mode]
        step_tf = nprob * float(n) ** 3 / (ms * 1e-3) / 1e12
        line = {
            'metric': 'GPR NLML+grad evals/s', 'value': nprob * 1e3 / ms, 'unit': 'evals/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak' if mode == 'independent' else 'strong',
            'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD % (n, d),
                       'parallelism': par,
                       'l2': 'inputs exceed L2 (K is %.1f GiB)' % (8.0 * n * n / 2 ** 30),
                       'flops_per_eval_model': float(n) ** 3},
            'e2e': {'value': nprob * 1e3 / e2e_ms, 'unit': 'evals/s',
                    'h2d_bytes_per_step': int(Xp.numel() * 8 + Yp.numel() * 8),
                    'd2h_bytes_per_step': int(8 * (1 + nparam))},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'kernel': 'gemm_nt_tma_kernel (FP64 DMMA, TMA-staged operands)',
                         'achieved': ach, 'peak': pk, 'unit': 'TFLOP/s', 'frac': ach / pk if pk else None,
                         'traffic': traffic, 'traffic_of': traffic_note,
                         'peak_source': peak['source'], 'peak_burst': peak['burst'],
                         'gemm_share_of_step': gemm_ms / (prof_ms * prof_steps) if prof_ms > 0 else None,
                         'scope': ('rank 0, separate profiled pass of %d step(s) after the timed region%s' % (
                             prof_steps, ' (overlapping streams: per-launch event time includes queueing '
                             'behind the other streams)' if world > 1 else ''))
                         if not profile_in_region else 'the GPU, events inside the timed region',
                         'step_tflops_vs_n3': step_tf,
                         'step_frac_of_aggregate_peak': step_tf / (pk * world)},
        }
        if parity is not None:
            worst = max(parity.values())
            line['parity_rel_err'] = worst
            line['parity'] = {'against': 'tests/golden/gpr_large_scalars.json (CPU oracle, LAPACK analytic gradient)',
                              'tolerance': PARITY_RTOL, 'ok': bool(worst < PARITY_RTOL), 'errors': parity}
            if not worst < PARITY_RTOL:
                rc = 3
        else:
            line['parity_rel_err'] = None
        if phases is not None:
            line['phases_ms'] = phases
            fac = phases.get('factor(lookahead)')
            if fac:
                tf = float(n) ** 3 / 3.0 / (fac * 1e-3) / 1e12
                line['potrf'] = {'n': n, 'ms': fac, 'tflops': tf, 'frac_of_fp64_peak': tf / (peak['burst'] * world),
                                 'flops': 'N^3/3', 'scope': 'distributed factorisation phase incl. its collectives '
                                 '(max over ranks), against %d x the FP64 peak measured in this process' % world}
        if world == 1:
            line['potrf'] = potrf_metric(gpf, model, dev, n, peak)
        if secondary is not None:
            line['secondary'] = secondary
        if world == 1 and not args.no_cpu_baseline:
            ns = min(args.cpu_n, n)
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            times = oracle_eval(ns, d, 2)[1:]
            t = float(np.mean(times))
            scale = cpu_scale(n, ns)
            line['cpu_baseline'] = {'value': 1.0 / (t * scale), 'unit': 'evals/s', 'cores': cores, 'kind': 'port',
                                    'sample': cpu_sample_note(ns, d, t, cores, n)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == '__main__':
    main()
