"""Multi-GPU execution: one process per GPU (torchrun-style launch), torch.distributed over
NCCL / NVLink for the plumbing.  The reference has no distributed code at all (SURVEY.md
section 2.2); this module is what shards its hot path where it shards naturally:

  * GPR (models/gpr.py:55-72): `parallel.init()` makes `GPR.objective` run the block-row
    distributed Gram / Cholesky / inverse of `_backend/dist_gpr.py` across the ranks of the
    group -- every rank passes the same (replicated) X, Y and parameters and receives the
    same objective and gradients.
  * SVGP (models/svgp.py:108-125): `parallel.svgp_objective_and_grads(model, Xb, Yb)` shards the
    minibatch rows across ranks, evaluates the ELBO terms locally and all-reduces one flat
    gradient buffer.
  * SGPR (models/sgpr.py:121-156): `parallel.sgpr_objective_and_grads(model)` shards the DATA rows
    (the columns of Kuf): A A^T and A err are summed over ranks by one all-reduce of
    M^2 + M R + 3 doubles, the two small Cholesky factorisations are replicated.
  * GPR prediction (models/gpr.py:118-131): under `parallel.init()` `GPR.predict_f` factors
    distributed and shares out the test points (`_backend/dist_gpr.predict`).
"""
import torch

_STATE = {'active': False, 'group': None, 'block': 512, 'lookahead': True, 'own_group': None}


def init(group=None, block=512, backend='nccl', device=None, lookahead=True):
    # lookahead: True / 'v1' = the three-stream schedule (default: the faster one on 8 x B200, see
    # profiles/r02_dist_8gpu_ab.txt), 'v2' = the five-stream pipeline, False = plain right-looking order
    """Activate the distributed paths.  If torch.distributed is not initialised yet and the
    torchrun environment variables are present, initialise it (backend NCCL)."""
    import os
    import torch.distributed as dist
    if not dist.is_initialized() and 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', 1)) > 1:
        if device is None and torch.cuda.is_available():
            device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
            torch.cuda.set_device(device)
        dist.init_process_group(backend, device_id=device if backend == 'nccl' else None)
    if group is None and dist.is_initialized() and dist.get_world_size() > 1 \
            and dist.get_backend() == 'nccl' and _STATE.get('own_group') is None:
        # three dedicated communicators whose NCCL streams have HIGH priority, one per kind of
        # collective of the distributed Cholesky (diagonal-block broadcast / top-block broadcast /
        # panel all-gather): a process group runs its collectives in issue order on ONE stream, so
        # with a single communicator a 100 MB panel all-gather would sit in front of the 2 MB
        # broadcast the serial chain is waiting for
        # GPSLIM_NCCL_COMMS=1 puts all three kinds on ONE high-priority communicator (the round-1
        # arrangement); GPSLIM_NCCL_CTAS caps the CTAs NCCL may use per communicator (0 = NCCL's own
        # choice).  Both are measurement knobs: profiles/r02_dist_8gpu_ab.txt holds the A/B runs.
        ctas = [int(c) for c in os.environ.get('GPSLIM_NCCL_CTAS', '4,4,16').split(',')]
        ncomm = int(os.environ.get('GPSLIM_NCCL_COMMS', '3'))
        ranks = list(range(dist.get_world_size()))
        own = {}
        for name, c in zip(('chain', 'tb', 'gather'), ctas):
            if ncomm == 1 and own:
                own[name] = own['chain']
                continue
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            if c > 0 and ncomm != 1:
                opts.config.max_ctas = c
            own[name] = dist.new_group(ranks=ranks, backend='nccl', pg_options=opts)
        _STATE['own_group'] = own
    if group is None:
        group = _STATE.get('own_group')
    _STATE.update(active=True, group=group, block=int(block),
                  lookahead=lookahead if lookahead in ('v1', 'v2') else bool(lookahead))


def _pg():
    """the plain process group (all-reduces, world size, rank)."""
    g = _STATE['group']
    return g.get('gather') if isinstance(g, dict) else g


def shutdown():
    _STATE.update(active=False, group=None)


def active():
    return _STATE['active']


def group():
    return _STATE['group']


def block():
    return _STATE['block']


def lookahead():
    return _STATE['lookahead']


def world_size():
    import torch.distributed as dist
    return dist.get_world_size(_pg()) if dist.is_available() and dist.is_initialized() else 1


def rank():
    import torch.distributed as dist
    return dist.get_rank(_pg()) if dist.is_available() and dist.is_initialized() else 0


def svgp_objective_and_grads(model, Xb, Yb, params=None):
    """Data-parallel SVGP step: this rank evaluates the variational
    expectations of ITS rows of the minibatch (Xb, Yb are the rank-local shard; the global batch
    is their concatenation over ranks) with Kuu / chol(Kuu) / KL replicated, and ONE all-reduce
    of a flat buffer [objective, d objective / d every trainable tensor] follows.  Returns
    (objective, grads) identical on every rank and equal to the single-process values on the
    concatenated batch (models/svgp.py:108-125: ELBO = sum var_exp * num_data / B - KL)."""
    import torch.distributed as dist
    from .misc import to_tensor
    params = list(params) if params is not None else model.trainable_tensors
    Xb, Yb = to_tensor(Xb), to_tensor(Yb)
    world = world_size()
    nloc = torch.tensor([float(Xb.shape[0])], dtype=torch.float64, device=Xb.device)
    if world > 1:
        dist.all_reduce(nloc, group=_pg())
    btot = float(nloc)
    fmean, fvar = model._build_predict(Xb, full_cov=False)
    var_exp = model.likelihood.variational_expectations(fmean, fvar, Yb)
    scale = float(model.num_data) / btot
    # the KL term is replicated: every rank contributes KL / world so that the sum is exact
    local = -(var_exp.sum() * scale) + (model.build_prior_KL() - model.prior_tensor) / world
    grads = torch.autograd.grad(local, params, allow_unused=True)
    flat = torch.cat([local.detach().reshape(1)] +
                     [(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(grads, params)])
    if world > 1:
        dist.all_reduce(flat, group=_pg())
    out, o = [], 1
    for p in params:
        out.append(flat[o:o + p.numel()].view_as(p))
        o += p.numel()
    return flat[0], out


def sgpr_objective_and_grads(model, params=None):
    """Data-parallel SGPR objective (last row; reference models/sgpr.py:121-156).
    `model` is an SGPR whose X, Y are THIS RANK's rows of the data set (the global data set is their
    concatenation over ranks; Z, kernel and noise are replicated).  With A = L^-1 Kuf / sigma the
    bound needs A A^T (M x M), A err (M x R), sum err^2, sum Kdiag and N: all sums over data rows,
    i.e. over ranks -- one all-reduce of M^2 + M R + 3 doubles.  What follows (B = A A^T + I,
    chol(B), c, the bound) is replicated.

    Gradient: bound = h(S(theta), theta) with S = sum_r s_r(theta).  Every rank differentiates h
    w.r.t. the reduced S (a leaf) and w.r.t. theta directly (replicated, identical on all ranks),
    pulls dh/dS back through ITS s_r, and one all-reduce sums those pull-backs:
        d/d theta = dh/d theta|_S + sum_r (d s_r / d theta)^T dh/dS.
    Returns (objective, grads) identical on every rank and equal to the single-process values."""
    import math
    import torch.distributed as dist
    from ._backend import ops as _ops
    from ._settings import SETTINGS as settings
    params = list(params) if params is not None else model.trainable_tensors
    world = world_size()
    X, Y = model.X, model.Y
    M, R = len(model.feature), Y.shape[1]
    err = Y - model.mean_function(X)
    Kfu = model.feature.Kfu(model.kern, X)                                      # Kuf^T, this rank's rows
    L = _ops.cholesky(model.feature.Kuu(model.kern, jitter=settings.numerics.jitter_level))
    var = model.likelihood.variance
    sigma = torch.sqrt(var)
    At = _ops.trsm_rlt(Kfu, L) / sigma
    Att = _ops.t(At)
    pieces = [_ops.matmul_nt(Att, Att).reshape(-1),                             # A A^T   (sgpr.py:140-141)
              _ops.matmul_nt(Att, _ops.t(err)).reshape(-1),                     # A err   (:144)
              (err ** 2).sum().reshape(1), model.kern.Kdiag(X).sum().reshape(1),
              torch.full((1,), float(X.shape[0]), dtype=At.dtype, device=At.device)]
    local = torch.cat(pieces)
    S = local.detach().clone()
    if world > 1:
        dist.all_reduce(S, group=_pg())
    S.requires_grad_(True)
    AAT, Aerr = S[:M * M].view(M, M), S[M * M:M * M + M * R].view(M, R)
    e2, kd, ntot = S[M * M + M * R], S[M * M + M * R + 1], float(S[M * M + M * R + 2])
    B = AAT + torch.eye(M, dtype=AAT.dtype, device=AAT.device)
    LB = _ops.cholesky(B)
    c = _ops.solve_lower(LB, Aerr) / sigma
    bound = (-0.5 * ntot * R * math.log(2 * math.pi) - R * torch.log(torch.diagonal(LB)).sum()
             - 0.5 * ntot * R * torch.log(var) - 0.5 * e2 / var + 0.5 * (c ** 2).sum()
             - 0.5 * R * kd / var + 0.5 * R * torch.diagonal(AAT).sum())            # sgpr.py:147-154
    obj = -(bound + model.prior_tensor)
    g = torch.autograd.grad(obj, [S] + params, allow_unused=True, retain_graph=True)   # sigma / var are shared with `local`
    gS, gdir = g[0], g[1:]
    gloc = torch.autograd.grad(local, params, grad_outputs=gS, allow_unused=True)
    flat = torch.cat([(t if t is not None else torch.zeros_like(p)).reshape(-1) for t, p in zip(gloc, params)])
    if world > 1:
        dist.all_reduce(flat, group=_pg())
    out, o = [], 0
    for p, gd in zip(params, gdir):
        t = flat[o:o + p.numel()].view_as(p)
        out.append(t if gd is None else t + gd)
        o += p.numel()
    return obj.detach(), out

