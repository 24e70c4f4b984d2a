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
"""
import torch

_STATE = {'active': False, 'group': None, 'block': 512, 'lookahead': True, 'own_group': None}


def init(group=None, block=512, backend='nccl', device=None, lookahead=True):
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
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        ranks = list(range(dist.get_world_size()))
        _STATE['own_group'] = {name: dist.new_group(ranks=ranks, backend='nccl', pg_options=opts)
                               for name in ('chain', 'tb', 'gather')}
    if group is None:
        group = _STATE.get('own_group')
    _STATE.update(active=True, group=group, block=int(block), lookahead=bool(lookahead))


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
