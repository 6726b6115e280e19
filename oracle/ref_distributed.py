"""CPU restatement of the reference's N-rank gradient synchronisation, emulated in ONE process.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``distributed.py:24-61`` (TimZaman/dotaclient): after backward, for every parameter in
``named_parameters()`` order, all-reduce(SUM) a 0/1 has-grad flag, skip if nobody has a gradient, else
all-reduce(SUM) the gradient (zeros where this rank has none) and divide by the count.  Each rank then
runs its own grad-norm / clip / Adam (``optimizer.py:674-681``).

This module keeps one ``RefOptimizer`` per emulated rank (same initial weights: ``sync_parameters``
broadcasts rank 0's, ``distributed.py:71-74``) and performs the sums directly, which is what gloo's
all-reduce computes.  One deliberate difference, mirrored by the product and documented in DESIGN.md:
the averaged gradient is installed on EVERY rank, including ranks that had none locally -- the
reference writes it into a temporary there (``distributed.py:50-56``) and replicas drift apart.
"""
import torch

from . import ref_optimizer as RO


def train_ranks(optimizers, shards):
    """One synchronous data-parallel ``train()`` step.  ``optimizers[r]`` trains on ``shards[r]``.
    Returns the per-rank (losses, entropies, grad_norms) like ``DotaOptimizer.train``."""
    world = len(optimizers)
    results = []
    for opt, xs in zip(optimizers, shards):
        (loss, p_loss, e_loss, v_loss, ents), _, _ = opt.loss_only(xs)
        opt.optimizer.zero_grad()
        loss.backward()
        results.append(({"loss": loss, "policy_loss": p_loss, "entropy_loss": e_loss, "value_loss": v_loss}, ents))
    names = [n for n, _ in optimizers[0].policy_base.named_parameters()]
    params = [dict(o.policy_base.named_parameters()) for o in optimizers]
    for n in names:
        have = [params[r][n].grad is not None for r in range(world)]
        count = sum(have)                                       # distributed.py:36-37
        if count == 0:                                          # distributed.py:40-42
            continue
        total = sum(params[r][n].grad if have[r] else torch.zeros_like(params[r][n]) for r in range(world))
        avg = total / count                                     # distributed.py:56-57
        for r in range(world):
            params[r][n].grad = avg.clone()
    out = []
    for r, opt in enumerate(optimizers):
        ps = list(opt.policy_base.parameters())
        gn = RO.mean_gradient_norm(ps)
        torch.nn.utils.clip_grad_norm_(ps, RO.MAX_GRAD_NORM)
        gn_clipped = RO.mean_gradient_norm(ps)
        opt.optimizer.step()
        out.append((results[r][0], results[r][1], {"unclipped": gn, "clipped": gn_clipped}))
    return out
