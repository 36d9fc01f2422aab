"""Multi-GPU data parallelism for the render path: rays shard, weights replicate.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The forward has NO collective
(every per-ray stage is independent; each rank renders its contiguous slice of the ray batch).  The
only exchange step of a training iteration is ONE all-reduce of a single flat fp32 buffer holding all
462,985 gradients (UDF MLP 462,980 + variance 2 + beta/gamma/zeta 3 = 1.85 MB): latency-bound, so the
win is not launching 32 per-parameter collectives.  Batch-level ratios (the two eikonal means) are made
exact across shards by all-reducing their two denominators (8 bytes) -- see ``global_denominators``.

Everything here is backend-agnostic torch.distributed code: the CPU test-suite exercises it with
``gloo`` and world_size 2; on the B200 box the same code runs over NCCL.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of n rays; the first (n % world) ranks get one extra ray."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(tensors: Sequence[torch.Tensor], rank: Optional[int] = None,
               world_size: Optional[int] = None) -> List[torch.Tensor]:
    """Slice every [B, ...] tensor of a ray batch (rays_o, rays_d, true_edge, depth_scale, t_rand...)
    to this rank's contiguous shard.  Draw per-ray randomness for the FULL batch first, then shard,
    so that concatenating the shards' forward outputs equals the single-GPU result bit for bit."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world_size)
    return [t[lo:hi].contiguous() for t in tensors]


GRAD_ARENA_SLACK = 32     # spare floats ops.udf_backward leaves behind its flat gradient for foreign gradients


class FlatGradAllReduce:
    """Average the gradients of ``params`` across ranks with ONE all-reduce of one flat buffer.

    The MLP backward (ops.udf_backward) returns all 462,980 gradients as ONE flat buffer and autograd hands the
    per-parameter views of it to ``p.grad`` without copying, so that buffer (the "arena") is all-reduced IN
    PLACE: the few gradients that live elsewhere (variance, beta, gamma, ...: 5 floats) are copied into the
    arena's spare tail with one batched copy and back with another.  Parameters whose gradients do not share
    an arena (any other model) fall back to packing into an own flat buffer with batched copies.  Either way:
    one collective, no per-parameter launches."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self._flat: Optional[torch.Tensor] = None

    # -- layout -------------------------------------------------------------------------------
    def _arena(self, grads: List[Optional[torch.Tensor]]):
        """(k, n): grads[0..k) tile one storage back to back (n elements from its storage offset 0..)."""
        g0 = grads[0] if grads else None
        if g0 is None or not g0.is_contiguous():
            return 0, 0
        base = g0.untyped_storage().data_ptr()
        start = g0.storage_offset()
        k, n = 0, 0
        for g in grads:
            if (g is None or not g.is_contiguous() or g.dtype != g0.dtype
                    or g.untyped_storage().data_ptr() != base or g.storage_offset() != start + n):
                break
            k += 1
            n += g.numel()
        return (k, n) if k > 1 else (0, 0)

    def flat_grads(self) -> torch.Tensor:
        """All gradients packed into an own flat buffer (missing ones as zeros) -- the fallback layout."""
        ps = self.params
        dev, dt = ps[0].device, ps[0].dtype
        if self._flat is None or self._flat.device != dev:
            self._flat = torch.zeros(self.numel, dtype=dt, device=dev)
        views, srcs, off = [], [], 0
        for p in ps:
            n = p.numel()
            if p.grad is None:
                self._flat[off:off + n].zero_()
            else:
                views.append(self._flat[off:off + n].view_as(p))
                srcs.append(p.grad)
            off += n
        if views:
            torch._foreach_copy_(views, srcs)
        return self._flat

    def scatter_(self, flat: torch.Tensor) -> None:
        dsts, srcs, off = [], [], 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                dsts.append(p.grad)
                srcs.append(flat[off:off + n].view_as(p))
            elif p.requires_grad:
                p.grad = flat[off:off + n].view_as(p).clone()
            off += n
        if dsts:
            torch._foreach_copy_(dsts, srcs)

    def _mean_all_reduce(self, buf: torch.Tensor, w: int) -> None:
        """mean over ranks, in place: NCCL averages inside the collective (one launch less); gloo has no AVG"""
        if buf.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(w)

    # -- the exchange step -------------------------------------------------------------------
    def allreduce_(self, local_weight: float = 1.0) -> None:
        """SUM over ranks of local_weight * grad, divided by the world size.  local_weight = 1 is the plain
        average (equal shards); for unequal ray shards pass B_local * W / B_global (``shard_weight``) so that
        per-shard means (mse, sparse error) combine to the mean over the whole batch."""
        _, w = world()
        if w == 1 and local_weight == 1.0:
            return
        grads = [p.grad for p in self.params]
        k, n = self._arena(grads)
        if k:
            g0 = grads[0]
            rest = [g for g in grads[k:] if g is not None]
            extra = sum(g.numel() for g in rest)
            cap = g0.untyped_storage().nbytes() // g0.element_size() - g0.storage_offset()
            if n + extra <= cap and all(g.dtype == g0.dtype and g.device == g0.device for g in rest):
                buf = torch.empty(0, dtype=g0.dtype, device=g0.device).set_(
                    g0.untyped_storage(), g0.storage_offset(), (n + extra,))
                tails, off = [], n
                for g in rest:
                    tails.append(buf[off:off + g.numel()].view_as(g))
                    off += g.numel()
                if rest:
                    torch._foreach_copy_(tails, rest)
                if local_weight != 1.0:
                    buf.mul_(float(local_weight))
                if w > 1:
                    self._mean_all_reduce(buf, w)
                if rest:
                    torch._foreach_copy_(rest, tails)
                return
        flat = self.flat_grads()
        if local_weight != 1.0:
            flat.mul_(float(local_weight))
        if w > 1:
            self._mean_all_reduce(flat, w)
        self.scatter_(flat)


def shard_weight(local_rays: int, global_rays: int, world_size: Optional[int] = None) -> float:
    """Weight of this rank's gradient for exact whole-batch means with unequal ray shards: per-shard means
    (mse, sparse_error: / B_local) averaged over W ranks equal the batch mean iff each is weighted by
    B_local * W / B_global (1.0 for equal shards)."""
    _, w = world()
    w = w if world_size is None else world_size
    return float(local_rays) * w / float(global_rays)


def global_denominators(local_sums: torch.Tensor, group=None) -> torch.Tensor:
    """SUM-all-reduce of the mask sums (sum relax_inside_sphere, sum near_surface) so that the eikonal
    means use the denominators of the whole batch (udf_renderer_blending.py:618-625)."""
    _, w = world()
    out = local_sums.clone()
    if w > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def globalize_eikonal(reduced: torch.Tensor, group=None) -> torch.Tensor:
    """Turn the per-shard eikonal means into terms whose rank-average is the GLOBAL-batch mean.

    ``reduced`` = [gerr, gerr_near, sparse, sum_relax, sum_near] of this rank's shard
    (emap_render_core_fwd).  gerr = N_local/(D_local+1e-5).  The reference's semantics on the full
    batch is N_global/(D_global+1e-5) (udf_renderer_blending.py:618-625).  With gradient averaging over
    W ranks, each rank must contribute  W*N_local/(D_global+1e-5):  their mean is exactly the global
    value and so is the averaged gradient.  Returns a new 5-vector whose entries 3,4 are the effective
    denominators (D_global+1e-5)/W - 1e-5 to be used by the backward kernel.  One 2-float all-reduce."""
    _, w = world()
    if w == 1:
        return reduced
    local_d = reduced[3:5]
    numer = reduced[0:2] * (local_d + 1e-5)
    glob_d = global_denominators(local_d, group)
    eff_d = (glob_d + 1e-5) / w - 1e-5
    out = reduced.clone()
    out[0:2] = numer / (eff_d + 1e-5)
    out[3:5] = eff_d
    return out
