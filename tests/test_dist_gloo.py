"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: ray sharding, single flat-buffer gradient
all-reduce, global denominators.  Same code runs over NCCL on the GPU box."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emap_b200 import parallel


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B = 11
        rays = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
        edge = torch.arange(B, dtype=torch.float32).reshape(B, 1)
        (ro, te) = parallel.shard_rays([rays, edge])
        lo, hi = parallel.shard_bounds(B, rank, world)
        assert torch.equal(ro, rays[lo:hi]) and torch.equal(te, edge[lo:hi])
        # shard-concat == full batch
        gathered = [None] * world
        dist.all_gather_object(gathered, ro)
        assert torch.equal(torch.cat(gathered), rays)

        # flat gradient all-reduce: 32 parameters of odd shapes, one collective
        params = [torch.nn.Parameter(torch.zeros(s)) for s in [(256, 63), (256,), (256, 1), (1,), (193, 256), (5,)]]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
        params[3].grad = None                                   # a frozen / unused parameter
        red = parallel.FlatGradAllReduce(params)
        calls = {"n": 0}
        orig = dist.all_reduce

        def counting(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)
        dist.all_reduce = counting
        red.allreduce_()
        dist.all_reduce = orig
        assert calls["n"] == 1
        mean_rank = sum(r + 1 for r in range(world)) / world
        for i, p in enumerate(params):
            if i == 3:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0
            else:
                assert torch.allclose(p.grad, torch.full_like(p, mean_rank * (i + 1)))

        # the production layout: the MLP backward hands out views of ONE flat buffer (with a spare tail) -- it is
        # all-reduced in place, the foreign gradients (scalar networks) ride in its tail; still one collective
        arena = torch.empty(20 + parallel.GRAD_ARENA_SLACK)[:20]
        arena.copy_(torch.arange(20.0) * (rank + 1))
        ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(1)),
              torch.nn.Parameter(torch.zeros(2))]
        ps[0].grad, ps[1].grad = arena[0:12].view(4, 3), arena[12:20]
        ps[2].grad, ps[3].grad = torch.tensor([7.0 * (rank + 1)]), torch.tensor([1.0, 2.0]) * (rank + 1)
        ptr0 = ps[0].grad.data_ptr()
        red2 = parallel.FlatGradAllReduce(ps)
        calls["n"] = 0
        dist.all_reduce = counting
        red2.allreduce_()
        dist.all_reduce = orig
        assert calls["n"] == 1
        assert ps[0].grad.data_ptr() == ptr0                    # reduced in place, no re-pack
        assert torch.allclose(torch.cat([ps[0].grad.reshape(-1), ps[1].grad]), torch.arange(20.0) * mean_rank)
        assert torch.allclose(ps[2].grad, torch.tensor([7.0 * mean_rank]))
        assert torch.allclose(ps[3].grad, torch.tensor([1.0, 2.0]) * mean_rank)
        # unequal ray shards: weights B_local * W / B_global make per-shard means combine to the batch mean
        lo_, hi_ = parallel.shard_bounds(11, rank, world)
        wgt = parallel.shard_weight(hi_ - lo_, 11)
        vals = torch.arange(11.0)
        pp = torch.nn.Parameter(torch.zeros(1))
        pp.grad = vals[lo_:hi_].mean().reshape(1)               # a per-shard mean
        parallel.FlatGradAllReduce([pp]).allreduce_(local_weight=wgt)
        assert torch.allclose(pp.grad, vals.mean().reshape(1))

        den = parallel.global_denominators(torch.tensor([float(rank + 1), 10.0 * (rank + 1)]))
        assert torch.allclose(den, torch.tensor([3.0, 30.0]))

        # exact full-batch eikonal means from per-shard ratios: mean over ranks == global ratio
        numer = torch.tensor([2.0 + rank, 5.0 + 3 * rank])
        dloc = torch.tensor([10.0 + 4 * rank, 3.0 + rank])
        red = torch.cat([numer / (dloc + 1e-5), torch.tensor([0.5]), dloc])
        g = parallel.globalize_eikonal(red)
        acc = g[0:2].clone()
        dist.all_reduce(acc)
        acc /= world
        n_all = torch.tensor([2.0 + 3.0, 5.0 + 8.0])
        d_all = torch.tensor([10.0 + 14.0, 3.0 + 4.0])
        assert torch.allclose(acc, n_all / (d_all + 1e-5), rtol=1e-6)
        # the backward divides the cotangent by (entry + 1e-5): must equal W / (D_global + 1e-5)
        assert torch.allclose(1.0 / (g[3:5] + 1e-5), world / (d_all + 1e-5), rtol=1e-6)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_bounds_cover():
    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
