"""Bodies of tests/test_gpu_graph.py, run in a CHILD process (python -m tests.graph_cases <name>): a fault inside a
CUDA-graph replay is sticky for the whole process, and must not be able to take the rest of the GPU suite down."""
import copy
import sys

import torch

from tests.helpers import maxdiff
from tests.test_gpu_render import build

dev = "cuda"


def _problem(B=192):
    from oracle import emap_oracle as O
    o, d = O.synthetic_rays(B, seed=5)
    near, far, ds = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0), torch.ones(B, 1)
    te = torch.rand(B, 1, generator=torch.Generator().manual_seed(3))
    return [t.to(dev) for t in (o, d, near, far, ds, te)]


def _iteration(r, opt):
    def fn(o, d, near, far, ds, te):
        out = r.render(o, d, near, far, ds, cos_anneal_ratio=1.0, perturb_overwrite=0, flip_saturation=0.9)
        loss = (torch.nn.functional.mse_loss(out["edge"], te) + 0.01 * out["gradient_error_near_surface"]
                + 0.1 * out["gradient_error"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return (loss.detach().reshape(1),)
    return fn


def _setup():
    net, var, beta, r = build(10, True, n_samples=64, n_importance=64, up_sample_steps=4)
    params = list(net.parameters()) + list(var.parameters()) + list(beta.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, capturable=True)
    return net, var, beta, r, params, opt


def graphed_train_iterations_match_eager():
    from emap_b200.graph import GraphedStep
    batch = _problem()
    # eager: 3 iterations (the warm-ups GraphedStep runs before capturing; the capture itself executes nothing) + 3 more
    net, var, beta, r, params, opt = _setup()
    init = copy.deepcopy([p.detach().clone() for p in params])
    fn = _iteration(r, opt)
    eager = [float(fn(*batch)[0]) for _ in range(6)]
    eager_params = [p.detach().clone() for p in params]
    r.check_numerics()

    net, var, beta, r, params, opt = _setup()
    for p, p0 in zip(params, init):
        assert torch.equal(p.detach(), p0)
    step = GraphedStep(_iteration(r, opt), batch, warmup=3, refold=[net])
    graphed = [float(step(*batch)[0]) for _ in range(3)]
    torch.cuda.synchronize()
    r.check_numerics()
    for a, b in zip(eager[3:], graphed):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (eager, graphed)
    for p, q in zip(params, eager_params):
        assert maxdiff(p.detach(), q) <= 1e-6
    # the parameters changed behind the fold cache's back: an eager call after the replays must see them
    with torch.no_grad():
        out = r.render(*batch[:5], cos_anneal_ratio=1.0, perturb_overwrite=0, flip_saturation=0.9)
    net.invalidate()
    with torch.no_grad():
        out2 = r.render(*batch[:5], cos_anneal_ratio=1.0, perturb_overwrite=0, flip_saturation=0.9)
    assert torch.equal(out["edge"], out2["edge"])


def graphed_inference_matches_eager_and_rejects_host_draws():
    from emap_b200.graph import GraphedStep
    o, d, near, far, ds, _ = _problem(160)
    net, var, beta, r = build(10, True, n_samples=64, n_importance=50, up_sample_steps=5)

    def infer(o, d, near, far, ds):
        with torch.no_grad():
            out = r.render(o, d, near, far, ds, cos_anneal_ratio=1.0, perturb_overwrite=0, flip_saturation=0.9)
        return out["edge"], out["depth"], out["weights"]

    ref = [t.clone() for t in infer(o, d, near, far, ds)]
    step = GraphedStep(infer, [o, d, near, far, ds], refold=[net])
    o2 = o + 0.01                                                       # new inputs through the static buffers
    got = [t.clone() for t in step(o2, d, near, far, ds)]
    ref2 = infer(o2, d, near, far, ds)
    for a, b in zip(got, ref2):
        assert torch.equal(a, b)
    got = step(o, d, near, far, ds)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)

    # perturb > 0 draws from the host generator unless perturb_on_device is set: refused (already in the warm-up)
    def infer_perturbed(o, d, near, far, ds):
        with torch.no_grad():
            return (r.render(o, d, near, far, ds, cos_anneal_ratio=1.0, flip_saturation=0.9)["edge"],)
    try:
        GraphedStep(infer_perturbed, [o, d, near, far, ds])
        raise AssertionError("a host-side draw was accepted")
    except RuntimeError as e:
        assert "perturb_on_device" in str(e)
    torch.cuda.synchronize()
    r.perturb_on_device = True
    step = GraphedStep(infer_perturbed, [o, d, near, far, ds], refold=[net])
    a = step(o, d, near, far, ds)[0].clone()
    b = step(o, d, near, far, ds)[0].clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b)            # fresh offsets on every replay


if __name__ == "__main__":
    {"train": graphed_train_iterations_match_eager,
     "infer": graphed_inference_matches_eager_and_rejects_host_draws}[sys.argv[1]]()
    torch.cuda.synchronize()
    print("ok")
