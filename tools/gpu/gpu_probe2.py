import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops
from oracle import emap_oracle as O
from tests.conftest import load_golden
dev = "cuda"
for k in (10, 16, 32):
    g = load_golden(f"sample_pdf_k{k}")
    z_new, inds = ops.sample_pdf_det(g["bins"].to(dev), g["weights"].to(dev), k)
    bad = (inds.cpu() != g["inds"]).nonzero()
    print("k", k, "mismatches", len(bad))
    w = g["weights"] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros(48, 1), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / k, 1 - 0.5 / k, steps=k)
    for b, j in bad[:6].tolist():
        i = int(g["inds"][b, j]); i2 = int(inds[b, j])
        print("  ray", b, "j", j, "ref ind", i, "got", i2, "u", u[j].item(), "cdf around", cdf[b, max(i-2,0):i+2].tolist())
g = load_golden("upsample_init_64_50_5")
o, d = g["rays_o"].to(dev), g["rays_d"].to(dev)
sd = torch.tensor([float(g["sample_dist"])], device=dev)
k = 10
u = torch.linspace(0.5 / k, 1 - 0.5 / k, steps=k).to(dev)
for i in range(5):
    zi = g["z0"] if i == 0 else g[f"z{i}"]
    ui = g["udf0"] if i == 0 else g[f"udf{i}"]
    inv_s, beta, gamma = O.upsample_schedule(i, 5)
    _, _, z_new, inds, w = ops.upsample_step(o, d, zi.to(dev), ui.to(dev), None, None, u, k, sd, inv_s, beta, gamma, want_inds=True, want_weights=True)
    zr, ir, wr = O.up_sample_unbias(g["rays_o"], g["rays_d"], zi, ui, float(g["sample_dist"]), k, inv_s, beta, gamma, return_aux=True)
    dw = (w.cpu() - wr).abs()
    print("step", i, "w err", dw.max().item(), "wmax", wr.abs().max().item(), "flips", int((inds.cpu() != ir).sum()), "z err", (z_new.cpu() - torch.sort(g[f"z_new{i}"], -1)[0]).abs().max().item())
    b, j = divmod(int(dw.argmax()), dw.shape[1])
    print("   worst at ray", b, "interval", j, "got", w[b, j].item(), "ref", wr[b, j].item(), "neighbors ref", wr[b, max(j-2,0):j+3].tolist())
