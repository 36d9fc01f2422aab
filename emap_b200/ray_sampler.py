"""Device-resident replacement for the per-iteration ray sampler of ``src/dataset/dataset.py``
(``Dataset.gen_random_rays_patches_at``, :221-305).

The reference does, on the HOST and per training iteration: ``np.ones_like`` over H*W, Python
``random.choices`` over H*W weights, a full-image ``meshgrid`` and ~10 tiny tensor ops + H2D copies.
Here the images/cameras live in HBM, the pixel -> ray arithmetic is one fused kernel
(``emap_rays_from_pixels``) and the importance draw is O(batch):

* uniform half / ``importance_sample=False``: ``torch.randint`` on the global CPU generator exactly like
  the reference (same seed -> same pixels), then one H2D copy of the indices;
* weighted half: the reference's weights take only two values (edge pixels 1-rho, others rho, rho = mean
  edge value), so a weighted draw over H*W is a Bernoulli choice of the class followed by a uniform
  draw inside the class (per-image index lists are built once and cached on the device).  Same
  distribution; the random STREAM differs from Python's ``random.choices`` (stated deviation).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


class RaySampler:
    def __init__(self, edges: torch.Tensor, intrinsics_all_inv: torch.Tensor, pose_all: torch.Tensor,
                 device="cuda", intrinsics_all: torch.Tensor = None):
        """edges [n_img,H,W,1] in [0,1]; intrinsics_all_inv, pose_all [n_img,4,4] (Dataset attributes)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("emap_b200.RaySampler runs on a CUDA device (no CPU path)")
        self.edges = edges.to(self.device, torch.float32).contiguous()
        self.n_images, self.H, self.W = edges.shape[0], edges.shape[1], edges.shape[2]
        self.image_pixels = self.H * self.W
        self.intrinsics_all_inv = intrinsics_all_inv.cpu().float()
        self.intrinsics_all = None if intrinsics_all is None else intrinsics_all.cpu().float()
        self.pose_all = pose_all.cpu().float()
        self._classes: Dict[int, tuple] = {}

    def _class_lists(self, img_idx):
        c = self._classes.get(img_idx)
        if c is None:
            img = self.edges[img_idx].reshape(-1)
            rho = float(img.mean())
            is_edge = img > 0.1
            idx_e = torch.where(is_edge)[0]
            idx_b = torch.where(~is_edge)[0]
            w_e, w_b = idx_e.numel() * (1.0 - rho), idx_b.numel() * rho
            p_edge = w_e / (w_e + w_b) if (w_e + w_b) > 0 else 0.0
            c = (idx_e, idx_b, p_edge)
            self._classes[img_idx] = c
        return c

    def gen_random_rays_patches_at(self, img_idx, batch_size, importance_sample=False, generator=None):
        if not importance_sample:
            px = torch.randint(low=0, high=self.W, size=[batch_size]).to(self.device)
            py = torch.randint(low=0, high=self.H, size=[batch_size]).to(self.device)
        else:
            h = batch_size // 2
            px1 = torch.randint(low=0, high=self.W, size=[h]).to(self.device)
            py1 = torch.randint(low=0, high=self.H, size=[h]).to(self.device)
            idx_e, idx_b, p_edge = self._class_lists(int(img_idx))
            u = torch.rand(h, device=self.device, generator=generator)
            r = torch.rand(h, device=self.device, generator=generator)
            pick_e = (u < p_edge) & (idx_e.numel() > 0)
            ie = idx_e[(r * max(idx_e.numel(), 1)).long().clamp_(max=max(idx_e.numel() - 1, 0))] \
                if idx_e.numel() else torch.zeros(h, dtype=torch.int64, device=self.device)
            ib = idx_b[(r * max(idx_b.numel(), 1)).long().clamp_(max=max(idx_b.numel() - 1, 0))] \
                if idx_b.numel() else ie
            flat = torch.where(pick_e, ie, ib)
            px = torch.cat([px1, flat % self.W])
            py = torch.cat([py1, torch.div(flat, self.W, rounding_mode="floor")])
        rays = ops.rays_from_pixels(px, py, self.edges[img_idx, :, :, 0], self.intrinsics_all_inv[img_idx],
                                    self.pose_all[img_idx])
        return {
            "rays": {"rays_o": rays["rays_o"], "rays_v": rays["rays_v"], "edge": rays["edge"]},
            "pose": self.pose_all[img_idx],
            "intrinsics": None if self.intrinsics_all is None else self.intrinsics_all[img_idx],
            "rays_ndc_uv": rays["rays_ndc_uv"],
            "rays_norm_XYZ_cam": rays["rays_norm_XYZ_cam"],
            "depth_scale": rays["depth_scale"],
        }
