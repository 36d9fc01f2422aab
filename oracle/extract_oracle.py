"""CPU oracle (TEST INFRASTRUCTURE ONLY) for SURVEY §8(f) rows 1-2: the grid UDF query of edge
extraction and the deterministic half of the ray sampler.

Plain torch-CPU restatement of the reference algorithms; every function cites the reference lines
it follows.  Only ``tests/`` may import this module -- the product path (``emap_b200/``) never does.
Pinned against outputs of the reference itself: ``tests/golden/make_golden.py`` runs the reference's
own ``get_udf_normals_grid`` / ``get_udf_normals_slow`` / ``gen_random_rays_patches_at`` (the latter
exec'd from its source file with a stand-in ``self``) and stores their outputs as
``tests/golden/extract_*.npz`` / ``raygen.npz``; ``tests/test_oracle_golden.py`` checks this file
against them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def grid_coordinates(N: int) -> torch.Tensor:
    """[N^3, 3] fp32 grid of [-1,1]^3, index order and rounding of
    src/edge_extraction/extract_pointcloud.py:36-54 (x slowest, z fastest; idx*voxel + (-1) in fp32)."""
    idx = torch.arange(0, N ** 3, 1, out=torch.LongTensor())
    s = torch.zeros(N ** 3, 3)
    s[:, 2] = idx % N
    s[:, 1] = torch.div(idx, N, rounding_mode="floor") % N
    s[:, 0] = torch.div(torch.div(idx, N, rounding_mode="floor"), N, rounding_mode="floor") % N
    voxel_size = 2.0 / (N - 1)
    for c in range(3):
        s[:, c] = (s[:, c] * voxel_size) + (-1)
    return s


def null_direction(grad_ld: torch.Tensor) -> torch.Tensor:
    """[M,S,3] -> [M,3]: F.normalize(svd(grad_ld).Vh[:, -1, :])  (extract_pointcloud.py:86-89)."""
    _, _, vh = torch.linalg.svd(grad_ld)
    return F.normalize(vh[:, -1, :].reshape(-1, 3), dim=1)


def udf_normals_grid(func, func_grad, N, udf_threshold, is_linedirection=False, sampling_N=50,
                     sampling_delta=0.005, max_batch=2 ** 12, offsets=None):
    """extract_pointcloud.py:5-95.  ``offsets`` ([M,sampling_N,3], standard normal) replaces the
    reference's in-place ``torch.randn`` draw so that two implementations can share it."""
    samples = torch.zeros(N ** 3, 12)
    samples[:, :3] = grid_coordinates(N)
    voxel_size = 2.0 / (N - 1)
    n = N ** 3
    for head in range(0, n, max_batch):
        tail = min(head + max_batch, n)
        df, _, _ = func(samples[head:tail, :3].clone())
        samples[head:tail, 3:4] = df.detach()
    norm_idx = torch.where(samples[:, 3] < udf_threshold)[0]
    for head in range(0, len(norm_idx), max_batch):
        tail = min(head + max_batch, len(norm_idx))
        sub = norm_idx[head:tail]
        pts = samples[sub, :3].clone()
        grad = func_grad(pts).detach()
        samples[sub, 4:7] = -F.normalize(grad, dim=1)[:, 0]
        if is_linedirection:
            off = offsets[head:tail] if offsets is not None else torch.randn(pts.shape[0], sampling_N, 3)
            pts_ld = pts.unsqueeze(1) + sampling_delta * off
            g = func_grad(pts_ld.reshape(-1, 3)).detach().reshape(pts.shape[0], sampling_N, 3)
            samples[sub, 8:11] = null_direction(g)
    return (samples[:, 3].reshape(N, N, N), samples[:, 8:11].reshape(N, N, N, 3),
            samples[:, 4:7].reshape(N, N, N, 3), samples, torch.tensor(voxel_size))


def udf_normals_slow(func, func_grad, xyz, is_linedirection, sampling_N=50, sampling_delta=0.005,
                     max_batch=2 ** 12, offsets=None):
    """extract_pointcloud.py:98-193 (df, -normalised gradient, optional line direction at xyz)."""
    n = xyz.shape[0]
    samples = torch.cat([xyz, torch.zeros(n, 10)], dim=-1)
    for head in range(0, n, max_batch):
        tail = min(head + max_batch, n)
        pts = samples[head:tail, 0:3].clone()
        with torch.no_grad():
            df, _, _ = func(pts)
        samples[head:tail, 3] = df.squeeze(-1).detach()
        grad = func_grad(pts).detach()[:, 0]
        samples[head:tail, 4:7] = -F.normalize(grad, dim=1)
        if is_linedirection:
            off = offsets[head:tail] if offsets is not None else torch.randn(pts.shape[0], sampling_N, 3)
            pts_ld = (pts.unsqueeze(1) + sampling_delta * off).reshape(-1, 3)
            g = func_grad(pts_ld.float()).detach()[:, 0].reshape(pts.shape[0], -1, 3)
            samples[head:tail, 7:10] = null_direction(g)
    return samples[:, 3], samples[:, 4:7], samples[:, 7:10], samples


def rays_from_pixels(pixels_x, pixels_y, edges_img, intrinsics_inv, pose, H, W):
    """Deterministic half of gen_random_rays_patches_at (src/dataset/dataset.py:268-305): everything
    after the pixel draw.  edges_img [H,W,1], intrinsics_inv [4,4], pose [4,4]."""
    ndc_u = 2 * pixels_x / (W - 1) - 1
    ndc_v = 2 * pixels_y / (H - 1) - 1
    rays_ndc_uv = torch.stack([ndc_u, ndc_v], dim=-1).view(-1, 2).float()
    edge = edges_img[(pixels_y, pixels_x)]
    p = torch.stack([pixels_x, pixels_y, torch.ones_like(pixels_y)], dim=-1).float()
    p = torch.matmul(intrinsics_inv[None, :3, :3], p[:, :, None]).squeeze()
    rays_v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    depth_scale = rays_v[:, 2:]
    rays_v = torch.matmul(pose[None, :3, :3], rays_v[:, :, None]).squeeze()
    rays_o = pose[None, :3, 3].expand(rays_v.shape)
    return {"rays_o": rays_o, "rays_v": rays_v, "edge": edge, "rays_ndc_uv": rays_ndc_uv,
            "rays_norm_XYZ_cam": p, "depth_scale": depth_scale}
