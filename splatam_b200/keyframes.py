"""Keyframe selection by re-projection overlap -- the scheduler that picks which keyframes a mapping phase
optimises over (reference R/utils/keyframe_selection.py:10-95,
call site R/scripts/splatam.py:800-817).

Same algorithm and the same random-number consumption as the reference (one `torch.randint` over the valid depth
pixels with the global torch generator, one `np.random.permutation` over the overlapping keyframes), so a seeded
run picks the same keyframes.  Differences in form only: device-agnostic (the reference hard-codes `.cuda()`), and
all keyframes are projected by ONE batched matmul instead of a Python loop of ~10 small launches per keyframe.
"""
import numpy as np
import torch


def get_pointcloud(depth, intrinsics, w2c, sampled_indices):
    """World-space points of the sampled pixels (rows = (y, x)), points at the camera origin removed
    (R/utils/keyframe_selection.py:10-37)."""
    CX, CY, FX, FY = intrinsics[0][2], intrinsics[1][2], intrinsics[0][0], intrinsics[1][1]
    xx = (sampled_indices[:, 1] - CX) / FX
    yy = (sampled_indices[:, 0] - CY) / FY
    depth_z = depth[0, sampled_indices[:, 0], sampled_indices[:, 1]]
    pts_cam = torch.stack((xx * depth_z, yy * depth_z, depth_z), dim=-1)
    pts4 = torch.cat([pts_cam, torch.ones_like(pts_cam[:, :1])], dim=1)
    pts = (torch.inverse(w2c) @ pts4.T).T[:, :3]
    # the reference drops every point that rounds (4 decimals, absolute value) to the origin -- and, through its
    # unique/counts construction, every point whose rounded |coordinates| occur more than once; both are kept here
    A = torch.abs(torch.round(pts, decimals=4))
    B = torch.zeros((1, 3), device=pts.device, dtype=pts.dtype)
    _, idx, counts = torch.cat([A, B], dim=0).unique(dim=0, return_inverse=True, return_counts=True)
    invalid = torch.isin(idx, torch.where(counts.gt(1))[0])[:len(A)]
    return pts[~invalid]


def overlap_fractions(pts, est_w2cs, intrinsics, width, height, edge=20):
    """Fraction of `pts` [N,3] that project inside each keyframe's image (margin `edge`), est_w2cs: [K,4,4]."""
    pts4 = torch.cat([pts, torch.ones_like(pts[:, :1])], dim=1)                      # [N,4]
    cam = torch.matmul(est_w2cs, pts4.T.unsqueeze(0))[:, :3, :]                      # [K,3,N]
    pix = torch.matmul(intrinsics.unsqueeze(0), cam)                                 # [K,3,N]
    z = pix[:, 2:3, :] + 1e-5
    uv = pix / z
    inside = (uv[:, 0] < width - edge) & (uv[:, 0] > edge) & (uv[:, 1] < height - edge) & (uv[:, 1] > edge)
    inside = inside & (z[:, 0] > 0)
    return inside.sum(dim=1) / pts.shape[0]


def keyframe_selection_overlap(gt_depth, w2c, intrinsics, keyframe_list, k, pixels=1600):
    """Ids (indices into keyframe_list) of up to k keyframes that overlap the current view, in random order.
    gt_depth [1,H,W]; keyframe_list: dicts with 'est_w2c' [4,4]."""
    width, height = gt_depth.shape[2], gt_depth.shape[1]
    valid = torch.stack(torch.where(gt_depth[0] > 0), dim=1)
    indices = torch.randint(valid.shape[0], (pixels,))                # global CPU generator, as the reference
    sampled = valid[indices.to(valid.device)]
    pts = get_pointcloud(gt_depth, intrinsics, w2c, sampled)
    if len(keyframe_list) == 0:
        return []
    est = torch.stack([kf["est_w2c"] for kf in keyframe_list]).to(pts.device)
    frac = overlap_fractions(pts, est, torch.as_tensor(intrinsics, device=pts.device, dtype=pts.dtype), width, height)
    frac = frac.tolist()                                              # one read-back for all keyframes
    order = sorted(range(len(frac)), key=lambda i: frac[i], reverse=True)     # stable, like sorted() of the dicts
    selected = [i for i in order if frac[i] > 0.0]
    return list(np.random.permutation(np.array(selected))[:k])
