"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by the product path).

Plain-torch restatement of the reference's map-growing logic, device-agnostic, used as the checker for
csrc/map_ops.cu.  It is pinned against tests/golden/host/host_ops.npz, which was produced by the reference's own
functions (tests/golden/host/make_golden_host.py).

  non_presence_mask      R/scripts/splatam.py:385-405   (inlined in add_new_gaussians)
  backproject            R/scripts/splatam.py:67-118    (get_pointcloud) + :348-375 (initialize_new_params)
  prune_keep_mask        R/utils/slam_external.py:177-186
"""
import torch


def depth_error(depth_sil, gt_depth):
    gt, rd = gt_depth.reshape(gt_depth.shape[-2:]), depth_sil[0]
    return (gt - rd).abs() * (gt > 0)


def non_presence_mask(depth_sil, gt_depth, sil_thres):
    gt, rd, sil = gt_depth.reshape(gt_depth.shape[-2:]), depth_sil[0], depth_sil[1]
    err = depth_error(depth_sil, gt_depth)
    grow = (sil < sil_thres) | ((rd > gt) & (err > 50 * err.median()))
    return (grow & (gt > 0)).reshape(-1)


def backproject(color, depth, K, w2c, mask=None, scale_dim=1):
    H, W = color.shape[-2:]
    dev = color.device
    u = torch.arange(W, device=dev).float().repeat(H)
    v = torch.arange(H, device=dev).float().repeat_interleave(W)
    z = depth.reshape(-1)
    cam = torch.stack(((u - K[0][2]) / K[0][0] * z, (v - K[1][2]) / K[1][1] * z, z, torch.ones_like(z)), dim=1)
    world = (torch.inverse(w2c) @ cam.T).T[:, :3]
    rgb = color.permute(1, 2, 0).reshape(-1, 3)
    msd = (z / ((K[0][0] + K[1][1]) / 2)) ** 2
    if mask is not None:
        world, rgb, msd = world[mask], rgb[mask], msd[mask]
    n = world.shape[0]
    rots = torch.zeros(n, 4, device=dev); rots[:, 0] = 1
    return dict(means3D=world, rgb_colors=rgb, unnorm_rotations=rots, logit_opacities=torch.zeros(n, 1, device=dev),
                log_scales=torch.log(torch.sqrt(msd))[:, None].repeat(1, scale_dim)), msd


def prune_keep_mask(logit_opacities, log_scales, opacity_threshold, big_threshold=None):
    remove = (torch.sigmoid(logit_opacities) < opacity_threshold).reshape(-1)
    if big_threshold is not None:
        remove = remove | (torch.exp(log_scales).max(dim=1).values > big_threshold)
    return ~remove
