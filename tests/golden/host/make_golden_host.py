"""Generates tests/golden/host/host_ops.npz and tests/golden/host/ref_params.npz by running the reference's OWN Python
functions (imported / compiled from /root/reference, CPU, this container) for the rows either side of the hot path:

* keyframe_selection_overlap                    R/utils/keyframe_selection.py   (imported as a module)
* save_params                                   R/utils/common_utils.py         (imported as a module)
* prune_gaussians / remove_points               R/utils/slam_external.py        (imported as a module)
* get_pointcloud / initialize_new_params        R/scripts/splatam.py            (the two function definitions are
  compiled from the file's AST: importing the whole script needs datasets / wandb / cv2, which are not installed)

The reference hard-codes `.cuda()`; for this CPU run `torch.Tensor.cuda` is patched to the identity.  Nothing is
copied into the repo: the functions run where they lie.  Run:  python tests/golden/host/make_golden_host.py
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def functions_from_source(path, names, namespace):
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), namespace)
    return [namespace[n] for n in names]


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self                 # the reference hard-codes .cuda()
    _zeros = torch.zeros
    torch.zeros = lambda *a, **k: _zeros(*a, **{kk: vv for kk, vv in k.items() if not (kk == "device" and vv == "cuda")})
    out = {}

    # ---- keyframe selection ---------------------------------------------------------------------------------
    ks = load_module(os.path.join(REF, "utils", "keyframe_selection.py"), "ref_keyframe_selection")
    g = torch.Generator().manual_seed(7)
    H, W = 96, 128
    depth = 1.0 + 2.0 * torch.rand(1, H, W, generator=g)
    depth[0, :10, :17] = 0.0                                        # invalid-depth corner
    K = torch.tensor([[100.0, 0, 63.5], [0, 100.0, 47.5], [0, 0, 1]])
    def pose(rx, ry, t):
        cx, sx, cy, sy = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry)
        R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = t
        return torch.tensor(M, dtype=torch.float32)
    w2c = pose(0.02, -0.03, [0.05, -0.02, 0.1])
    kfs = [pose(0.0, 0.0, [0, 0, 0]), pose(0.05, 0.3, [0.4, 0, 0.2]), pose(0.0, 3.1, [0, 0, 0]),
           pose(-0.1, -0.6, [-0.8, 0.1, 0.0]), pose(0.0, 1.2, [0.0, 0.0, 1.5]), pose(0.3, 0.05, [0.1, 0.5, -0.2])]
    keyframe_list = [{"est_w2c": m} for m in kfs]
    torch.manual_seed(11); np.random.seed(12)
    sel = ks.keyframe_selection_overlap(depth, w2c, K, keyframe_list, k=3, pixels=400)
    torch.manual_seed(11)
    valid = torch.stack(torch.where(depth[0] > 0), dim=1)
    sampled = valid[torch.randint(valid.shape[0], (400,))]
    pts = ks.get_pointcloud(depth, K, w2c, sampled)
    out.update(kf_depth=depth.numpy(), kf_K=K.numpy(), kf_w2c=w2c.numpy(), kf_list=torch.stack(kfs).numpy(),
               kf_selected=np.array(sel, dtype=np.int64), kf_pts=pts.numpy())

    # ---- prune_gaussians (with a live torch Adam state) --------------------------------------------------------
    se = load_module(os.path.join(REF, "utils", "slam_external.py"), "ref_slam_external")
    P = 400
    g = torch.Generator().manual_seed(21)
    init = dict(means3D=torch.randn(P, 3, generator=g), rgb_colors=torch.rand(P, 3, generator=g),
                unnorm_rotations=torch.randn(P, 4, generator=g), logit_opacities=3.0 * torch.randn(P, 1, generator=g),
                log_scales=torch.log(0.02 + 0.3 * torch.rand(P, 1, generator=g)))
    grads = {k: torch.randn(v.shape, generator=g) for k, v in init.items()}
    for tag, it, prune_dict in [
            ("a", 20, dict(start_after=0, remove_big_after=0, stop_after=20, prune_every=20, removal_opacity_threshold=0.005,
                           final_removal_opacity_threshold=0.05, reset_opacities=False, reset_opacities_every=500)),
            ("b", 40, dict(start_after=0, remove_big_after=100, stop_after=200, prune_every=20, removal_opacity_threshold=0.3,
                           final_removal_opacity_threshold=0.05, reset_opacities=True, reset_opacities_every=40))]:
        params = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
        params["cam_unnorm_rots"] = torch.nn.Parameter(torch.zeros(1, 4, 3))
        params["cam_trans"] = torch.nn.Parameter(torch.zeros(1, 3, 3))
        opt = torch.optim.Adam([{"params": [v], "name": k, "lr": 1e-3} for k, v in params.items()], lr=0.0, eps=1e-15)
        for k in init:
            params[k].grad = grads[k].clone()
        params["cam_unnorm_rots"].grad = torch.zeros(1, 4, 3); params["cam_trans"].grad = torch.zeros(1, 3, 3)
        opt.step()
        stepped = {k: params[k].detach().clone() for k in init}
        variables = dict(means2D_gradient_accum=torch.zeros(P), denom=torch.zeros(P), max_2D_radius=torch.zeros(P),
                         timestep=torch.zeros(P), scene_radius=torch.tensor(2.0))
        params, variables = se.prune_gaussians(params, variables, opt, it, prune_dict)
        for k in init:
            out[f"prune_{tag}_{k}"] = params[k].detach().numpy()
            st = opt.state[params[k]]
            out[f"prune_{tag}_{k}_exp_avg"] = st["exp_avg"].numpy()
            out[f"prune_{tag}_{k}_exp_avg_sq"] = st["exp_avg_sq"].numpy()
            out[f"prune_{tag}_{k}_before"] = stepped[k].numpy()
        out[f"prune_{tag}_iter"] = np.int64(it)
        out[f"prune_{tag}_dict"] = np.array([repr(prune_dict)])
    for k in init:
        out["prune_init_" + k] = init[k].numpy()
        out["prune_grad_" + k] = grads[k].numpy()
    out["prune_scene_radius"] = np.float32(2.0)

    # ---- get_pointcloud + initialize_new_params -------------------------------------------------------------
    ns = {"torch": torch, "np": np}
    get_pointcloud, initialize_new_params = functions_from_source(
        os.path.join(REF, "scripts", "splatam.py"), ["get_pointcloud", "initialize_new_params"], ns)
    g = torch.Generator().manual_seed(31)
    H, W = 40, 56
    color = torch.rand(3, H, W, generator=g)
    depth = 0.5 + 3.0 * torch.rand(1, H, W, generator=g)
    depth[0, 5:9, 20:30] = 0.0
    K = torch.tensor([[60.0, 0, 27.5], [0, 62.0, 19.5], [0, 0, 1]])
    w2c = pose(0.1, -0.2, [0.3, -0.1, 0.25])
    mask = (torch.rand(H * W, generator=g) < 0.3) & (depth[0].reshape(-1) > 0)
    for tag, dist in [("iso", "isotropic"), ("aniso", "anisotropic")]:
        pt_cld, msd = get_pointcloud(color, depth, K, w2c, mask=mask, compute_mean_sq_dist=True,
                                     mean_sq_dist_method="projective")
        new = initialize_new_params(pt_cld, msd, dist)
        for k, v in new.items():
            out[f"bp_{tag}_{k}"] = v.detach().numpy()
        out[f"bp_{tag}_mean_sq_dist"] = msd.numpy()
    full, msd_full = get_pointcloud(color, depth, K, w2c, mask=None, compute_mean_sq_dist=True,
                                    mean_sq_dist_method="projective")
    out.update(bp_color=color.numpy(), bp_depth=depth.numpy(), bp_K=K.numpy(), bp_w2c=w2c.numpy(), bp_mask=mask.numpy(),
               bp_full_pts=full.numpy(), bp_full_mean_sq_dist=msd_full.numpy())
    np.savez_compressed(os.path.join(HERE, "host_ops.npz"), **out)

    # ---- save_params -----------------------------------------------------------------------------------------
    cu = load_module(os.path.join(REF, "utils", "common_utils.py"), "ref_common_utils")
    import shutil, tempfile
    g = torch.Generator().manual_seed(41)
    params = dict(means3D=torch.randn(7, 3, generator=g), rgb_colors=torch.rand(7, 3, generator=g),
                  unnorm_rotations=torch.randn(7, 4, generator=g), logit_opacities=torch.randn(7, 1, generator=g),
                  log_scales=torch.randn(7, 1, generator=g), cam_unnorm_rots=torch.randn(1, 4, 5, generator=g),
                  cam_trans=torch.randn(1, 3, 5, generator=g), timestep=torch.arange(7).float(),
                  intrinsics=np.eye(3, dtype=np.float32), org_width=1200, org_height=680)
    d = tempfile.mkdtemp()
    cu.save_params(params, d)
    shutil.copy(os.path.join(d, "params.npz"), os.path.join(HERE, "ref_params.npz"))
    print("wrote host_ops.npz (%d arrays) and ref_params.npz; keyframes selected: %s" % (len(out), sel))


if __name__ == "__main__":
    main()
