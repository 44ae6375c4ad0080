"""Generates tests/golden/host/densify.npz by running the reference's OWN densify (R/utils/slam_external.py:191-243,
imported unmodified, CPU, this container) on a small map with a live torch Adam state: two calls (an accumulate-only
iteration and a densifying one).  `.cuda()` / device="cuda" are patched to stay on the CPU; torch's global RNG is
seeded right before each call so that the split samples can be reproduced.
Run:  python tests/golden/host/make_golden_densify.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "zeros_like", "ones", "eye"):
        orig = getattr(torch, name)
        setattr(torch, name, (lambda o: lambda *a, **k: o(*a, **{kk: vv for kk, vv in k.items()
                                                            if not (kk == "device" and str(vv).startswith("cuda"))}))(orig))
    torch.cuda.empty_cache = lambda: None
    spec = importlib.util.spec_from_file_location("ref_slam_external", os.path.join(REF, "utils", "slam_external.py"))
    se = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(se)
    P = 300
    g = torch.Generator().manual_seed(61)
    init = dict(means3D=torch.randn(P, 3, generator=g), rgb_colors=torch.rand(P, 3, generator=g),
                unnorm_rotations=torch.randn(P, 4, generator=g), logit_opacities=2.0 * torch.randn(P, 1, generator=g),
                log_scales=torch.log(0.005 + 0.08 * torch.rand(P, 1, generator=g)))
    grads = {k: torch.randn(v.shape, generator=g) for k, v in init.items()}
    m2d = [0.002 * torch.rand(P, 3, generator=g) for _ in range(2)]
    seen = [torch.rand(P, generator=g) < 0.8 for _ in range(2)]
    dd = dict(start_after=1, remove_big_after=0, stop_after=10, densify_every=2, grad_thresh=0.0006, num_to_split_into=2,
              removal_opacity_threshold=0.1, final_removal_opacity_threshold=0.1, reset_opacities=True,
              reset_opacities_every=2)
    scene_radius = 3.0
    params = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    params["cam_unnorm_rots"] = torch.nn.Parameter(torch.zeros(1, 4, 3))
    params["cam_trans"] = torch.nn.Parameter(torch.zeros(1, 3, 3))
    opt = torch.optim.Adam([{"params": [v], "name": k, "lr": 1e-3} for k, v in params.items()], lr=0.0, eps=1e-15)
    for k in init:
        params[k].grad = grads[k].clone()
    params["cam_unnorm_rots"].grad = torch.zeros(1, 4, 3); params["cam_trans"].grad = torch.zeros(1, 3, 3)
    opt.step()
    out = {"scene_radius": np.float32(scene_radius), "dict": np.array([repr(dd)])}
    for k in init:
        out["init_" + k], out["grad_" + k] = init[k].numpy(), grads[k].numpy()
    # NB no 'timestep' entry: the reference's densify does not grow variables['timestep'] when it clones, so with
    # SplaTAM's own `variables` dict its remove_points fails on the shape mismatch (slam_external.py:158-159) -- one more
    # sign that this branch is dead in SplaTAM (use_gaussian_splatting_densification=False everywhere)
    variables = dict(means2D_gradient_accum=torch.zeros(P), denom=torch.zeros(P), max_2D_radius=torch.zeros(P),
                     scene_radius=torch.tensor(scene_radius))
    for call, it in enumerate((1, 2)):
        m = torch.zeros(P, 3, requires_grad=True)
        m.grad = m2d[call].clone()
        variables["means2D"], variables["seen"] = m, seen[call]
        torch.manual_seed(70 + call)
        params, variables = se.densify(params, variables, opt, it, dd)
        out["m2d_%d" % call], out["seen_%d" % call], out["iter_%d" % call] = m2d[call].numpy(), seen[call].numpy(), np.int64(it)
        if call == 0:
            assert params["means3D"].shape[0] == P        # accumulate only
    for k in init:
        out["out_" + k] = params[k].detach().numpy()
        st = opt.state[params[k]]
        out["out_" + k + "_exp_avg"], out["out_" + k + "_exp_avg_sq"] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
    np.savez_compressed(os.path.join(HERE, "densify.npz"), **out)
    print("wrote densify.npz: %d -> %d Gaussians" % (P, params["means3D"].shape[0]))


if __name__ == "__main__":
    main()
