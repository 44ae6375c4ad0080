"""Synthetic scenes and cameras for BASELINE.json's configs.

Shared by tests/, bench.py, __graft_entry__.smoke() and tests/golden/make_golden.py.  Everything is
generated on the CPU with a seeded ``torch.Generator`` so the GPU box, this container and the golden
fixtures all see bit-identical inputs.  No dataset access (Replica / TUM are not available offline).
"""
import math

import numpy as np
import torch


def camera_matrices(w, h, fx, fy, cx, cy, w2c=None, near=0.01, far=100.0):
    """Float32 (viewmatrix, projmatrix, campos) exactly as R/utils/recon_helpers.py:4-27 builds them
    (w2c transposed; full projection = w2c^T @ opengl_proj^T), computed on the CPU."""
    w2c = torch.eye(4) if w2c is None else torch.as_tensor(w2c, dtype=torch.float32)
    w2c = w2c.float()
    cam_center = torch.inverse(w2c)[:3, 3]
    view = w2c.unsqueeze(0).transpose(1, 2)
    opengl_proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                                [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]]).float().unsqueeze(0).transpose(1, 2)
    full_proj = view.bmm(opengl_proj)
    return view.contiguous(), full_proj.contiguous(), cam_center.contiguous()


class Scene:
    """Inputs of one raster call in the operator's own parametrisation (post-exp scales,
    post-sigmoid opacities, normalised quaternions), as CPU float32 tensors."""

    def __init__(self, name, w, h, fx, fy, cx, cy, means3D, colors, opacities, scales, rotations,
                 w2c=None, bg=(0.0, 0.0, 0.0)):
        self.name, self.w, self.h = name, int(w), int(h)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.means3D, self.colors, self.opacities = means3D.float(), colors.float(), opacities.float().reshape(-1, 1)
        self.scales, self.rotations = scales.float(), rotations.float()
        self.view, self.proj, self.campos = camera_matrices(w, h, fx, fy, cx, cy, w2c)
        self.bg = torch.tensor(bg, dtype=torch.float32)
        self.tanfovx, self.tanfovy = w / (2 * fx), h / (2 * fy)

    @property
    def P(self):
        return self.means3D.shape[0]

    def settings(self, cls, device):
        """Build a GaussianRasterizationSettings-like NamedTuple of class `cls` on `device`."""
        return cls(image_height=self.h, image_width=self.w, tanfovx=self.tanfovx, tanfovy=self.tanfovy,
                   bg=self.bg.to(device), scale_modifier=1.0, viewmatrix=self.view.to(device),
                   projmatrix=self.proj.to(device), sh_degree=0, campos=self.campos.to(device), prefiltered=False)

    def inputs(self, device, requires_grad=False):
        d = dict(means3D=self.means3D, colors_precomp=self.colors, opacities=self.opacities, scales=self.scales,
                 rotations=self.rotations)
        out = {k: v.to(device).clone().requires_grad_(requires_grad) for k, v in d.items()}
        out["means2D"] = torch.zeros_like(out["means3D"], requires_grad=requires_grad)
        return out

    def oracle_cam(self):
        from oracle import oracle as O
        return O.make_cam(self.w, self.h, self.tanfovx, self.tanfovy, self.bg.numpy(), self.view.numpy().reshape(-1),
                          self.proj.numpy().reshape(-1))

    def oracle(self):
        from oracle import oracle as O
        return O.Oracle(self.oracle_cam(), self.means3D.numpy(), self.colors.numpy(), self.opacities.numpy(),
                        self.scales.numpy(), self.rotations.numpy())


def _rand(gen, *shape):
    return torch.rand(*shape, generator=gen)


def config1(seed=0, P=256, w=64, h=64, bg=(0.0, 0.0, 0.0)):
    """BASELINE config[0]: 256 mixed isotropic/anisotropic Gaussians, 64x64, fx=fy=64."""
    g = torch.Generator().manual_seed(seed)
    fx = fy = 64.0 * w / 64.0
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    z = 1.0 + 3.0 * _rand(g, P)
    u = -4.0 + (w + 8.0) * _rand(g, P)
    v = -4.0 + (h + 8.0) * _rand(g, P)
    means = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    s = 0.5 + 5.5 * _rand(g, P, 3)
    half = P // 2
    s[:half] = s[:half, :1]
    scales = s * z[:, None] / fx
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    q[:half] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    opac = torch.sigmoid(1.5 * torch.randn(P, generator=g))
    col = _rand(g, P, 3)
    return Scene("config1", w, h, fx, fy, cx, cy, means, col, opac, scales, q, bg=bg)


REPLICA = dict(w=1200, h=680, fx=600.0, fy=600.0, cx=599.5, cy=339.5)      # R/configs/data/replica.yaml
TUM_FR1 = dict(w=640, h=480, fx=517.3, fy=516.5, cx=318.6, cy=255.3)       # R/configs/data/TUM/freiburg1_desk.yaml


def config3(seed=2, P=1_000_000, cam=REPLICA, sigma_px=(0.7, 2.5), zr=(1.0, 5.0)):
    """BASELINE config[2]: P isotropic Gaussians, Replica intrinsics, sigma_px~U(0.7,2.5), z~U(1,5),
    opacity~U(.05,.95), colour~U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    w, h, fx, fy, cx, cy = cam["w"], cam["h"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    z = zr[0] + (zr[1] - zr[0]) * _rand(g, P)
    u, v = w * _rand(g, P), h * _rand(g, P)
    means = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    sig = sigma_px[0] + (sigma_px[1] - sigma_px[0]) * _rand(g, P)
    scales = (sig * z / fx)[:, None].repeat(1, 3)
    q = torch.zeros(P, 4)
    q[:, 0] = 1.0
    opac = 0.05 + 0.9 * _rand(g, P)
    col = _rand(g, P, 3)
    return Scene("config3_%d" % P, w, h, fx, fy, cx, cy, means, col, opac, scales, q)


def room(seed=1, P=50_000, cam=REPLICA, anisotropic=False, w2c=None):
    """BASELINE config[1]-like "room": points on the faces of a 6x4x3 m box around the camera plus
    clutter, sigma ~1.5-3 px, opacity 0.9; also used for TUM-sized scenes."""
    g = torch.Generator().manual_seed(seed)
    w, h, fx, fy, cx, cy = cam["w"], cam["h"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    n_wall = int(P * 0.85)
    face = torch.randint(0, 6, (n_wall,), generator=g)
    a, b = _rand(g, n_wall) * 2 - 1, _rand(g, n_wall) * 2 - 1
    half = torch.tensor([3.0, 2.0, 1.5])
    pts = torch.zeros(n_wall, 3)
    for f in range(6):
        m = face == f
        ax = f // 2
        sgn = 1.0 if f % 2 else -1.0
        o = [i for i in range(3) if i != ax]
        pts[m, ax] = sgn * half[ax]
        pts[m, o[0]] = a[m] * half[o[0]]
        pts[m, o[1]] = b[m] * half[o[1]]
    clutter = (torch.rand(P - n_wall, 3, generator=g) * 2 - 1) * half * 0.8
    pts = torch.cat([pts, clutter], 0)
    # camera at the origin looking down +z (x right, y down), as SplaTAM's first frame
    z = pts[:, 2].clamp(min=0.05)
    sig = 1.5 + 1.5 * _rand(g, P)
    base = sig * z.abs().clamp(min=0.3) / fx
    if anisotropic:
        scales = base[:, None] * (0.4 + 1.6 * _rand(g, P, 3))
        q = torch.randn(P, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
    else:
        scales = base[:, None].repeat(1, 3)
        q = torch.zeros(P, 4)
        q[:, 0] = 1.0
    opac = torch.full((P,), 0.9)
    col = 0.5 + 0.5 * torch.sin(pts * 3.0 + torch.tensor([0.0, 2.0, 4.0]))
    return Scene("room_%d" % P, w, h, fx, fy, cx, cy, pts, col, opac, scales, q, w2c=w2c)


def _surface_depth(u, v):
    """Smooth synthetic depth map (metres) standing in for a Replica / TUM depth frame."""
    return 2.5 + 0.8 * torch.sin(u / 130.0) * torch.cos(v / 90.0) + 0.3 * torch.sin(u / 23.0 + v / 31.0)


def view_filling(seed=12, cam=REPLICA, P=None, stride=1, anisotropic=False, w2c=None, opacity=(0.45, 0.6), margin=0,
                 cover=False):
    """A map that FILLS the view the way SplaTAM's maps do: one Gaussian per pixel of the first frame,
    back-projected through the depth map, sigma_world = depth / f (one pixel std on screen), logit opacity ~ 0
    (R/scripts/splatam.py:67-118,120-157,196-203: get_pointcloud + initialize_params with the "projective"
    mean_sq_dist), plus -- when P exceeds the pixel count -- further Gaussians at random sub-pixel positions of the
    same surface with sizes 0.8-1.3 px, standing in for the back-projections that later keyframes add
    (R/scripts/splatam.py:378-420).  Every Gaussian is in view, so num_rendered ~ 2 x P and every tile list is
    hundreds of entries long -- unlike `room`, where ~97 % of the Gaussians lie outside the frustum."""
    g = torch.Generator().manual_seed(seed)
    w, h, fx, fy, cx, cy = cam["w"], cam["h"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    # margin > 0 extends the surface beyond the image borders (pixels), so a moving camera keeps seeing a covered view
    vv, uu = torch.meshgrid(torch.arange(-margin, h + margin, stride, dtype=torch.float32),
                            torch.arange(-margin, w + margin, stride, dtype=torch.float32), indexing="ij")
    u, v = uu.reshape(-1), vv.reshape(-1)
    n_grid = u.numel()
    P = n_grid if P is None else int(P)
    if P < n_grid:
        keep = torch.randperm(n_grid, generator=g)[:P].sort().values
        u, v = u[keep], v[keep]
        # cover: a map with fewer Gaussians than pixels still fills the view (as a map built at a lower densification
        # resolution does): splat size grows with the pixel spacing
        size = torch.full((P,), math.sqrt(n_grid / P) if cover else 1.0)
    else:
        extra = P - n_grid
        u = torch.cat([u, -margin + (w + 2 * margin) * _rand(g, extra)])
        v = torch.cat([v, -margin + (h + 2 * margin) * _rand(g, extra)])
        size = torch.cat([torch.ones(n_grid), 0.8 + 0.5 * _rand(g, extra)])
    z = _surface_depth(u, v) * (1.0 + 0.004 * torch.randn(P, generator=g))
    means = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    base = size * z / ((fx + fy) / 2.0)
    if anisotropic:
        scales = base[:, None] * (0.5 + 1.0 * _rand(g, P, 3))
        q = torch.randn(P, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
    else:
        scales = base[:, None].repeat(1, 3)
        q = torch.zeros(P, 4)
        q[:, 0] = 1.0
    opac = opacity[0] + (opacity[1] - opacity[0]) * _rand(g, P)
    col = 0.5 + 0.5 * torch.sin(means * 6.0 + torch.tensor([0.0, 2.0, 4.0]))
    return Scene("view_filling_%d" % P, w, h, fx, fy, cx, cy, means, col, opac, scales, q, w2c=w2c)


def edge_cases(seed=5, w=97, h=45):
    """Ragged image size (not a multiple of 16), Gaussians behind / on the near plane, far off-screen,
    huge and sub-pixel splats, opacity below 1/255, exact depth ties, non-zero background, and a
    non-identity view matrix."""
    g = torch.Generator().manual_seed(seed)
    P = 300
    fx = fy = 80.0
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    z = 0.5 + 3.0 * _rand(g, P)
    u, v = -30 + (w + 60) * _rand(g, P), -30 + (h + 60) * _rand(g, P)
    means = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    means[:10, 2] = -1.0                      # behind the camera
    means[10:14, 2] = 0.2                     # exactly on the near plane (culled: z <= 0.2)
    means[14:18, 2] = 0.2000001               # just inside
    means[20:40, 2] = 2.0                     # exact depth ties -> order by index
    s = (0.3 + 4 * _rand(g, P, 3)) * z[:, None] / fx
    s[40:44] *= 40.0                          # huge splats covering the whole image
    s[44:50] *= 0.01                          # sub-pixel
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    opac = _rand(g, P)
    opac[50:56] = 0.003                       # below 1/255: never contributes
    opac[56:60] = 1.0
    col = _rand(g, P, 3)
    ang = 0.1
    w2c = torch.tensor([[math.cos(ang), 0, math.sin(ang), 0.05], [0, 1, 0, -0.02],
                        [-math.sin(ang), 0, math.cos(ang), 0.1], [0, 0, 0, 1]], dtype=torch.float32)
    return Scene("edge", w, h, fx, fy, cx, cy, means, col, opac, s, q, w2c=w2c, bg=(0.2, 0.5, 0.7))


def dense_opaque(seed=7, P=4000, w=128, h=96):
    """Many near-opaque overlapping splats so pixels saturate (T < 1e-4) and the early-out paths run."""
    g = torch.Generator().manual_seed(seed)
    fx = fy = 100.0
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    z = 1.0 + 2.0 * _rand(g, P)
    u, v = w * _rand(g, P), h * _rand(g, P)
    means = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    s = (3.0 + 5.0 * _rand(g, P, 3)) * z[:, None] / fx
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    opac = 0.9 + 0.1 * _rand(g, P)
    col = _rand(g, P, 3)
    return Scene("dense", w, h, fx, fy, cx, cy, means, col, opac, s, q)


def algorithmic_bytes(P, R, w, h):
    """B_algo of SURVEY.md section 8(d) / BASELINE.md section 4, bytes per fwd+bwd render."""
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    n, msb = tiles, 0
    # getHigherMsb (rasterizer_impl.cu:35-50)
    msb, step = 16, 16
    while step > 1:
        step //= 2
        msb = msb + step if (n >> msb) else msb - step
    if n >> msb:
        msb += 1
    passes = (32 + msb + 7) // 8
    return 180 * P + (72 + 24 * passes) * R + 44 * h * w
