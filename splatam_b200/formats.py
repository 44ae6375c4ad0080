"""Data formats either side of the hot path: SplaTAM's `params.npz` checkpoints
and the 3DGS-viewer PLY export, byte-compatible with what the reference writes.

* ``save_params`` / ``save_params_ckpt`` / ``load_params``   R/utils/common_utils.py:25-52 (np.savez of every entry of
  the params dict, tensors as contiguous CPU numpy arrays, non-tensors stored as they are)
* ``save_ply`` / ``load_ply`` / ``export_ply``               R/scripts/export_ply.py:20-77 (binary little-endian PLY,
  one `vertex` element of 17 float32 properties; the reference writes it through the `plyfile` package, which is
  not a dependency here: the header and the packed records are produced directly)
"""
import os

import numpy as np

C0 = 0.28209479177387814            # SH band-0 constant (export_ply.py:9)

PLY_ATTRS = ("x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity",
             "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3")


def params2cpu(params):
    out = {}
    for k, v in params.items():
        if hasattr(v, "detach"):                                  # torch.Tensor / Parameter
            out[k] = np.ascontiguousarray(v.detach().cpu().numpy())
        else:
            out[k] = v
    return out


def save_params(output_params, output_dir, name="params.npz"):
    """np.savez(<output_dir>/params.npz, **params) -- the file R/scripts/export_ply.py, the viewers and
    post_splatam_opt.py load."""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, name)
    np.savez(path, **params2cpu(output_params))
    return path


def save_params_ckpt(output_params, output_dir, time_idx):
    return save_params(output_params, output_dir, name="params" + str(time_idx) + ".npz")


def load_params(path, device=None):
    """dict of numpy arrays (device=None) or torch tensors on `device`; scalar / object entries (e.g. the stored
    intrinsics, `org_width`) are returned as numpy, like dict(np.load(path, allow_pickle=True))."""
    raw = dict(np.load(path, allow_pickle=True))
    if device is None:
        return raw
    import torch
    out = {}
    for k, v in raw.items():
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device).float()
        else:
            out[k] = v
    return out


def rgb_to_spherical_harmonic(rgb):
    return (rgb - 0.5) / C0


def spherical_harmonic_to_rgb(sh):
    return sh * C0 + 0.5


def _ply_header(n):
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    lines += [f"property float {a}" for a in PLY_ATTRS]
    lines += ["end_header", ""]
    return "\n".join(lines).encode("ascii")


def save_ply(path, means, scales, rotations, rgbs, opacities, normals=None):
    """Arguments as R/scripts/export_ply.py:20 -- the RAW stored parameters: `scales` are log-scales [N,1] or [N,3],
    `rotations` the unnormalised quaternions, `opacities` the logits [N,1]; colours are converted to SH band 0."""
    means = np.asarray(means, dtype=np.float32)
    n = means.shape[0]
    normals = np.zeros_like(means) if normals is None else np.asarray(normals, dtype=np.float32)
    colors = rgb_to_spherical_harmonic(np.asarray(rgbs, dtype=np.float32))
    scales = np.asarray(scales, dtype=np.float32)
    if scales.shape[1] == 1:
        scales = np.tile(scales, (1, 3))
    table = np.concatenate((means, normals, colors, np.asarray(opacities, dtype=np.float32).reshape(n, 1), scales,
                            np.asarray(rotations, dtype=np.float32)), axis=1).astype("<f4")
    assert table.shape == (n, len(PLY_ATTRS))
    with open(path, "wb") as f:
        f.write(_ply_header(n))
        f.write(np.ascontiguousarray(table).tobytes())
    return path


def load_ply(path):
    """Inverse of save_ply for files with exactly the 17 float properties (any order of header whitespace)."""
    with open(path, "rb") as f:
        blob = f.read()
    end = blob.index(b"end_header\n") + len(b"end_header\n")
    header = blob[:end].decode("ascii").split("\n")
    assert header[0] == "ply" and header[1].startswith("format binary_little_endian")
    n = int([h for h in header if h.startswith("element vertex")][0].split()[-1])
    props = [h.split()[-1] for h in header if h.startswith("property")]
    assert all(h.split()[1] == "float" for h in header if h.startswith("property"))
    table = np.frombuffer(blob, dtype="<f4", count=n * len(props), offset=end).reshape(n, len(props))
    col = {p: table[:, i] for i, p in enumerate(props)}
    pick = lambda names: np.stack([col[a] for a in names], axis=1)
    return dict(means3D=pick(("x", "y", "z")), normals=pick(("nx", "ny", "nz")),
                rgb_colors=spherical_harmonic_to_rgb(pick(("f_dc_0", "f_dc_1", "f_dc_2"))),
                logit_opacities=pick(("opacity",)), log_scales=pick(("scale_0", "scale_1", "scale_2")),
                unnorm_rotations=pick(("rot_0", "rot_1", "rot_2", "rot_3")))


def export_ply(params_path, ply_path=None):
    """params.npz -> splat.ply beside it (the __main__ of R/scripts/export_ply.py:55-77)."""
    p = dict(np.load(params_path, allow_pickle=True))
    ply_path = ply_path or os.path.join(os.path.dirname(params_path), "splat.ply")
    return save_ply(ply_path, p["means3D"], p["log_scales"], p["unnorm_rotations"], p["rgb_colors"], p["logit_opacities"])
