"""ctypes wrapper of oracle/splat_oracle.c (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsplat_oracle.so")


class OracleCam(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("tanfovx", ctypes.c_float),
                ("tanfovy", ctypes.c_float), ("scale_modifier", ctypes.c_float), ("bg", ctypes.c_float * 3),
                ("view", ctypes.c_float * 16), ("proj", ctypes.c_float * 16)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.POINTER(OracleCam), ctypes.c_int] + [ctypes.c_void_p] * 5
        L.oracle_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_num_rendered.argtypes = [ctypes.c_void_p]
        L.oracle_get_geometry.argtypes = [ctypes.c_void_p] * 7
        L.oracle_get_binning.argtypes = [ctypes.c_void_p] * 4
        L.oracle_render.argtypes = [ctypes.c_void_p] * 5
        L.oracle_backward.argtypes = [ctypes.c_void_p] * 8
        L.oracle_backward.restype = ctypes.c_int
        L.oracle_mark_visible.argtypes = [ctypes.POINTER(OracleCam), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def make_cam(width, height, tanfovx, tanfovy, bg, viewmatrix, projmatrix, scale_modifier=1.0):
    """viewmatrix / projmatrix: the 16 floats of the reference's [1,4,4] tensors in flat order."""
    cam = OracleCam()
    cam.width, cam.height = int(width), int(height)
    cam.tanfovx, cam.tanfovy, cam.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    cam.bg[:] = [float(x) for x in np.asarray(bg, dtype=np.float32).reshape(3)]
    cam.view[:] = [float(x) for x in np.asarray(viewmatrix, dtype=np.float32).reshape(16)]
    cam.proj[:] = [float(x) for x in np.asarray(projmatrix, dtype=np.float32).reshape(16)]
    return cam


def _f32(a, shape):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))
    return a


class Oracle:
    """One scene/camera pair: preprocess + binning run at construction."""

    def __init__(self, cam, means3D, colors, opacities, scales, rotations):
        L = lib()
        self.cam = cam
        self.P = int(np.asarray(means3D).shape[0])
        P = self.P
        self._in = [_f32(means3D, (P, 3)), _f32(colors, (P, 3)), _f32(opacities, (P,)), _f32(scales, (P, 3)),
                    _f32(rotations, (P, 4))]
        self._h = L.oracle_create(ctypes.byref(cam), P, *[a.ctypes.data for a in self._in])
        self.R = L.oracle_num_rendered(self._h)
        self.W, self.H = cam.width, cam.height
        self.tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def geometry(self):
        P = self.P
        out = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
                   conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
                   cov3D=np.zeros((P, 6), np.float32))
        lib().oracle_get_geometry(self._h, out["radii"].ctypes.data, out["means2D"].ctypes.data,
                                  out["depths"].ctypes.data, out["conic_opacity"].ctypes.data,
                                  out["tiles_touched"].ctypes.data, out["cov3D"].ctypes.data)
        return out

    def binning(self):
        R = max(self.R, 1)
        keys, lst = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
        ranges = np.zeros((self.tiles, 2), np.uint32)
        lib().oracle_get_binning(self._h, keys.ctypes.data, lst.ctypes.data, ranges.ctypes.data)
        return dict(keys=keys[:self.R], point_list=lst[:self.R], ranges=ranges)

    def render(self):
        H, W = self.H, self.W
        out = dict(color=np.zeros((3, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
                   final_T=np.zeros((H, W), np.float32), n_contrib=np.zeros((H, W), np.uint32))
        lib().oracle_render(self._h, out["color"].ctypes.data, out["depth"].ctypes.data,
                            out["final_T"].ctypes.data, out["n_contrib"].ctypes.data)
        return out

    def backward(self, dL_dcolor):
        P = self.P
        g = _f32(dL_dcolor, (3, self.H, self.W))
        out = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
                   colors=np.zeros((P, 3), np.float32), opacities=np.zeros((P, 1), np.float32),
                   scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32))
        rc = lib().oracle_backward(self._h, g.ctypes.data, out["means3D"].ctypes.data, out["means2D"].ctypes.data,
                                   out["colors"].ctypes.data, out["opacities"].ctypes.data,
                                   out["scales"].ctypes.data, out["rotations"].ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_backward needs render() first")
        return out


def mark_visible(cam, means3D):
    m = _f32(means3D, (-1, 3))
    out = np.zeros(m.shape[0], np.uint8)
    lib().oracle_mark_visible(ctypes.byref(cam), m.shape[0], m.ctypes.data, out.ctypes.data)
    return out.astype(bool)
