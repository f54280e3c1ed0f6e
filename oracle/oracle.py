"""ctypes front end of oracle/drt_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

Each wrapper names the reference code (/root/reference/...) its C function restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
EXT_IOR = 1.00029  # DiffRender.py:21


def build(force=False):
    src = os.path.join(_HERE, "drt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.orc_bvh_build.restype = vp
        L.orc_bvh_build.argtypes = [vp, i32, vp, i32]
        L.orc_bvh_free.argtypes = [vp]
        L.orc_bvh_num_nodes.restype = i32
        L.orc_bvh_num_nodes.argtypes = [vp]
        L.orc_closest_hit.argtypes = [vp, C.c_int, vp, i64, vp, vp, vp, vp]
        L.orc_trace_fwd.argtypes = [vp, C.c_int, vp, vp, vp, i64, f64, f64, vp, vp, vp, vp, vp, vp, vp]
        L.orc_chain_fwd.argtypes = [vp, vp, vp, vp, i64, f64, f64, vp, vp, vp, vp, vp]
        L.orc_trace_bwd.argtypes = [vp, vp, i32, vp, vp, i64, f64, f64, vp, vp, vp, vp, vp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class OracleMesh:
    """Query structure over the fp32 cast of the vertices, like optix_mesh (optix_extend.cpp:6-83)
    after Scene.update_mesh / update_verticex (DiffRender.py:311-313, 379-380)."""

    def __init__(self, vertices, faces):
        self.V64 = _c(vertices, np.float64)
        self.V32 = self.V64.astype(np.float32)
        self.F = _c(faces, np.int32)
        self.h = lib().orc_bvh_build(_p(self.V32), len(self.V32), _p(self.F), len(self.F))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_bvh_free(self.h)
            self.h = None

    @property
    def num_nodes(self):
        return lib().orc_bvh_num_nodes(self.h)

    def closest_hit(self, ray6, use_bvh=True, counters=False):
        """optix_mesh.intersect (optix_extend.cpp:29-57): f32[N,6] -> (T f32[N], ID i32[N])."""
        ray6 = _c(ray6, np.float32)
        n = len(ray6)
        T = np.empty(n, np.float32)
        ID = np.empty(n, np.int32)
        nn, nt = C.c_int64(0), C.c_int64(0)
        lib().orc_closest_hit(self.h, int(use_bvh), _p(ray6), n, _p(T), _p(ID), C.byref(nn), C.byref(nt))
        if counters:
            return T, ID, nn.value, nt.value
        return T, ID

    def trace_fwd(self, origin, ray_dir, int_ior, ext_ior=EXT_IOR, use_bvh=True, V64=None):
        """Scene.render_transparent (DiffRender.py:420-432) -> dict(out_ori, out_dir, mask, tri1, tri2,
        stage, counters[6] = nodes,tris of Q1,Q2,Q3 on the canonical LBVH)."""
        o = _c(origin, np.float64)
        d = _c(ray_dir, np.float64)
        V = self.V64 if V64 is None else _c(V64, np.float64)
        n = len(o)
        out = dict(out_ori=np.empty((n, 3)), out_dir=np.empty((n, 3)), mask=np.empty((n, 3), np.uint8),
                   tri1=np.empty(n, np.int32), tri2=np.empty(n, np.int32), stage=np.empty(n, np.uint8),
                   counters=np.zeros(6, np.int64))
        lib().orc_trace_fwd(self.h, int(use_bvh), _p(V), _p(o), _p(d), n, ext_ior, int_ior, _p(out["out_ori"]),
                            _p(out["out_dir"]), _p(out["mask"]), _p(out["tri1"]), _p(out["tri2"]), _p(out["stage"]),
                            _p(out["counters"]))
        out["mask"] = out["mask"].astype(bool)
        return out

    def chain_fwd(self, origin, ray_dir, tri1, tri2, int_ior, ext_ior=EXT_IOR, V64=None):
        """Differentiable part only, hit ids given: JIT_Dintersect + refract_ray twice
        (DiffRender.py:492-546)."""
        o = _c(origin, np.float64)
        d = _c(ray_dir, np.float64)
        V = self.V64 if V64 is None else _c(V64, np.float64)
        t1, t2 = _c(tri1, np.int32), _c(tri2, np.int32)
        n = len(o)
        oo, od, tf = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.uint8)
        lib().orc_chain_fwd(_p(V), _p(self.F), _p(o), _p(d), n, ext_ior, int_ior, _p(t1), _p(t2), _p(oo), _p(od), _p(tf))
        return oo, od, tf

    def trace_bwd(self, origin, ray_dir, tri1, tri2, g_ori, g_dir, int_ior, ext_ior=EXT_IOR, V64=None):
        """vertices.grad of sum(out_ori*g_ori + out_dir*g_dir) -- what loss.backward() (optim.py:210)
        produces through the autograd graph of DiffRender.py:492-546."""
        o = _c(origin, np.float64)
        d = _c(ray_dir, np.float64)
        V = self.V64 if V64 is None else _c(V64, np.float64)
        t1, t2 = _c(tri1, np.int32), _c(tri2, np.int32)
        go = None if g_ori is None else _c(g_ori, np.float64)
        gd = _c(g_dir, np.float64)
        gV = np.zeros_like(V)
        lib().orc_trace_bwd(_p(V), _p(self.F), len(V), _p(o), _p(d), len(o), ext_ior, int_ior, _p(t1), _p(t2), _p(go),
                            _p(gd), _p(gV))
        return gV


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
