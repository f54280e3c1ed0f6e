"""`optix` plugin module of the reference, re-implemented on libdrt_b200 (no OptiX).

The reference builds this module at import time from optix_extend.cpp (DiffRender.py:5-6) and uses
exactly one class, `optix.optix_mesh`, with four methods (optix_extend.cpp:77-83).  Same names,
argument meaning and dtypes here; errors are Python exceptions instead of C asserts.
"""
import ctypes as C

import torch

from . import _lib


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def check_on(device, **tensors):
    """Every tensor handed to a kernel by raw pointer must live on the handle's CUDA device: a CPU tensor or one on
    another GPU would be an illegal address inside the kernel (a sticky context error), not a Python exception."""
    device = torch.device(device)
    for name, t in tensors.items():
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
        if t.device != device:
            raise ValueError(f"{name} is on {t.device}, the mesh lives on {device}")


class optix_mesh:  # noqa: N801  (name fixed by the reference, optix_extend.cpp:6)
    """Ray-query object over one triangle mesh on one CUDA device."""

    def __init__(self, cuda_device=0):
        if not torch.cuda.is_available():
            raise _lib.DrtError("optix_mesh needs a CUDA device (drt_b200 has no CPU path)")
        self.device = torch.device("cuda", int(cuda_device))
        h = C.c_void_p()
        _lib.call("drt_bvh_create", int(cuda_device), C.byref(h))
        self._h = h
        self._faces = None   # int32 [F,3] kept for callers that want the face list back
        self.n_faces = 0
        self.n_verts = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().drt_bvh_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- optix_extend.cpp:14-21 -------------------------------------------------------------
    def update_mesh(self, F, V):
        """F int32 [F,3], V float32 (or float64) [V,3], CUDA, contiguous: set triangles + build BVH."""
        self._check(F, "F", (torch.int32,))
        self._check(V, "V", (torch.float32, torch.float64))
        F = F.contiguous()
        V = V.contiguous()
        fn = "drt_bvh_build" if V.dtype == torch.float32 else "drt_bvh_build_f64"
        _lib.call(fn, self._h, _ptr(F), F.shape[0], _ptr(V), V.shape[0], _stream_ptr(self.device))
        # the library copies F and V on the stream; keep them alive until then
        F.record_stream(torch.cuda.current_stream(self.device))
        V.record_stream(torch.cuda.current_stream(self.device))
        self._faces = F
        self.n_faces, self.n_verts = F.shape[0], V.shape[0]

    # -- optix_extend.cpp:23-27 -------------------------------------------------------------
    def update_vert(self, V, refit=False):
        """New vertex positions for the current faces; full rebuild like the reference unless refit."""
        self._check(V, "V", (torch.float32, torch.float64))
        V = V.contiguous()
        v32 = _ptr(V) if V.dtype == torch.float32 else C.c_void_p(0)
        v64 = _ptr(V) if V.dtype == torch.float64 else C.c_void_p(0)
        _lib.call("drt_bvh_update_vert", self._h, v32, v64, V.shape[0], int(bool(refit)), _stream_ptr(self.device))
        V.record_stream(torch.cuda.current_stream(self.device))

    # -- optix_extend.cpp:29-57 -------------------------------------------------------------
    def intersect(self, Ray):
        """Ray float32 [N,6] (origin, direction) -> [T float32 [N], ID int32 [N]], both strided views of
        one [N,2] buffer exactly like the reference's {float t; int id} hit records; miss: T=-1, ID=-1."""
        self._check(Ray, "Ray", (torch.float32,), cols=6)
        Ray = Ray.contiguous()
        n = Ray.shape[0]
        hit = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        hit_i = hit.view(torch.int32)
        _lib.call("drt_closest_hit", self._h, _ptr(Ray), n, _ptr(hit), C.c_void_p(hit_i.data_ptr() + 4), 2, 2,
                  _stream_ptr(self.device))
        Ray.record_stream(torch.cuda.current_stream(self.device))
        return [hit[:, 0], hit_i[:, 1]]

    def set_image_size(self, resy, resx):
        """Render resolution hint (the reference's module globals DiffRender.resy / resx, DiffRender.py:16-17): when a
        batch of rays is whole resy x resx images the entry query works on 32-pixel tiles (4x8, else 8x4).  Same results either way."""
        if (resy, resx) != getattr(self, "_image_size", None):
            _lib.call("drt_bvh_set_image_size", self._h, int(resx), int(resy))
            self._image_size = (resy, resx)

    # -- helpers ----------------------------------------------------------------------------
    def info(self):
        a = (C.c_int64 * 8)()
        _lib.call("drt_bvh_info", self._h, a)
        keys = ("n_faces", "n_verts", "n_nodes", "built", "node_bytes", "tri_bytes", "builds", "refits")
        return dict(zip(keys, list(a)))

    def last_counts(self):
        """stage counters of the latest fused ray-loss step (synchronises): entry hits, alive after both refractions, valid
        paths, tiles seen / kept by the beam pass, lanes the call ran on (0: one-thread-per-path route)"""
        a = (C.c_int64 * 6)()
        _lib.call("drt_bvh_last_counts", self._h, _stream_ptr(self.device), a)
        return dict(zip(("entry_hits", "alive", "valid_paths", "tiles", "tiles_kept", "lanes"), list(a)))

    def bad_indices(self):
        out = C.c_int(0)
        _lib.call("drt_bvh_bad_indices", self._h, _stream_ptr(self.device), C.byref(out))
        return out.value

    def _check(self, t, name, dtypes, cols=3):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dim() != 2 or t.shape[1] != cols:
            raise ValueError(f"{name} must have shape [n,{cols}], got {tuple(t.shape)}")  # assert(size(1)==..)
        if t.dtype not in dtypes:
            raise TypeError(f"{name} must be {' or '.join(str(d) for d in dtypes)}, got {t.dtype}")
        if t.device != self.device:
            raise ValueError(f"{name} is on {t.device}, the mesh lives on {self.device}")
