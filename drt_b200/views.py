"""Synthetic turntable views (the captured .h5 data of the reference is not distributed:
captured_data.py:94-108, README.md:18).  Rays follow captured_data.generate_ray's convention
(captured_data.py:23-40): integer pixel coordinates (x, y, 1), K^-1 then the camera-to-world
transform, direction normalised, origin = camera centre broadcast to every pixel.
"""
import math

import numpy as np
import torch


def turntable_cameras(vertices, resy, resx, n_views=72, fill=0.87, focal_scale=1.4, step_deg=None):
    """-> list of (R 4x4 world->camera, K 3x3, R_inverse 4x4, K_inverse 3x3), float64 numpy
    (the camera_M tuple of captured_data.py:112-118).  Turntable about +y through the bbox centre,
    view k at azimuth step_deg*k (default 360/n_views; 72 views = 5 degrees)."""
    v = np.asarray(vertices, dtype=np.float64)
    lo, hi = v.min(0), v.max(0)
    ctr = 0.5 * (lo + hi)
    half_y = 0.5 * (hi[1] - lo[1])
    half_h = 0.5 * math.hypot(hi[0] - lo[0], hi[2] - lo[2])
    f = focal_scale * max(resx, resy)
    dist = f * max(half_y / (0.5 * fill * resy), half_h / (0.5 * fill * resx))
    K = np.array([[f, 0, resx / 2.0], [0, f, resy / 2.0], [0, 0, 1.0]])
    K_inv = np.linalg.inv(K)
    step = 360.0 / n_views if step_deg is None else step_deg
    cams = []
    for k in range(n_views):
        a = math.radians(step * k)
        eye = ctr + dist * np.array([math.sin(a), 0.0, math.cos(a)])
        fwd = (ctr - eye) / np.linalg.norm(ctr - eye)
        right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        Rm = np.stack([right, -up, fwd], axis=0)  # x right, y down, z forward
        R = np.eye(4)
        R[:3, :3] = Rm
        R[:3, 3] = -Rm @ eye
        R_inv = np.eye(4)
        R_inv[:3, :3] = Rm.T
        R_inv[:3, 3] = eye
        cams.append((R, K, R_inv, K_inv))
    return cams


def generate_ray(resy, resx, K_inverse, R_inverse, device="cpu", dtype=torch.float64):
    """Per-pixel rays for one view, row-major (y outer, x inner) like captured_data.py:23-40.
    -> (origin [resy*resx,3], ray_dir [resy*resx,3])."""
    Ki = torch.as_tensor(np.asarray(K_inverse), dtype=dtype, device=device)
    Ri = torch.as_tensor(np.asarray(R_inverse), dtype=dtype, device=device)
    ys = torch.arange(resy, dtype=dtype, device=device)
    xs = torch.arange(resx, dtype=dtype, device=device)
    py, px = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([px, py, torch.ones_like(px)], dim=-1).reshape(-1, 3)
    cam = pix @ Ki.T
    world = cam @ Ri[:3, :3].T + Ri[:3, 3]
    origin = Ri[:3, 3].expand_as(world)
    d = world - origin
    d = d / d.norm(dim=1, keepdim=True)
    return origin.contiguous(), d.contiguous()


def generate_ray_device(resy, resx, K_inverse, R_inverse, device, out_dir=None):
    """captured_data.generate_ray (captured_data.py:23-40) as ONE kernel (drt_generate_rays): -> (origin [1,3],
    ray_dir [resy*resx,3]) on `device`; `origin.expand_as(ray_dir)` is what the reference returns.  Agrees with
    `generate_ray` to a few ulp (the reference's two matmuls leave the summation order to the BLAS)."""
    import ctypes as C

    from . import _lib
    dev = torch.device(device)
    Ki = torch.as_tensor(np.asarray(K_inverse), dtype=torch.float64).to(dev).contiguous()
    Ri = torch.as_tensor(np.asarray(R_inverse), dtype=torch.float64).to(dev).contiguous()
    if Ki.shape != (3, 3) or Ri.shape != (4, 4):
        raise ValueError("K_inverse must be [3,3] and R_inverse [4,4]")
    origin = torch.empty((1, 3), dtype=torch.float64, device=dev)
    d = out_dir if out_dir is not None else torch.empty((resy * resx, 3), dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        _lib.call("drt_generate_rays", int(resy), int(resx), p(Ki), p(Ri), p(origin), p(d),
                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return origin, d


def view_batch(cams, resy, resx, device="cpu", out_origin=None, out_dir=None):
    """Rays of several views concatenated view-major: [len(cams)*resy*resx, 3] each."""
    n = resy * resx
    if out_origin is None:
        out_origin = torch.empty((len(cams) * n, 3), dtype=torch.float64, device=device)
        out_dir = torch.empty_like(out_origin)
    for k, (_, _, R_inv, K_inv) in enumerate(cams):
        o, d = generate_ray(resy, resx, K_inv, R_inv, device=device)
        out_origin[k * n:(k + 1) * n] = o
        out_dir[k * n:(k + 1) * n] = d
    return out_origin, out_dir


def stratified_sample(cfg, target=65536, device="cpu"):
    """Every `stride`-th ray of the config's whole ray set (all views, global scanline order), stride = total // target:
    the sample the roofline denominator is frozen on (profiles/canonical_counters.json) and the headline
    configs are parity-checked on.  -> (origin [n,3], ray_dir [n,3], stride) float64 numpy.  `device`: where the
    per-view rays are generated before the sample is picked (a CUDA matmul may round the last ulp differently)."""
    n_pix = cfg["resy"] * cfg["resx"]
    total = n_pix * cfg["n_views"]
    stride = max(1, total // target)
    os_, ds_ = [], []
    for k, (_, _, R_inv, K_inv) in enumerate(cfg["cams"]):
        first = (-(k * n_pix)) % stride  # global index k*n_pix + j must be a multiple of stride
        idx = np.arange(first, n_pix, stride)
        if len(idx) == 0:
            continue
        o, d = generate_ray(cfg["resy"], cfg["resx"], K_inv, R_inv, device=device)
        pick = torch.as_tensor(idx, device=device)
        os_.append(o[pick].cpu().numpy())
        ds_.append(d[pick].cpu().numpy())
    return np.concatenate(os_), np.concatenate(ds_), stride
