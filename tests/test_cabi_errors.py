"""Error behaviour of the C ABI (include/drt_b200.h): status codes + drt_last_error instead of the
reference's C asserts (optix_extend.cpp:17-18,25,30-31).  Called through ctypes directly."""
import ctypes as C

import pytest
import torch

from drt_b200 import _lib


def _err():
    return _lib.load().drt_last_error().decode()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_without_gpu_fails_cleanly():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.drt_bvh_create(0, C.byref(h))
    assert rc != 0 and not h.value and len(_err()) > 0
    assert lib.drt_bvh_destroy(None) == 0


def test_null_and_range_arguments_cpu():
    lib = _lib.load()
    assert lib.drt_bvh_create(0, None) == 1 and "out is null" in _err()
    assert lib.drt_bvh_info(None, None) == 1
    assert lib.drt_closest_hit(None, None, 0, None, None, 1, 1, None) == 1
    assert lib.drt_trace_fwd(None, None, None, None, 0, 1.0, 1.5, None, None, None, None, None, None, None) == 1
    assert lib.drt_trace_bwd(None, None, None, None, 0, 1.0, 1.5, None, None, None, None, None, None) == 1
    assert lib.drt_ray_loss_grad(None, None, None, None, None, -1, None, None, None) == 1 and "N < 0" in _err()
    assert lib.drt_ray_loss_grad(None, None, None, None, None, 0, None, None, None) == 0
    assert lib.drt_ray_loss_step(None, None, None, 1, None, 0, 1.0, 1.5, 0, None, None, None, None, 0, 0, 0, None, None, None, None, None) == 1
    assert "null handle" in _err()
    assert lib.drt_generate_rays(4, 4, None, None, None, None, None) == 1 and "null buffer" in _err()
    assert lib.drt_kernel_launches() >= 0
    # the optional modes and the tuning switch (pure host-side argument checks)
    assert lib.drt_trace_fwd_smooth(None, None, None, None, None, 0, 1.0, 1.5, None, None, None, None, None, None) == 1
    assert lib.drt_trace_bwd_smooth(None, None, None, None, None, 0, 1.0, 1.5, None, None, None, None, None, None, None) == 1
    assert lib.drt_plane_hit(None, None, None, -1, None, None, None, None) == 1 and "N < 0" in _err()
    assert lib.drt_plane_hit(None, None, None, 0, None, None, None, None) == 0
    assert lib.drt_plane_hit_bwd(None, None, None, 3, None, None, None, None, None) == 1 and "null buffer" in _err()
    old = lib.drt_tuning_get(b"direct_max_rays")
    assert old >= 0
    assert lib.drt_tuning_set(b"direct_max_rays", 12345) == 0 and lib.drt_tuning_get(b"direct_max_rays") == 12345
    assert lib.drt_tuning_set(b"direct_max_rays", -1) == 1 and lib.drt_tuning_get(b"direct_max_rays") == 12345
    assert lib.drt_tuning_set(b"no_such_switch", 1) == 1 and "unknown key" in _err()
    assert lib.drt_tuning_set(None, 1) == 1 and lib.drt_tuning_get(b"no_such_switch") == -1
    assert lib.drt_tuning_set(b"direct_max_rays", old) == 0


@pytest.mark.gpu
def test_state_and_argument_errors_gpu(cuda_device):
    lib = _lib.load()
    dev = cuda_device
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    h = C.c_void_p()
    assert lib.drt_bvh_create(99, C.byref(h)) == 1 and "out of range" in _err()
    assert lib.drt_bvh_create(dev.index or 0, C.byref(h)) == 0
    ray = torch.zeros((4, 6), dtype=torch.float32, device=dev)
    hit = torch.zeros((4, 2), dtype=torch.float32, device=dev)
    # queries before any build: DRT_ERR_STATE (reference: assert(builded))
    assert lib.drt_closest_hit(h, p(ray), 4, p(hit), p(hit), 2, 2, st) == 3
    V = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=torch.float32, device=dev)
    assert lib.drt_bvh_update_vert(h, p(V), None, 3, 0, st) == 3
    F = torch.tensor([[0, 1, 2]], dtype=torch.int32, device=dev)
    assert lib.drt_bvh_build(h, p(F), -1, p(V), 3, st) == 1
    assert lib.drt_bvh_build(h, p(F), 1, None, 3, st) == 1
    assert lib.drt_bvh_build(h, p(F), 1, p(V), 0, st) == 1 and "without vertices" in _err()
    assert lib.drt_bvh_build(h, p(F), 1, p(V), 3, st) == 0
    info = (C.c_int64 * 8)()
    assert lib.drt_bvh_info(h, info) == 0 and list(info)[:4] == [1, 3, 1, 1]
    # update_vert: vertex count must match, exactly one of V32/V64
    assert lib.drt_bvh_update_vert(h, p(V), None, 4, 0, st) == 1 and "vertex count" in _err()
    assert lib.drt_bvh_update_vert(h, p(V), p(V), 3, 0, st) == 1
    assert lib.drt_bvh_update_vert(h, None, None, 3, 0, st) == 1
    V64 = V.double()
    assert lib.drt_bvh_update_vert(h, None, p(V64), 3, 1, st) == 0
    # closest hit argument checks
    assert lib.drt_closest_hit(h, p(ray), -1, p(hit), p(hit), 2, 2, st) == 1
    assert lib.drt_closest_hit(h, p(ray), 4, p(hit), p(hit), 0, 2, st) == 1 and "strides" in _err()
    assert lib.drt_closest_hit(h, C.c_void_p(ray.data_ptr() + 4), 3, p(hit), p(hit), 2, 2, st) == 1 and "aligned" in _err()
    assert lib.drt_closest_hit(h, None, 4, p(hit), p(hit), 2, 2, st) == 1
    assert lib.drt_closest_hit(h, None, 0, None, None, 2, 2, st) == 0
    # trace: rec / rec_count come as a pair, rec 16-byte aligned
    o = torch.zeros((4, 3), dtype=torch.float64, device=dev)
    m = torch.zeros((4, 3), dtype=torch.bool, device=dev)
    rec = torch.zeros((5, 4), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    args = lambda r, c: (h, p(V64), p(o), p(o), 4, 1.00029, 1.5, p(o.clone()), p(o.clone()), p(m), r, c, None, st)  # noqa: E731
    assert lib.drt_trace_fwd(*args(p(rec), None)) == 1 and "both" in _err()
    assert lib.drt_trace_fwd(*args(C.c_void_p(rec.data_ptr() + 4), p(cnt))) == 1 and "16-byte" in _err()
    assert lib.drt_trace_fwd(*args(p(rec), p(cnt))) == 0
    assert lib.drt_trace_fwd(*args(None, None)) == 0
    assert lib.drt_trace_fwd(h, None, p(o), p(o), 4, 1.0, 1.5, p(o), p(o), p(m), None, None, None, st) == 1
    assert lib.drt_trace_bwd(h, p(V64), p(o), p(o), 4, 1.0, 1.5, p(rec), p(cnt), None, None, p(V64), st) == 1
    torch.cuda.synchronize()
    assert cnt.item() == 0
    assert lib.drt_bvh_destroy(h) == 0
