"""Index arithmetic of the kernels restated in numpy and checked exhaustively on the CPU:
  * TileMap::ray_of (drt_b200/csrc/trace.cuh) -- work item -> ray for 4x8 / 8x4 / 16x2 pixel tiles is a bijection per image,
    every 32 consecutive work items are one tile, images stay separate;
  * tgt_bucket_kernel + TargetSrc::get (drt_b200/csrc/loss_step.cuh) -- the bucketed lookup of the sparse screen targets
    finds exactly the rays that have a target, for any N (also not a multiple of the bucket size) and any target set."""
import numpy as np
import pytest

K_TGT_SHIFT = 6


def ray_of(item, img_w, img_hw, tw_log2):
    v, r = np.divmod(item, img_hw)
    t, w = r >> 5, r & 31
    tpr = img_w >> tw_log2
    ty, tx = np.divmod(t, tpr)
    return v * img_hw + (ty * (32 >> tw_log2) + (w >> tw_log2)) * img_w + (tx << tw_log2) + (w & ((1 << tw_log2) - 1))


@pytest.mark.parametrize("tw_log2", [2, 3, 4])
@pytest.mark.parametrize("res", [(720, 960), (96, 128), (16, 32), (1080, 1920)])
def test_tile_map_is_a_bijection_of_whole_tiles(tw_log2, res):
    h, w = res
    tw, th = 1 << tw_log2, 32 >> tw_log2
    assert w % tw == 0 and h % th == 0
    n_views = 3
    items = np.arange(n_views * h * w, dtype=np.int64)
    rays = ray_of(items, w, h * w, tw_log2)
    assert np.array_equal(np.sort(rays), items)                               # every ray exactly once
    assert np.array_equal(rays // (h * w), items // (h * w))                  # images do not mix
    y, x = np.divmod(rays % (h * w), w)
    yb, xb = y.reshape(-1, 32), x.reshape(-1, 32)                             # one warp batch per row
    assert ((yb.max(1) - yb.min(1)) == th - 1).all() and ((xb.max(1) - xb.min(1)) == tw - 1).all()
    assert (yb.min(1) % th == 0).all() and (xb.min(1) % tw == 0).all()
    assert np.array_equal(xb[:, :tw], xb[:, :1] + np.arange(tw))              # x fastest inside a tile: contiguous runs of tw rays


def bucket_table(idx, n_rays):
    nb = (n_rays >> K_TGT_SHIFT) + 2
    keys = np.arange(nb, dtype=np.int64) << K_TGT_SHIFT
    return np.searchsorted(idx, keys, side="left").astype(np.int64)          # what tgt_bucket_kernel's binary search computes


def lookup(i, idx, bucket):
    lo, hi = bucket[i >> K_TGT_SHIFT], bucket[(i >> K_TGT_SHIFT) + 1]
    while lo < hi:                                                           # TargetSrc::get
        mid = (lo + hi) >> 1
        if idx[mid] < i:
            lo = mid + 1
        else:
            hi = mid
    return lo if lo < len(idx) and idx[lo] == i else -1


@pytest.mark.parametrize("n_rays", [1, 63, 64, 65, 1000, 4096, 70001])
def test_bucketed_target_lookup_finds_exactly_the_targets(n_rays):
    rng = np.random.default_rng(n_rays)
    for density in (0.0, 0.02, 0.5, 1.0):
        has = rng.random(n_rays) < density
        if density == 0.02 and n_rays > 2:
            has[[0, n_rays - 1]] = True                                       # first and last ray
        idx = np.nonzero(has)[0].astype(np.int64)
        bucket = bucket_table(idx, n_rays)
        assert len(bucket) == (n_rays >> K_TGT_SHIFT) + 2 and bucket[0] == 0 and bucket[-1] == len(idx)
        assert (np.diff(bucket) >= 0).all() and (np.diff(bucket) <= 1 << K_TGT_SHIFT).all()
        probe = np.arange(n_rays) if n_rays <= 4096 else rng.integers(0, n_rays, 5000)
        for i in probe:
            k = lookup(int(i), idx, bucket)
            assert (k >= 0) == bool(has[i]) and (k < 0 or idx[k] == i)
