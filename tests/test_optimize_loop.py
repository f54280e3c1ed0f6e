"""Config C3 shape: the optim.py loop (examples/optimize_synthetic.py) runs on the drop-in Scene and
reduces the ray loss (reference optim.py:145-219)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_optimisation_loop_reduces_ray_loss(cuda_device):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import optimize_synthetic as ex
    from drt_b200 import configs, synthetic_data
    hp = {"IOR": 1.4723, "ray_w": 40, "sm_w": 0.08, "vh_w": 2e-3, "momentum": 0.95, "start_lr": 0.1}
    v, f = configs.load_mesh("hand_vh")
    target = configs.perturbed_target_mesh(v, scale=0.6)
    data = synthetic_data.SyntheticData(target, f, 120, 160, n_views=12, num_view=12, int_ior=hp["IOR"])
    assert len(data.Views) == 12 and data.Views[0][3].is_pinned()
    losses = {}
    for fused in (True, False):
        data.rng = np.random.default_rng(0)
        scene, hist = ex.optimize(v, f, data, hp, iters=24, log_every=0, fused_loss=fused)
        h = np.array(hist)
        assert np.isfinite(h).all()
        first, last = h[:6, 0].mean(), h[-6:, 0].mean()
        assert last < 0.8 * first, (first, last)
        losses[fused] = h[:, 0]
    # the fused RayLoss and the reference-style expression drive the same trajectory
    assert np.allclose(losses[True], losses[False], rtol=1e-6)
