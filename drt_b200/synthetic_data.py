"""Synthetic stand-in for captured_data.Data_Pointgray / Data_Redmi (captured_data.py:85-165): the
captured .h5 sets are not distributed (README.md:18), so the 72 views are rendered from a TARGET mesh
with this package's own tracer.  Same interface the reference's Loss_calculator uses (optim.py:59-108):
`Views[i] = (screen_pixel, valid, mask, origin, ray_dir, camera_M)` in pinned host memory, `get_view(i)`
uploading one view, `ray_view_generator()` / `silh_view_generator()` as shuffled infinite generators."""
import numpy as np
import torch

from . import views as _views
from . import DiffRender as R
from .captured_data import CompactViews


class SyntheticData(CompactViews):
    def __init__(self, target_vertices, faces, resy, resx, n_views=72, num_view=72, cuda_device=0, screen_dist=100.0,
                 int_ior=None, seed=0):
        self.resy, self.resx, self.num_view, self.n_views = resy, resx, num_view, n_views
        self.device = torch.device("cuda", cuda_device)
        self.rng = np.random.default_rng(seed)
        self.keep_on_device, self._resident = False, {}
        if int_ior is not None:
            R.intIOR = int_ior
        scene = R.Scene(vertices=target_vertices, faces=faces, cuda_device=cuda_device)
        cams = _views.turntable_cameras(np.asarray(target_vertices), resy, resx, n_views)
        pin = lambda t: t.cpu().pin_memory()  # noqa: E731
        self.Views = []
        with torch.no_grad():
            for cam in cams:
                Rm, K, R_inv, K_inv = (torch.tensor(m, dtype=torch.float64) for m in cam)
                origin, ray_dir = _views.generate_ray(resy, resx, cam[3], cam[2], device=self.device)
                out_ori, out_dir, mask3 = scene.render_transparent(origin, ray_dir)
                screen = (out_ori + screen_dist * out_dir) * mask3[:, :1]   # zeros where nothing was measured
                valid = screen[:, 0] != 0                                    # captured_data.py:104
                sil = scene.render_mask(origin, ray_dir)                     # [0,1] silhouette image, flat
                self.Views.append((pin(screen), pin(valid), pin(sil), pin(origin), pin(ray_dir),
                                   (pin(Rm), pin(K), pin(R_inv), pin(K_inv))))

    def get_view(self, V_index):  # captured_data.py:44-59
        """Uploads one view.  With `self.keep_on_device = True` every view is uploaded once and then served from
        HBM: 72 views of 960x1280 float64 rays + targets are 8.6 GB, nothing on a 180 GB part (the reference
        re-uploads ~100 MB per iteration because it had to)."""
        if self.keep_on_device and V_index in self._resident:
            return self._resident[V_index]
        screen, valid, mask, origin, ray_dir, cam = self.Views[V_index]
        up = lambda t: t.to(self.device, non_blocking=True)  # noqa: E731
        view = (up(screen), up(valid), up(mask), up(origin), up(ray_dir), tuple(up(m) for m in cam))
        if self.keep_on_device:
            self._resident[V_index] = view
        return view

    def _cycle(self, index):
        index = list(index)
        while True:
            self.rng.shuffle(index)
            for i in index:
                yield int(i) % self.n_views

    def ray_view_generator(self):  # captured_data.py:60-74
        return self._cycle(np.arange(0, self.n_views, max(1, self.n_views // self.num_view)))

    def silh_view_generator(self):  # captured_data.py:76-82
        return self._cycle(np.arange(self.n_views))
