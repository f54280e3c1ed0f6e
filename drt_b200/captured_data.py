"""View-set loaders with the reference's `captured_data` interface.

Schema of the captured sets (captured_data.py:94-108, 136-149): per object one file with
`cam_proj [72,4,4]` (world->camera), `cam_k [3,3]`, `screen_position [72,N,3]` (measured 3-D screen point per
pixel, zero where nothing was measured), `mask [72,resy,resx]` (uint8 silhouette), and for the Point Grey sets
calibrated per-pixel `ray_origin/ray_dir [72,N,3]` (the Redmi sets derive rays from K and R, :149).
The .h5 files themselves are not distributed (README.md:18) and h5py is not in this image: the loader reads them with
the in-tree HDF5 reader (`h5lite`; h5py is used when present), or the same arrays from an .npz / any mapping, and keeps
the 72 views in pinned host memory exactly like the reference (:112-120).
"""
import numpy as np
import torch

from . import views as _views

Float = torch.float64


def process_mask(M):
    """Soft silhouette mask in [0,1] from a binary one: signed 1-px Euclidean distance ramp across the
    boundary, last row forced to 0.5 (captured_data.py:12-20; cv2.distanceTransform(DIST_L2, precise) ==
    scipy's exact EDT)."""
    from scipy.ndimage import distance_transform_edt as edt
    M = np.asarray(M)
    if M.max() == 255:
        M = M // 255
    assert M.max() == 1
    dist = edt(M).clip(0, 1) - (edt(1 - M) - 1).clip(0, 1)
    mask = (dist + 1) / 2
    mask[-1] = 0.5
    return mask


def open_view_file(path):
    """-> mapping with the schema above: `f[name][i]`, `f[name][:]` as the reference uses its h5py.File (captured_data.py:94-108).
    .npz: numpy; .h5 / .hdf5: h5py when it is installed, else the in-tree reader `drt_b200.h5lite` (old- and new-style groups,
    contiguous / chunked + gzip + shuffle numeric datasets -- what h5py writes by default)."""
    if str(path).endswith((".h5", ".hdf5")):
        try:
            import h5py
            return h5py.File(path, "r")
        except ImportError:
            from . import h5lite
            return h5lite.File(path)
    return np.load(path)


class CompactView:
    """One view in the lossless compact form the fused ray-loss step consumes (drt_b200.losses.ray_loss_view):
    `origin` [1,3] when every ray of the view starts at the same point (a pinhole view: captured_data.py:38
    expands ONE camera centre to [N,3]) else [N,3]; `ray_dir` [N,3]; `targets` = SparseTargets of the measured
    pixels (captured_data.py:104: valid = screen_pixel[:,0] != 0).  73 B per ray of the reference layout become
    24 B + 28 B per MEASURED pixel; nothing is rounded."""

    __slots__ = ("origin", "ray_dir", "targets", "mask", "camera_M", "image_size", "tile_beams")

    def __init__(self, origin, ray_dir, targets, mask=None, camera_M=None, image_size=None, tile_beams=None):
        self.origin, self.ray_dir, self.targets, self.mask, self.camera_M = origin, ray_dir, targets, mask, camera_M
        self.image_size = image_size  # (resy, resx) when the rays are whole scanline-ordered images (lets Q1 work on pixel tiles)
        self.tile_beams = tile_beams  # device-resident views only: losses.prepare_tile_beams of these rays (prepare_beams())

    def prepare_beams(self):
        """For a view that STAYS on the device (Data.keep_on_device): the per-tile direction intervals of its rays, computed once
        (drt_tile_beams); every later losses.ray_loss_view on it skips the beam pass's scan of the ray directions."""
        from .losses import prepare_tile_beams
        if self.tile_beams is None and self.ray_dir.is_cuda:
            self.tile_beams = prepare_tile_beams(self.origin, self.ray_dir, self.image_size)
        return self

    @staticmethod
    def from_reference_view(view, image_size=None):
        """(screen_pixel, valid, mask, origin, ray_dir, camera_M) as captured_data.Data.Views holds it -> CompactView."""
        from .losses import SparseTargets
        screen, valid, mask, origin, ray_dir, cam = view
        one = origin.shape[0] > 0 and bool((origin == origin[:1]).all())
        if image_size is not None and image_size[0] * image_size[1] != ray_dir.shape[0]:
            image_size = None
        return CompactView(origin[:1].clone() if one else origin, ray_dir, SparseTargets.from_dense(screen, valid), mask, cam, image_size)

    @staticmethod
    def concat(views):
        """Several single-origin views as ONE batch for one drt_ray_loss_step call: origin [n_views,3] (ray i starts
        at row i // rays_per_view), ray_dir concatenated view-major, target indices offset into the batch.  All views
        must have the same ray count and one origin row each."""
        from .losses import SparseTargets
        n = views[0].ray_dir.shape[0]
        if any(v.origin.shape[0] != 1 or v.ray_dir.shape[0] != n for v in views):
            raise ValueError("concat needs single-origin views of equal ray count")
        idx = torch.cat([v.targets.idx + k * n for k, v in enumerate(views)])
        size = views[0].image_size if all(v.image_size == views[0].image_size for v in views) else None
        return CompactView(torch.cat([v.origin for v in views]), torch.cat([v.ray_dir for v in views]),
                           SparseTargets(idx, torch.cat([v.targets.xyz for v in views])), image_size=size)

    def pin_memory(self):
        pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t  # noqa: E731
        return CompactView(pin(self.origin), pin(self.ray_dir), self.targets.pin_memory() if torch.cuda.is_available()
                           else self.targets, self.mask, self.camera_M, self.image_size)

    def to(self, device, non_blocking=True):
        up = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)  # noqa: E731
        cam = None if self.camera_M is None else tuple(up(m) for m in self.camera_M)
        return CompactView(up(self.origin), up(self.ray_dir), self.targets.to(device, non_blocking), up(self.mask), cam, self.image_size)

    def h2d_bytes(self):
        """bytes the ray path needs on the device (mask / camera_M belong to the silhouette path)"""
        return sum(t.numel() * t.element_size() for t in (self.origin, self.ray_dir, self.targets.idx, self.targets.xyz))


class CompactViews:
    """Mixin for the loaders: compact pinned host copies built once per view, uploaded per iteration."""

    def compact(self, V_index):
        cache = self.__dict__.setdefault("_compact", {})
        if V_index not in cache:
            size = (self.resy, self.resx) if hasattr(self, "resy") else None
            cache[V_index] = CompactView.from_reference_view(self.Views[V_index], size).pin_memory()
        return cache[V_index]

    def get_view_compact(self, V_index):
        return self.compact(V_index).to(self.device)

    def compact_batch(self, V_indices):
        """pinned host batch of several views (CompactView.concat), cached per index tuple"""
        key = tuple(int(i) for i in V_indices)
        cache = self.__dict__.setdefault("_compact_batches", {})
        if key not in cache:
            cache[key] = CompactView.concat([self.compact(i) for i in key]).pin_memory()
        return cache[key]


class Data(CompactViews):
    """captured_data.Data (captured_data.py:43-82): get_view uploads one view; shuffled infinite generators."""

    device = "cuda"
    keep_on_device = False  # True: upload every view once and serve it from HBM afterwards (8.6 GB for 72 views)

    def get_view(self, V_index):
        cache = self.__dict__.setdefault("_resident", {})
        if self.keep_on_device and V_index in cache:
            return cache[V_index]
        screen_pixel, valid, mask, origin, ray_dir, camera_M = self.Views[V_index]
        up = lambda t: t.to(self.device, non_blocking=True)  # noqa: E731
        view = (up(screen_pixel), up(valid), up(mask), up(origin), up(ray_dir), tuple(up(m) for m in camera_M))
        if self.keep_on_device:
            cache[V_index] = view
        return view

    def _cycle(self, index):
        index = list(index)
        while True:
            np.random.shuffle(index)
            for i in index:
                yield int(i) % 72

    def ray_view_generator(self):
        index = list(np.arange(0, 72, 72 // self.num_view))
        if self.name == "mouse":  # captured_data.py:66-68: the reference restricts the mouse to these views
            index = list(np.arange(-5, 10)) + list(np.arange(22, 40))
        return self._cycle(index)

    def silh_view_generator(self):
        return self._cycle(np.arange(72))

    def _load(self, data, calibrated_rays):
        pin = lambda a, dt=Float: torch.tensor(np.asarray(a), dtype=dt).pin_memory() if torch.cuda.is_available() \
            else torch.tensor(np.asarray(a), dtype=dt)  # noqa: E731
        K = np.asarray(data["cam_k"])
        K_inv = np.linalg.inv(K)
        self.Views = []
        for i in range(len(data["cam_proj"])):
            R = np.asarray(data["cam_proj"][i])
            R_inv = np.linalg.inv(R)
            screen = np.asarray(data["screen_position"][i]).reshape(-1, 3)
            valid = screen[:, 0] != 0
            mask = process_mask(np.asarray(data["mask"][i]))
            if calibrated_rays:
                origin, ray_dir = np.asarray(data["ray_origin"][i]), np.asarray(data["ray_dir"][i])
            else:
                o, d = _views.generate_ray(self.resy, self.resx, K_inv, R_inv)
                origin, ray_dir = o.numpy(), d.numpy()
            self.Views.append((pin(screen), pin(valid, torch.bool), pin(mask), pin(origin), pin(ray_dir),
                               (pin(R), pin(K), pin(R_inv), pin(K_inv))))


class Data_Pointgray(Data):
    """960x1280 sets with calibrated per-pixel rays (hand, mouse, dog, monkey) -- captured_data.py:85-120."""

    def __init__(self, HyperParams, path=None, data=None):
        self.resy, self.resx = 960, 1280
        self.num_view, self.name = HyperParams["num_view"], HyperParams["name"]
        self._load(data if data is not None else open_view_file(path), calibrated_rays=True)


class Data_Redmi(Data):
    """1080x1920 pinhole sets (tiger, pig, horse, rabbit) -- captured_data.py:122-165."""

    def __init__(self, HyperParams, path=None, data=None, res=(1080, 1920)):
        self.resy, self.resx = res
        self.num_view, self.name = HyperParams["num_view"], HyperParams["name"]
        self._load(data if data is not None else open_view_file(path), calibrated_rays=False)
