"""drt_b200.h5lite -- the HDF5 subset the captured view sets need (captured_data.py:94-108, 136-149 read them
with h5py, which this image does not have).

The reader is checked against a file written by libhdf5 itself: tests/golden/h5/testhdf5_7.4_GLNX86.mat is a MATLAB v7.3 file
(= an HDF5 file behind a 512-byte user block) from scipy's own test data (scipy/io/matlab/tests/data, BSD licence), whose one
variable is `testdouble = 0:pi/4:2*pi` like its v4-v7 siblings in that directory.  Everything else is written by h5lite.write_h5
and read back."""
import os

import numpy as np
import pytest

from drt_b200 import h5lite

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reads_a_file_written_by_libhdf5():
    f = h5lite.File(os.path.join(HERE, "golden", "h5", "testhdf5_7.4_GLNX86.mat"))
    assert f.base == 512 and (f.O, f.L) == (8, 8)                 # superblock found behind the user block
    assert list(f.keys()) == ["testdouble"] and "testdouble" in f and "nope" not in f
    ds = f["testdouble"]
    assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8") and ds.chunks is None
    assert np.array_equal(ds[:, 0], np.arange(9) * (np.pi / 4))   # MATLAB stores column-major: [1,9] arrives as (9,1)
    assert ds[3, 0] == 3 * np.pi / 4 and np.asarray(ds).shape == (9, 1)
    with pytest.raises(KeyError):
        f["nope"]
    f.close()


@pytest.mark.parametrize("mode", ["contiguous", "chunked", "gzip"])
def test_roundtrip_every_layout_and_index_form(tmp_path, mode):
    rng = np.random.default_rng(1)
    arrays = {
        "cam_proj": rng.normal(size=(7, 4, 4)),
        "cam_k": rng.normal(size=(3, 3)),
        "screen_position": rng.normal(size=(7, 150, 3)) * (rng.uniform(size=(7, 150, 1)) > 0.4),
        "mask": (rng.uniform(size=(7, 20, 30)) > 0.5).astype(np.uint8) * 255,
        "ids": rng.integers(-1000, 1000, size=(300,), dtype=np.int32),
        "f32": rng.normal(size=(5, 6)).astype(np.float32),
        "flag": rng.uniform(size=(11,)) > 0.5,
        "be": rng.normal(size=(4, 2)).astype(">f8"),
        "scalar": np.float64(2.5),
        "empty": np.zeros((0, 3)),
    }
    chunks = None if mode == "contiguous" else {"screen_position": (2, 64, 3), "mask": (1, 8, 16), "cam_proj": (3, 4, 4),
                                                 "ids": (4,)}  # ids: 75 chunks -> a two-level chunk B-tree
    p = str(tmp_path / "set.h5")
    h5lite.write_h5(p, arrays, chunks=chunks, compression="gzip" if mode == "gzip" else None, userblock=0 if mode != "chunked" else 512)
    with h5lite.File(p) as f:
        assert sorted(f.keys()) == sorted(arrays)
        for name, a in arrays.items():
            ds = f[name]
            ref = a.astype(np.uint8) if a.dtype == np.bool_ else np.asarray(a)
            assert ds.shape == ref.shape and ds.dtype.itemsize == ref.dtype.itemsize, name
            assert np.array_equal(np.asarray(ds), ref), name
            if ref.ndim == 0:
                assert ds[()] == ref
                continue
            if ref.shape[0] == 0:
                continue
            assert np.array_equal(ds[:], ref) and np.array_equal(ds[1], ref[1]) and np.array_equal(ds[-1], ref[-1]), name
            assert np.array_equal(ds[1:4], ref[1:4]) and np.array_equal(ds[::2], ref[::2]) and np.array_equal(ds[...], ref), name
            if ref.ndim == 3:
                assert np.array_equal(ds[2, 1:9], ref[2, 1:9]) and np.array_equal(ds[:, 3, 1], ref[:, 3, 1]), name
                assert np.array_equal(ds[4, ..., 0], ref[4, ..., 0]), name
            with pytest.raises(IndexError):
                ds[ref.shape[0]]
        if chunks:
            assert f["screen_position"].chunks == (2, 64, 3) and f["cam_k"].chunks is None
            assert len(f["ids"]._chunks()) == 75
        if mode == "gzip":
            assert [fl[0] for fl in f["mask"]._filters] == [2, 1]   # shuffle, then deflate
            assert os.path.getsize(p) < sum(np.asarray(a).nbytes for a in arrays.values())


def test_unsupported_features_raise(tmp_path):
    p = str(tmp_path / "x.h5")
    open(p, "wb").write(b"not hdf5 at all" * 100)
    with pytest.raises(h5lite.H5FormatError):
        h5lite.File(p)
    with pytest.raises(NotImplementedError):
        h5lite.write_h5(p, {"s": np.array(["a", "b"])})
    with pytest.raises(NotImplementedError):
        h5lite._parse_datatype(bytes([0x16, 0, 0, 0, 8, 0, 0, 0]))   # compound


def test_captured_set_from_h5_equals_npz(tmp_path):
    """The loaders (captured_data.py:85-165) over an .h5 capture file: same Views as from the .npz twin, view by view."""
    import torch
    from drt_b200 import captured_data as cd, meshgen, views
    rng = np.random.default_rng(0)
    v, _ = meshgen.icosahedron()
    cams = views.turntable_cameras(v, 12, 16, 4)
    screen = rng.normal(size=(4, 12 * 16, 3))
    screen[:, ::5] = 0
    masks = (rng.uniform(size=(4, 12, 16)) > 0.5).astype(np.uint8)
    masks[:, 0, 0] = 1
    rays = [views.generate_ray(12, 16, c[3], c[2]) for c in cams]
    arrays = dict(cam_proj=np.stack([c[0] for c in cams]), cam_k=cams[0][1], screen_position=screen, mask=masks,
                  ray_origin=np.stack([r[0].numpy() for r in rays]), ray_dir=np.stack([r[1].numpy() for r in rays]))
    np.savez(str(tmp_path / "set.npz"), **arrays)
    h5lite.write_h5(str(tmp_path / "set.h5"), arrays, chunks={"screen_position": (1, 100, 3), "ray_dir": (1, 192, 3)}, compression="gzip")
    for cls, kw in ((cd.Data_Redmi, {"res": (12, 16)}), (cd.Data_Pointgray, {})):
        a = cls({"num_view": 4, "name": "horse"}, path=str(tmp_path / "set.npz"), **kw)
        b = cls({"num_view": 4, "name": "horse"}, path=str(tmp_path / "set.h5"), **kw)
        assert len(a.Views) == len(b.Views) == 4
        for va, vb in zip(a.Views, b.Views):
            for ta, tb in zip(va[:5], vb[:5]):
                assert torch.equal(ta, tb)
            for ta, tb in zip(va[5], vb[5]):
                assert torch.equal(ta, tb)
