"""Raw volume files streamed into HBM (vnr_volume_set_groundtruth_file): StaticSampler::load semantics
(core/samplers/neural_sampler.cpp:176-288): value range from the data or given, clamp((v - min) / (max - min), 0, 1)."""
import numpy as np
import pytest

import instantvnr_b200 as vnr

pytestmark = pytest.mark.gpu
CFG = vnr.model_json(n_levels=2, n_features=2, log2_hashmap=8, base_res=4, n_hidden=1)


def _voxel_centres(dims):
    dx, dy, dz = dims
    zz, yy, xx = np.meshgrid((np.arange(dz) + 0.5) / dz, (np.arange(dy) + 0.5) / dy, (np.arange(dx) + 0.5) / dx, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.float32)


@pytest.mark.parametrize("dtype,big_endian,offset", [("uint8", False, 0), ("uint16", True, 128), ("int16", False, 0), ("float32", False, 64),
                                                      ("float64", True, 0), ("int32", False, 0)])
def test_file_volume_equals_host_normalised_volume(tmp_path, dtype, big_endian, offset):
    dims = (40, 24, 17)
    rng = np.random.default_rng(3)
    n = dims[0] * dims[1] * dims[2]
    if np.dtype(dtype).kind == "f":
        raw = (rng.standard_normal(n) * 50 + 10).astype(dtype)
    else:
        info = np.iinfo(dtype)
        raw = rng.integers(max(info.min, -30000), min(info.max, 60000), n, dtype=dtype)
    path = tmp_path / "vol.raw"
    with open(path, "wb") as f:
        f.write(b"\x7f" * offset)
        f.write(raw.astype(np.dtype(dtype).newbyteorder(">" if big_endian else "<")).tobytes())
    vol = vnr.NeuralVolume(CFG, dims)
    lo, hi = vol.set_groundtruth_file(path, dtype, offset=offset, big_endian=big_endian)
    f32 = raw.astype(np.float32)
    assert (lo, hi) == (float(f32.min()), float(f32.max()))
    want = np.clip((f32 - np.float32(lo)) / (np.float32(hi) - np.float32(lo)), 0, 1).astype(np.float32)
    got = vol.sample_at(_voxel_centres(dims))          # the filter returns the voxel value at voxel centres
    assert np.array_equal(got, want)
    # a given range clamps
    lo2, hi2 = vol.set_groundtruth_file(path, dtype, offset=offset, big_endian=big_endian, value_range=(float(lo) + 5.0, float(hi) - 5.0))
    want2 = np.clip((f32 - np.float32(lo2)) / (np.float32(hi2) - np.float32(lo2)), 0, 1).astype(np.float32)
    assert np.array_equal(vol.sample_at(_voxel_centres(dims)), want2)
    assert want2.min() == 0.0 and want2.max() == 1.0


def test_multi_chunk_file_and_training_from_it(tmp_path):
    """a file larger than one 64 MiB staging chunk (uint8, 416^3 = 72 MB) streams through both pinned buffers"""
    dims = (416, 416, 416)
    n = dims[0] * dims[1] * dims[2]
    raw = (np.arange(n, dtype=np.uint64) * 2654435761 >> 13).astype(np.uint8)
    path = tmp_path / "big.raw"
    raw.tofile(path)
    vol = vnr.NeuralVolume(CFG, dims)
    lo, hi = vol.set_groundtruth_file(path, "uint8")
    assert (lo, hi) == (0.0, 255.0)
    idx = np.random.default_rng(0).integers(0, n, 50000)
    z, r = np.divmod(idx, dims[0] * dims[1]); y, x = np.divmod(r, dims[0])
    xyz = np.stack([(x + 0.5) / dims[0], (y + 0.5) / dims[1], (z + 0.5) / dims[2]], 1).astype(np.float32)
    assert np.array_equal(vol.sample_at(xyz), raw[idx].astype(np.float32) / np.float32(255.0))
    vol.init_params(1)
    vol.train(3, batch=4096)
    assert vol.stats()[0] == 3


def test_file_errors(tmp_path):
    vol = vnr.NeuralVolume(CFG, (16, 16, 16))
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_file(tmp_path / "missing.raw", "uint8")
    assert e.value.code == -1 and "cannot open" in str(e.value)
    short = tmp_path / "short.raw"
    short.write_bytes(b"\x01" * 100)
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_file(short, "uint8", value_range=(0, 255))
    assert e.value.code == -1 and "shorter" in str(e.value)
    const = tmp_path / "const.raw"
    const.write_bytes(b"\x05" * 4096)
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_file(const, "uint8")
    assert "empty value range" in str(e.value)
    with pytest.raises(vnr.VnrError) as e:
        vnr._check(vnr.lib().vnr_volume_set_groundtruth_file(vol._h, str(const).encode(), 9, 0, 0, 0, 1, None))   # float2: not a scalar type
    assert e.value.code == -3
