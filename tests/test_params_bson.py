"""params.json (BSON) interchange with the reference: the golden file was written by the reference's own
serializer (nlohmann json.hpp from /root/reference/tcnn/dependencies, tools/make_golden_bson.cpp) in the
layout of NeuralVolume::save_params_to_json (core/network.cu:827-857)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import instantvnr_b200 as vnr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "params_ref_small.bson")
DIMS = (40, 24, 17)
N_PARAMS = 2048 + 256


def _expected_params():
    i = np.arange(N_PARAMS, dtype=np.int64)
    return (0x2000 + (i * 7) % 0x1C00).astype(np.uint16)


def _peek(blob):
    dx, dy, dz = C.c_int(), C.c_int(), C.c_int()
    model = C.c_char_p()
    rc = vnr.lib().vnr_params_peek(blob, C.c_size_t(len(blob)), C.byref(dx), C.byref(dy), C.byref(dz), C.byref(model))
    return rc, (dx.value, dy.value, dz.value), (model.value or b"").decode()


def test_peek_reads_reference_written_file_without_a_device():
    blob = open(GOLDEN, "rb").read()
    rc, dims, model = _peek(blob)
    assert rc == 0 and dims == DIMS
    m = json.loads(model)
    assert m["encoding"] == {"base_resolution": 4, "log2_hashmap_size": 6, "n_features_per_level": 2, "n_levels": 2, "otype": "HashGrid"}
    assert m["network"]["n_hidden_layers"] == 1 and m["loss"]["otype"] == "L1"


def test_peek_rejects_garbage_and_missing_volume_tag():
    rc, _, _ = _peek(b"\x05\x00\x00\x00\x00")          # empty document: no "volume"
    assert rc == -1 and b"volume dims" in vnr.lib().vnr_last_error()
    rc, _, _ = _peek(b"\xff\xff\xff\x7f" + b"junk" * 8)
    assert rc == -1
    rc, _, _ = _peek(open(GOLDEN, "rb").read()[:100])   # truncated
    assert rc == -1


@pytest.mark.gpu
def test_load_reference_file_and_write_it_back_byte_identical():
    blob = open(GOLDEN, "rb").read()
    rc, dims, model = _peek(blob)
    assert rc == 0
    vol = vnr.NeuralVolume(model, dims)
    vol.load_params(blob)
    assert np.array_equal(vol.get_params_f16(), _expected_params())
    md, vr, _ = vol.get_macrocell()
    assert md == (3, 2, 2)
    assert np.array_equal(vr, np.arange(24, dtype=np.float32) * 0.125 - 1.0)
    out = vol.save_params()
    assert out == blob                                   # same bytes as nlohmann::json::to_bson wrote


@pytest.mark.gpu
def test_load_errors():
    blob = open(GOLDEN, "rb").read()
    _, dims, model = _peek(blob)
    other = vnr.NeuralVolume(model, (32, 32, 32))
    with pytest.raises(vnr.VnrError) as e:
        other.load_params(blob)
    assert e.value.code == -1 and "mismatch data dimension" in str(e.value)
    fresh = vnr.NeuralVolume(model, dims)
    with pytest.raises(vnr.VnrError) as e:
        fresh.save_params()
    assert e.value.code == -4


@pytest.mark.gpu
def test_trained_volume_round_trips_and_renders_identically():
    from instantvnr_b200 import synthetic as syn
    dims = (32, 32, 32)
    cfg = vnr.model_json(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2)
    a = vnr.NeuralVolume(cfg, dims)
    a.set_groundtruth(syn.make_volume(dims, seed=5))
    a.init_params(3)
    rgb, alpha = syn.make_tfn(32)
    a.set_transfer_function(rgb, alpha)
    a.train(30, batch=4096, fast_mode=False)
    blob = a.save_params()
    rc, pdims, model = _peek(blob)
    assert rc == 0 and pdims == dims
    b = vnr.NeuralVolume(model, pdims)
    b.set_transfer_function(rgb, alpha)
    b.load_params(blob)
    assert np.array_equal(a.get_params_f16(), b.get_params_f16())
    ma, mb = a.get_macrocell(), b.get_macrocell()
    assert ma[0] == mb[0] and np.array_equal(ma[1], mb[1]) and np.array_equal(ma[2], mb[2])
    frames = []
    for v in (a, b):
        r = vnr.Renderer(v)
        r.set_size(40, 30)
        r.set_camera(*syn.default_camera(dims, 2))
        r.render()
        frames.append(r.map_frame())
    assert np.array_equal(frames[0], frames[1])
