"""GPU parity: fused hash-grid + tcgen05 MLP decode vs the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import instantvnr_b200 as vnr
import oracle as O

pytestmark = pytest.mark.gpu

# fp16 tolerance of decoded values: |d| <= 2^-9 * max(1, |y|) against the oracle that
# accumulates like a tensor core with an fp32 accumulator (acc_mode 0); reported vs acc_mode 1.
TOL = 2.0 ** -9


def _scaled_params(m, seed, grid_scale=2000.0):
    """Random-init parameters with the grid scaled up so decoded values are O(0.1) rather than 1e-5."""
    p32, _ = O.init_params(m, seed)
    p32 = p32.copy()
    p32[m.n_mlp:] *= grid_scale
    return O.f32_to_f16(p32)


CONFIGS = [
    dict(n_levels=8, n_features=8, log2_hashmap=19, base_res=16, n_hidden=4),   # example-model.json
    dict(n_levels=16, n_features=2, log2_hashmap=19, base_res=16, n_hidden=2),  # BASELINE.json text variant
    dict(n_levels=8, n_features=4, log2_hashmap=15, base_res=8, n_hidden=3),
    dict(n_levels=16, n_features=1, log2_hashmap=14, base_res=4, n_hidden=1),
    dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=16, n_hidden=4, per_level_scale=1.5),
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_decode_matches_oracle(cfg):
    m = O.ModelCfg(cfg["n_levels"], cfg["n_features"], cfg["log2_hashmap"], cfg["base_res"], cfg.get("per_level_scale", 2.0), cfg["n_hidden"])
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), (64, 64, 64))
    assert vol.n_params == m.n_params and vol.n_mlp_params == m.n_mlp
    p16 = _scaled_params(m, 7)
    vol.set_params_f16(p16)
    rng = np.random.default_rng(3)
    n = 128 * 37 + 53          # ragged: last tile partially filled
    xyz = rng.random((n, 3), dtype=np.float32)
    xyz[:8] = [[0, 0, 0], [1, 1, 1], [0.999999, 0.5, 0.25], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.25, 0.75, 1.0]]
    got, enc = vol.decode_debug(xyz)
    want_enc = O.encode(m, p16, xyz)
    # hash-grid gather: bit exact (index math, fp16 accumulation order)
    assert np.array_equal(enc, want_enc)
    want = O.decode(m, p16, xyz, acc_mode=0)
    want_ref_like = O.decode(m, p16, xyz, acc_mode=1)
    err = np.abs(got - want)
    assert np.all(err <= TOL * np.maximum(1.0, np.abs(want))), f"max err {err.max()}"
    assert np.abs(want).max() > 1e-3          # the test is not vacuous
    # against the fp16-accumulating emulation of the reference's wmma path the bound is the same
    err1 = np.abs(got - want_ref_like)
    assert np.all(err1 <= 4 * TOL * np.maximum(1.0, np.abs(want_ref_like))), f"max err vs fp16-acc {err1.max()}"
    # decode_host == decode_debug
    assert np.array_equal(vol.decode_host(xyz), got)


def test_decode_fresh_init_params_match_oracle():
    """vnr_volume_init_params restates Trainer::initialize_params: same fp16 blob as the oracle."""
    m = O.ModelCfg()
    vol = vnr.NeuralVolume(vnr.example_model_json(), (32, 32, 32))
    vol.init_params(1337)
    _, p16 = O.init_params(m, 1337)
    assert np.array_equal(vol.get_params_f16(), p16)


def test_decode_empty_and_single():
    vol = vnr.NeuralVolume(vnr.example_model_json(), (32, 32, 32))
    vol.init_params(1)
    assert vol.decode_host(np.zeros((0, 3), np.float32)).shape == (0,)
    m = O.ModelCfg()
    p16 = vol.get_params_f16()
    x = np.array([[0.3, 0.6, 0.9]], np.float32)
    assert abs(vol.decode_host(x)[0] - O.decode(m, p16, x)[0]) <= TOL


def test_decode_large_property():
    """2^22 samples: determinism and permutation equivariance (size-independent properties)."""
    vol = vnr.NeuralVolume(vnr.example_model_json(), (256, 256, 256))
    m = O.ModelCfg()
    vol.set_params_f16(_scaled_params(m, 11))
    rng = np.random.default_rng(5)
    n = 1 << 22
    xyz = rng.random((n, 3), dtype=np.float32)
    a = vol.decode_host(xyz)
    b = vol.decode_host(xyz)
    assert np.array_equal(a, b)
    perm = rng.permutation(n)
    c = vol.decode_host(xyz[perm])
    assert np.array_equal(c, a[perm])
    idx = rng.choice(n, 4096, replace=False)
    want = O.decode(m, vol.get_params_f16(), xyz[idx])
    assert np.all(np.abs(a[idx] - want) <= TOL * np.maximum(1.0, np.abs(want)))
