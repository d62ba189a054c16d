"""Quality evaluators and volume export behind the C ABI (evaluate.cu): vnrNeuralVolumeGetSSIM / GetTestingLoss /
DecodeInference / DecodeReference (api.h:130-131,139-140; core/network.cu:70-125,261-288,327-408,474-549)."""
import numpy as np
import pytest

import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
CFG = dict(n_levels=4, n_features=4, log2_hashmap=12, base_res=8, n_hidden=2)


def _trained(dims, steps=150):
    vol = vnr.NeuralVolume(vnr.model_json(**CFG), dims)
    gt = syn.make_volume(dims, seed=11)
    vol.set_groundtruth(gt)
    vol.init_params(5)
    vol.train(steps, batch=8192, fast_mode=True)
    return vol, gt


def _decoded(vol, dims):
    for _ in range(vol.num_blobs()):
        vol.decode_progressive()
    return vol.get_decoded()


def test_ssim_matches_the_oracle_per_window_and_in_the_mean():
    dims = (40, 33, 26)                      # odd sizes: ragged CTA tiles in x and y
    vol, gt = _trained(dims)
    dec = _decoded(vol, dims)
    got, gmap = vol.ssim(return_map=True)
    want, wmap = O.ssim(gt, dec, return_map=True)
    assert gmap.shape == wmap.shape == (20, 27, 34)
    assert np.array_equal(gmap, wmap)        # same fp32 operation order, every fused multiply-add explicit on both sides
    assert abs(got - want) <= 1e-9
    assert 0.0 < got < 1.0
    assert vol.ssim() == got                 # idempotent; no map requested


def test_ssim_is_one_when_the_network_is_exact_and_rejects_thin_volumes():
    # a network with zero weights decodes 0 everywhere; against an all-zero ground truth every window gives C1*C2/(C1*C2)
    dims = (16, 16, 16)
    vol = vnr.NeuralVolume(vnr.model_json(**CFG), dims)
    vol.set_groundtruth(np.zeros(dims[::-1], np.float32))
    vol.set_params_f16(np.zeros(vol.n_params, np.uint16))
    assert vol.ssim() == 1.0
    thin = vnr.NeuralVolume(vnr.model_json(**CFG), (16, 16, 6))
    thin.set_groundtruth(np.zeros((6, 16, 16), np.float32))
    thin.init_params(1)
    with pytest.raises(vnr.VnrError):
        thin.ssim()
    with pytest.raises(vnr.VnrError):         # no ground truth: "[error]: missing a reference volume."
        v2 = vnr.NeuralVolume(vnr.model_json(**CFG), dims); v2.init_params(1); v2.ssim()


def test_testing_loss_draws_the_next_sampler_batch():
    dims = (32, 32, 32)
    vol, gt = _trained(dims, steps=50)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p16 = vol.get_params_f16()
    n = 4096
    # the sampler stream after 50 training steps of 8192 samples (3 floats each)
    rng = O.Rng(1337); rng.uniform(50 * 8192 * 3)
    xyz, tgt = O.sample_batch(rng, n, gt, dims)
    want = float(np.mean(np.abs(O.decode(m, p16, xyz).astype(np.float64) - tgt.astype(np.float64))))
    got = vol.test_loss(n)
    assert abs(got - want) <= 2.0 ** -9       # decode tolerance (fp16 MLP) on a mean of |.|
    # the draw advanced the stream: the next call sees the next batch
    xyz2, tgt2 = O.sample_batch(rng, n, gt, dims)
    want2 = float(np.mean(np.abs(O.decode(m, p16, xyz2).astype(np.float64) - tgt2.astype(np.float64))))
    assert abs(vol.test_loss(n) - want2) <= 2.0 ** -9
    assert abs(want - want2) > 0 and vol.test_loss() > 0      # default batch 65536


@pytest.mark.parametrize("dims", [(32, 16, 9), (20, 10, 5)])     # 512 = 2 x 256 (no padding); 200 -> 256 (padded records)
def test_export_writes_the_reference_record_layout(tmp_path, dims):
    vol, gt = _trained(dims, steps=20)
    dec = _decoded(vol, dims)
    sl = dims[0] * dims[1]
    rec = (sl + 255) // 256 * 256
    for which, src in ((0, dec.reshape(-1)), (1, gt.reshape(-1))):
        path = tmp_path / f"vol{which}.bin"
        lo, hi = vol.export(path, which)
        raw = np.fromfile(path, dtype=np.float32)
        assert raw.size == rec * dims[2]                           # sizeof(float) * count * dims.z (network.cu:342)
        raw = raw.reshape(dims[2], rec)
        assert np.array_equal(raw[:, :sl].reshape(-1), src)
        if rec > sl and dims[2] > 1:                                # the padding of record z holds the first voxels of slice z+1
            nxt = src.reshape(dims[2], sl)[1:, : rec - sl]
            assert np.array_equal(raw[:-1, sl:], nxt)
        assert lo == raw.min() and hi == raw.max()
    with pytest.raises(vnr.VnrError):
        vol.export(tmp_path / "no_such_dir" / "x.bin", 0)
