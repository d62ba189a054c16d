"""ctypes front-end of the CPU oracle (oracle/vnr_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by instantvnr_b200.
Parity: pinned to the reference's own tiny-cuda-nn and renderer sources, except the out-of-core sampler restatement
(PARITY UNPINNED for that one function group): see the header of vnr_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvnr_oracle.so")
_lib = None

FRAME_FLOATS = 64
FRAME_INTS = 16


def build(force=False):
    src = os.path.join(_HERE, "vnr_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_train_create.restype = C.c_void_p
        _lib.orc_train_step.restype = C.c_double
        _lib.orc_train_grads.restype = C.c_double
        _lib.orc_train_set_grads.restype = None
        _lib.orc_train_apply.restype = None
        _lib.orc_lcg_tea16_first.restype = C.c_float
        _lib.orc_grid_index.restype = C.c_uint32
        _lib.orc_hadd.restype = C.c_uint16
        _lib.orc_hadd_exact.restype = C.c_uint16
        _lib.orc_f16_selftest.restype = C.c_uint64
        _lib.orc_ssim.restype = C.c_double
    return _lib


def use_host_cores(n=None):
    """give the oracle's OpenMP loops `n` threads (default: the cores this process may run on); launchers like torchrun
    export OMP_NUM_THREADS=1"""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except Exception:
            n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(int(n)))
    return int(lib().orc_num_threads())


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class ModelCfg:
    """example-model.json subset that defines the decoder (L, F, log2T, base, per_level_scale,
    n_hidden_layers, width)."""

    def __init__(self, n_levels=8, n_features=8, log2_hashmap=19, base_res=16, per_level_scale=2.0,
                 n_hidden=4, width=64):
        self.L, self.F, self.log2T, self.base = n_levels, n_features, log2_hashmap, base_res
        self.pls, self.n_hidden, self.width = float(per_level_scale), n_hidden, width
        self._cfg = np.array([self.L, self.F, self.log2T, self.base, self.n_hidden, self.width], dtype=np.int32)
        offs = np.zeros(self.L + 1, dtype=np.uint32)
        scales = np.zeros(self.L, dtype=np.float32)
        res = np.zeros(self.L, dtype=np.uint32)
        n_mlp, n_grid, enc_pad = C.c_uint64(), C.c_uint64(), C.c_int()
        lib().orc_model_info(_p(self._cfg), C.c_float(self.pls), _p(offs), _p(scales), _p(res),
                             C.byref(n_mlp), C.byref(n_grid), C.byref(enc_pad))
        self.offsets, self.scales, self.res = offs, scales, res
        self.n_mlp, self.n_grid, self.enc_pad = n_mlp.value, n_grid.value, enc_pad.value
        self.n_params = self.n_mlp + self.n_grid

    @property
    def cfg(self):
        return _p(self._cfg)


def f32_to_f16(a):
    a = _f32(a)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().orc_f32_to_f16(_p(a), _p(out), C.c_size_t(a.size))
    return out


def f16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    out = np.empty(a.shape, dtype=np.float32)
    lib().orc_f16_to_f32(_p(a), _p(out), C.c_size_t(a.size))
    return out


def pcg32_uints(initstate, initseq, n, advance=0):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_pcg32_uints(C.c_uint64(initstate), C.c_uint64(initseq), C.c_int64(advance), _p(out), C.c_size_t(n))
    return out


def pcg32_floats(initstate, initseq, n, advance=0):
    out = np.empty(n, dtype=np.float32)
    lib().orc_pcg32_floats(C.c_uint64(initstate), C.c_uint64(initseq), C.c_int64(advance), _p(out), C.c_size_t(n))
    return out


class Rng:
    """The sampler's process-wide pcg32 (neural_sampler.cu:36: `static default_rng_t rng{1337}`)."""

    def __init__(self, seed=1337, seq=1):
        self.state = np.zeros(2, dtype=np.uint64)
        lib().orc_pcg32_seed(C.c_uint64(seed), C.c_uint64(seq), _p(self.state))

    def uniform(self, n, lower=0.0, upper=1.0):
        out = np.empty(n, dtype=np.float32)
        lib().orc_random_uniform(_p(self.state), C.c_size_t(n), _p(out), C.c_float(lower), C.c_float(upper))
        return out


def grid_index(hashmap_size, res, x, y, z):
    return lib().orc_grid_index(C.c_uint32(hashmap_size), C.c_uint32(res), C.c_uint32(x), C.c_uint32(y), C.c_uint32(z))


def init_params(m, seed=1337):
    p32 = np.empty(m.n_params, dtype=np.float32)
    p16 = np.empty(m.n_params, dtype=np.uint16)
    lib().orc_init_params(m.cfg, C.c_float(m.pls), C.c_uint32(seed), _p(p32), _p(p16))
    return p32, p16


def encode(m, params_f16, coords):
    coords = _f32(coords).reshape(-1, 3)
    out = np.empty((coords.shape[0], m.enc_pad), dtype=np.uint16)
    lib().orc_encode(m.cfg, C.c_float(m.pls), _p(params_f16), _p(coords), C.c_size_t(coords.shape[0]), _p(out))
    return out


def decode(m, params_f16, coords, acc_mode=0):
    coords = _f32(coords).reshape(-1, 3)
    out = np.empty(coords.shape[0], dtype=np.float32)
    lib().orc_decode(m.cfg, C.c_float(m.pls), _p(params_f16), _p(coords), C.c_size_t(coords.shape[0]), _p(out), C.c_int(acc_mode))
    return out


def mlp(m, params_f16, enc, acc_mode=0):
    enc = np.ascontiguousarray(enc, dtype=np.uint16).reshape(-1, m.enc_pad)
    out = np.empty(enc.shape[0], dtype=np.float32)
    lib().orc_mlp(m.cfg, C.c_float(m.pls), _p(params_f16), _p(enc), C.c_size_t(enc.shape[0]), _p(out), C.c_int(acc_mode))
    return out


def sample_batch(rng, n, volume, dims, lower=(0, 0, 0), upper=(1, 1, 1), tex_round=0):
    volume = _f32(volume)
    d = np.array(dims, dtype=np.int32)
    lo, up = _f32(lower), _f32(upper)
    coords = np.empty((n, 3), dtype=np.float32)
    targets = np.empty(n, dtype=np.float32)
    lib().orc_sample_batch(_p(rng.state), C.c_size_t(n), _p(volume), _p(d), _p(lo), _p(up), C.c_int(tex_round), _p(coords), _p(targets))
    return coords, targets


def tex3d(volume, dims, coords, tex_round=0):
    volume = _f32(volume)
    coords = _f32(coords).reshape(-1, 3)
    d = np.array(dims, dtype=np.int32)
    out = np.empty(coords.shape[0], dtype=np.float32)
    lib().orc_tex3d(_p(volume), _p(d), _p(coords), C.c_size_t(coords.shape[0]), C.c_int(tex_round), _p(out))
    return out


def macrocell_dims(dims):
    return tuple((int(d) + 15) // 16 for d in dims)


def macrocell_update_explicit(coords, values, dims, mc):
    coords = _f32(coords).reshape(-1, 3)
    values = _f32(values)
    d = np.array(dims, dtype=np.int32)
    md = np.array(macrocell_dims(dims), dtype=np.int32)
    lib().orc_macrocell_update_explicit(_p(coords), _p(values), C.c_size_t(coords.shape[0]), _p(d), _p(md), _p(mc))


def macrocell_update_implicit(volume, dims):
    volume = _f32(volume)
    d = np.array(dims, dtype=np.int32)
    mdt = macrocell_dims(dims)
    md = np.array(mdt, dtype=np.int32)
    mc = np.zeros(2 * mdt[0] * mdt[1] * mdt[2], dtype=np.float32)
    lib().orc_macrocell_update_implicit(_p(volume), _p(d), _p(md), _p(mc))
    return mc


def macrocell_max_opacity(mc, alphas, lo=0.0, hi=1.0):
    mc = _f32(mc)
    alphas = _f32(alphas)
    cells = mc.size // 2
    out = np.empty(cells, dtype=np.float32)
    lib().orc_macrocell_max_opacity(_p(mc), C.c_size_t(cells), _p(alphas), C.c_int(alphas.size), C.c_float(lo), C.c_float(hi), _p(out))
    return out


def ooc_sample(state_inc, n, first_voxel, length, block_rows, raw_f32, dims, vmin, vmax):
    """OutOfCoreSampler::sample restated: returns (coords[n,3], values[n], bound violations); state_inc is advanced by 5 n."""
    first_voxel = np.ascontiguousarray(first_voxel, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    raw = _f32(raw_f32)
    d = np.array(dims, dtype=np.int32)
    coords = np.empty((n, 3), dtype=np.float32)
    values = np.empty(n, dtype=np.float32)
    lib().orc_ooc_sample.restype = C.c_uint64
    bad = lib().orc_ooc_sample(_p(state_inc), C.c_size_t(n), _p(first_voxel), _p(length), C.c_uint32(first_voxel.size), C.c_int(block_rows),
                               _p(raw), _p(d), C.c_float(vmin), C.c_float(vmax), _p(coords), _p(values))
    return coords, values, int(bad)


class Frame:
    """Per-frame constants (LaunchParams / DeviceVolume subset, instantvnr_types.h:89-149)."""

    def __init__(self, dims, width, height, cam_from, cam_at, cam_up, fovy=60.0, sampling_rate=1.0,
                 tfn_range=(0.0, 1.0), frame_index=1, n_iters=16, tex_round=0, shade_mode=0, light_dir=None):
        self.f = np.zeros(FRAME_FLOATS, dtype=np.float32)
        self.i = np.zeros(FRAME_INTS, dtype=np.int32)
        d = np.array(dims, dtype=np.int32)
        lib().orc_frame_setup(_p(_f32(cam_from)), _p(_f32(cam_at)), _p(_f32(cam_up)), C.c_float(fovy), C.c_int(width), C.c_int(height),
                              _p(d), C.c_float(sampling_rate), _p(_f32(tfn_range)), _p(self.f))
        md = macrocell_dims(dims)
        self.i[:10] = [width, height, frame_index, n_iters, tex_round, md[0], md[1], md[2], 0, 0]
        self.dims = tuple(int(x) for x in dims)
        self.width, self.height = width, height
        # shaded modes (0 none, 1 gradient shading, 2 single-shade heuristic + shadow pass); light_dir is the renderer's
        # persistent light direction before this frame's sign correction (renderer.cpp:98-101)
        self.light_dir = np.zeros(3, dtype=np.float32)
        lib().orc_frame_shading(_p(self.f), _p(self.i), C.c_int(shade_mode), _p(d), _p(_f32(light_dir)) if light_dir is not None else None,
                                _p(self.light_dir))

    def set_tfn_sizes(self, n_color, n_alpha):
        self.i[8], self.i[9] = n_color, n_alpha


def render(m, params_f16, frame, mc_max_opacity, colors_rgba, alphas, acc_mode=0, volume=None, jitter_mode=0,
           accum=None):
    """Sample-streaming marcher.  volume given -> ground-truth sampling instead of the network."""
    colors_rgba = _f32(colors_rgba).reshape(-1, 4)
    alphas = _f32(alphas)
    frame.set_tfn_sizes(colors_rgba.shape[0], alphas.size)
    npix = frame.width * frame.height
    if accum is None:
        accum = np.zeros((npix, 4), dtype=np.float32)
    out = np.zeros((npix, 4), dtype=np.float32)
    stats = np.zeros(4, dtype=np.uint64)
    mc_max_opacity = _f32(mc_max_opacity)
    vol = _f32(volume) if volume is not None else None
    gd = np.array(frame.dims, dtype=np.int32)
    lib().orc_render(m.cfg, C.c_float(m.pls), _p(params_f16), C.c_int(acc_mode), _p(frame.f), _p(frame.i), _p(mc_max_opacity),
                     _p(colors_rgba), _p(alphas), C.c_int(0 if vol is None else 1), _p(vol), _p(gd), C.c_int(jitter_mode),
                     _p(accum), _p(out), _p(stats))
    return out.reshape(frame.height, frame.width, 4), accum, {"rays_hit": int(stats[0]), "samples_decoded": int(stats[1]),
                                                              "samples_composited": int(stats[2]), "rounds": int(stats[3])}


def render_single_kernel(frame, mc_max_opacity, colors_rgba, alphas, volume, jitter_mode=0, accum=None):
    """The single-kernel marcher of the decoding / SimpleVolume in-shader modes (method_raymarching.cu:400-545)."""
    colors_rgba = _f32(colors_rgba).reshape(-1, 4)
    alphas = _f32(alphas)
    frame.set_tfn_sizes(colors_rgba.shape[0], alphas.size)
    npix = frame.width * frame.height
    if accum is None:
        accum = np.zeros((npix, 4), dtype=np.float32)
    out = np.zeros((npix, 4), dtype=np.float32)
    stats = np.zeros(4, dtype=np.uint64)
    mc_max_opacity = _f32(mc_max_opacity)
    vol = _f32(volume)
    gd = np.array(frame.dims, dtype=np.int32)
    lib().orc_render_single_kernel(_p(frame.f), _p(frame.i), _p(mc_max_opacity), _p(colors_rgba), _p(alphas), _p(vol), _p(gd),
                                   C.c_int(jitter_mode), _p(accum), _p(out), _p(stats))
    return out.reshape(frame.height, frame.width, 4), accum, {"rays_hit": int(stats[0]), "samples_decoded": int(stats[1]),
                                                              "samples_composited": int(stats[2]), "rounds": int(stats[3])}


def render_pathtracing(frame, mc_max_opacity, colors_rgba, alphas, volume=None, m=None, params_f16=None, streaming=True, density_scale=1.0,
                       light_ambient=1.5, light_rgb=(1.0, 1.0, 1.0), acc_mode=0, accum=None):
    """The path tracer (method_pathtracing.cu): streaming=True -> the sample-streaming state machine (mode 14 / 15 on a network),
    False -> the single-kernel tracer (mode 13 / 15 on a SimpleVolume).  volume given -> trilinear lookups, else the network."""
    colors_rgba = _f32(colors_rgba).reshape(-1, 4)
    alphas = _f32(alphas)
    frame.set_tfn_sizes(colors_rgba.shape[0], alphas.size)
    npix = frame.width * frame.height
    if accum is None:
        accum = np.zeros((npix, 4), dtype=np.float32)
    out = np.zeros((npix, 4), dtype=np.float32)
    stats = np.zeros(2, dtype=np.uint64)
    mc_max_opacity = _f32(mc_max_opacity)
    vol = _f32(volume) if volume is not None else None
    gd = np.array(frame.dims, dtype=np.int32)
    lights = _f32([density_scale, light_ambient, *light_rgb, *frame.f[38:41]])
    cfg = m.cfg if m is not None else None
    lib().orc_render_pathtracing(cfg, C.c_float(m.pls if m is not None else 2.0), _p(params_f16) if params_f16 is not None else None, C.c_int(acc_mode),
                                 _p(frame.f), _p(frame.i), _p(mc_max_opacity), _p(colors_rgba), _p(alphas), C.c_int(0 if vol is None else 1),
                                 _p(vol), _p(gd), C.c_int(1 if streaming else 0), _p(lights), _p(accum), _p(out), _p(stats))
    return out.reshape(frame.height, frame.width, 4), accum, {"rays_hit": int(stats[0]), "samples_decoded": int(stats[1])}


def rays(frame):
    out = np.empty((frame.width * frame.height, 8), dtype=np.float32)
    lib().orc_rays(_p(frame.f), _p(frame.i), _p(out))
    return out


def classify(frame, colors_rgba, alphas, values, dts):
    colors_rgba = _f32(colors_rgba).reshape(-1, 4)
    alphas = _f32(alphas)
    frame.set_tfn_sizes(colors_rgba.shape[0], alphas.size)
    values, dts = _f32(values), _f32(dts)
    out = np.empty((values.size, 4), dtype=np.float32)
    lib().orc_classify(_p(frame.f), _p(frame.i), _p(colors_rgba), _p(alphas), _p(values), _p(dts), C.c_size_t(values.size), _p(out))
    return out


def ssim(reference, prediction, return_map=False):
    """Mean SSIM of two [z][y][x] float volumes (compute_ssim / get_mssim, core/network.cu:70-125,474-549)."""
    a, b = _f32(reference), _f32(prediction)
    assert a.shape == b.shape and a.ndim == 3
    dims = np.array(a.shape[::-1], dtype=np.int32)
    out = np.empty(tuple(d - 6 for d in a.shape), dtype=np.float32) if return_map else None
    v = lib().orc_ssim(_p(a), _p(b), _p(dims), _p(out))
    return (v, out) if return_map else v


def lcg_tea16_first(v0, v1):
    return float(lib().orc_lcg_tea16_first(C.c_uint32(v0), C.c_uint32(v1)))


DEFAULT_HYPER = dict(lr=5e-3, beta1=0.9, beta2=0.999, eps=1e-15, l2_reg=1e-6, decay_base=0.99, decay_start=2000,
                     decay_interval=1000)   # example-model.json:2-15


class Trainer:
    def __init__(self, m, params_f32, hyper=None):
        h = dict(DEFAULT_HYPER)
        h.update(hyper or {})
        hv = np.array([h["lr"], h["beta1"], h["beta2"], h["eps"], h["l2_reg"], h["decay_base"], h["decay_start"],
                       h["decay_interval"]], dtype=np.float32)
        self.m = m
        self.h = C.c_void_p(lib().orc_train_create(m.cfg, C.c_float(m.pls), _p(_f32(params_f32)), _p(hv)))

    def step(self, coords, targets, acc_mode=0, grad_mode=0, do_step=True):
        coords = _f32(coords).reshape(-1, 3)
        targets = _f32(targets)
        return float(lib().orc_train_step(self.h, _p(coords), _p(targets), C.c_size_t(coords.shape[0]), C.c_int(acc_mode),
                                          C.c_int(grad_mode), C.c_int(1 if do_step else 0)))

    def grads_only(self, coords, targets, n_global, acc_mode=0, grad_mode=0):
        """forward + loss + backward with the loss normalised by n_global (data-parallel rank share); returns the
        rank's share of the loss; gradients are read with grads()."""
        coords = _f32(coords).reshape(-1, 3)
        targets = _f32(targets)
        return float(lib().orc_train_grads(self.h, _p(coords), _p(targets), C.c_size_t(coords.shape[0]), C.c_size_t(n_global),
                                           C.c_int(acc_mode), C.c_int(grad_mode)))

    def set_wgrad_slices(self, n):
        """grad_mode 2: number of K-slices of the half-accumulated weight gradients (the product: its CTA count, one per SM)"""
        lib().orc_train_set_wgrad_slices(self.h, C.c_int(n))

    def apply(self, grads):
        """optimizer step (ExponentialDecay + Adam) on externally provided (e.g. all-reduced) gradients"""
        g = _f32(grads)
        assert g.size == self.m.n_params
        lib().orc_train_set_grads(self.h, _p(g))
        lib().orc_train_apply(self.h)

    def params(self):
        p16 = np.empty(self.m.n_params, dtype=np.uint16)
        p32 = np.empty(self.m.n_params, dtype=np.float32)
        lib().orc_train_get_params(self.h, _p(p16), _p(p32))
        return p16, p32

    def grads(self):
        g = np.empty(self.m.n_params, dtype=np.float32)
        lib().orc_train_get_grads(self.h, _p(g))
        return g

    def __del__(self):
        try:
            lib().orc_train_destroy(self.h)
        except Exception:
            pass
