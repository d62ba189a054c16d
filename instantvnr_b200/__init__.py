"""instantvnr_b200 -- B200-native (sm_100a) implementation of instantvnr's hot path.

This Python package is only the harness around the product: it loads the C-ABI shared
library (csrc -> _build/libvnr_b200.so, declared in include/vnr_c.h) through ctypes and
offers thin object wrappers used by tests/, bench.py and the multi-GPU (torch.distributed)
drivers.  The product itself is the CUDA/C++ library; there is no CPU or PyTorch fallback:
importing works without a GPU (so the symbol table can be checked), every compute call
raises VnrError if the CUDA library or a device is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "_build", "libvnr_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "vnr_c.h")
EXAMPLE_MODEL = os.path.join(_HERE, "configs", "example-model.json")

VNR_OK = 0


class VnrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[vnr error {code}] {msg}")
        self.code = code


def build(verbose=False):
    """Compile libvnr_b200.so in-tree (nvcc, sm_100a)."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VnrError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no fallback path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vnr_last_error.restype = C.c_char_p
        _lib.vnr_map_frame.restype = C.POINTER(C.c_float)
        _lib.vnr_volume_release.restype = None
        _lib.vnr_renderer_release.restype = None
        _lib.vnr_peer_barrier_release.restype = None
    return _lib


def _check(code):
    if code != VNR_OK:
        raise VnrError(code, lib().vnr_last_error().decode("utf-8", "replace"))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):          # torch tensor
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _stream(s):
    """None -> NULL (the handle's own stream); 0 (torch's default stream) -> cudaStreamLegacy."""
    if s is None:
        return None
    s = int(s)
    return C.c_void_p(1 if s == 0 else s)


def device_count():
    return int(lib().vnr_device_count())


def probe_memory(kind, table_bytes, n_ops, repeats=5):
    """Independent memory-system microbenchmark (csrc/probe.cu): kind "loads" / "reds" / "copy" -> (best ms, mean ms)."""
    k = {"loads": 0, "reds": 1, "copy": 2, "loads32": 3, "host_scanline": 4, "host_tiles8x4": 5, "host_scanline32": 6, "host_tiles16x2": 7, "mixed": 8}.get(kind, kind)
    best, mean = C.c_float(), C.c_float()
    _check(lib().vnr_probe_memory(C.c_int(k), C.c_size_t(int(table_bytes)), C.c_size_t(int(n_ops)), C.c_int(repeats), C.byref(best), C.byref(mean)))
    return best.value, mean.value


def probe_levels(level_entries, n_samples, repeats=5):
    """csrc/probe.cu: n_samples x 8 random 16-byte loads per level of a table laid out like a hash grid -> (best ms, mean ms)"""
    le = np.ascontiguousarray(level_entries, dtype=np.uint32)
    best, mean = C.c_float(), C.c_float()
    _check(lib().vnr_probe_levels(_ptr(le), C.c_int(le.size), C.c_size_t(int(n_samples)), C.c_int(repeats), C.byref(best), C.byref(mean)))
    return best.value, mean.value


def hash_grid_level_entries(n_levels=8, log2_hashmap=19, base_res=16, per_level_scale=2.0):
    """entries per level of the hash grid (tcnn encodings/grid.h:591-611: dense while (res)^3 fits, rounded up to 8, capped at T)"""
    out = []
    for l in range(n_levels):
        scale = base_res * per_level_scale ** l - 1.0
        res = int(np.ceil(scale)) + 1
        out.append(int(min(-(-res ** 3 // 8) * 8, 1 << log2_hashmap)))
    return out


def example_model_json():
    with open(EXAMPLE_MODEL) as f:
        return f.read()


def model_json(n_levels=8, n_features=8, log2_hashmap=19, base_res=16, n_hidden=4, per_level_scale=None):
    """A model config in the reference's example-model.json schema."""
    import json
    cfg = json.loads("\n".join(l for l in example_model_json().splitlines() if not l.strip().startswith("//")))
    cfg["encoding"].update(n_levels=n_levels, n_features_per_level=n_features, log2_hashmap_size=log2_hashmap,
                           base_resolution=base_res)
    if per_level_scale is not None:
        cfg["encoding"]["per_level_scale"] = per_level_scale
    cfg["network"]["n_hidden_layers"] = n_hidden
    return json.dumps(cfg)


class Comm:
    """vnr_comm_t: multi-GPU behind the same calls (include/vnr_c.h).  `Comm.init_rank` = one process per GPU,
    `Comm.init_local` = one process driving n devices (returns one communicator per rank)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def init_rank(rank, world, name):
        h = C.c_void_p()
        _check(lib().vnr_comm_init_rank(C.c_int(rank), C.c_int(world), str(name).encode(), C.byref(h)))
        return Comm(h)

    @staticmethod
    def init_local(n_devices):
        arr = (C.c_void_p * n_devices)()
        _check(lib().vnr_comm_init(C.c_int(n_devices), arr))
        return [Comm(C.c_void_p(a)) for a in arr]

    def info(self):
        r, w, d = C.c_int(), C.c_int(), C.c_int()
        _check(lib().vnr_comm_info(self._h, C.byref(r), C.byref(w), C.byref(d)))
        return r.value, w.value, d.value

    def set_device(self):
        _check(lib().vnr_comm_set_device(self._h))

    def barrier(self):
        _check(lib().vnr_comm_barrier(self._h))

    def close(self):
        if self._h:
            lib().vnr_comm_release.restype = None
            lib().vnr_comm_release(self._h)
            self._h = None


class NeuralVolume:
    """vnrVolume (neural) -- api.h:122-143."""

    uniforms_per_sample = 3          # StaticSampler: x, y, z; the out-of-core sampler draws 5 (+ slab and voxel selectors)

    def __init__(self, model_json_text, dims):
        self._h = C.c_void_p()
        _check(lib().vnr_volume_create(model_json_text.encode(), int(dims[0]), int(dims[1]), int(dims[2]), C.byref(self._h)))
        self.dims = tuple(int(d) for d in dims)
        self._read_model_info()

    def _read_model_info(self):
        n, nm = C.c_uint64(), C.c_uint64()
        L, F, H = C.c_int(), C.c_int(), C.c_int()
        _check(lib().vnr_volume_model_info(self._h, C.byref(n), C.byref(nm), C.byref(L), C.byref(F), C.byref(H)))
        self.n_params, self.n_mlp_params = n.value, nm.value
        self.n_levels, self.n_features, self.n_hidden = L.value, F.value, H.value
        self.enc_pad = ((self.n_levels * self.n_features + 15) // 16) * 16

    def set_model(self, model_json_text, seed=1337):
        """vnrNeuralVolumeSetModel (api.h:126): a new network under the same dims / ground truth / sampler / macrocells."""
        _check(lib().vnr_volume_set_model(self._h, model_json_text.encode(), C.c_uint32(seed)))
        self._read_model_info()

    def close(self):
        if self._h:
            lib().vnr_volume_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters
    def init_params(self, seed=1337):
        _check(lib().vnr_volume_init_params(self._h, C.c_uint32(seed)))

    def set_params_f16(self, params_u16):
        p = np.ascontiguousarray(params_u16, dtype=np.uint16)
        _check(lib().vnr_volume_set_params_f16(self._h, _ptr(p), C.c_size_t(p.size)))

    def get_params_f16(self):
        p = np.empty(self.n_params, dtype=np.uint16)
        _check(lib().vnr_volume_get_params_f16(self._h, _ptr(p), C.c_size_t(p.size)))
        return p

    def load_params(self, blob):
        _check(lib().vnr_volume_load_params(self._h, blob, C.c_size_t(len(blob))))

    def save_params(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().vnr_volume_save_params(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    # -- decode (NeuralVolume::inference)
    def decode(self, d_xyz, d_out, n, stream=None):
        _check(lib().vnr_volume_decode(self._h, _ptr(d_xyz), _ptr(d_out), C.c_size_t(n), _stream(stream)))

    def gather_probe(self, d_xyz, d_out, n, stream=None):
        _check(lib().vnr_volume_gather_probe(self._h, _ptr(d_xyz), _ptr(d_out), C.c_size_t(n), _stream(stream)))

    def decode_host(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.empty(xyz.shape[0], dtype=np.float32)
        _check(lib().vnr_volume_decode_host(self._h, _ptr(xyz), _ptr(out), C.c_size_t(xyz.shape[0])))
        return out

    def decode_debug(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.empty(xyz.shape[0], dtype=np.float32)
        enc = np.empty((xyz.shape[0], self.enc_pad), dtype=np.uint16)
        _check(lib().vnr_volume_decode_debug(self._h, _ptr(xyz), _ptr(out), _ptr(enc), C.c_size_t(xyz.shape[0])))
        return out, enc

    # -- ground truth, macrocell, transfer function
    def set_groundtruth(self, volume):
        v = _f32(volume)
        if v.size != self.dims[0] * self.dims[1] * self.dims[2]:
            raise VnrError(-1, "ground-truth volume size does not match dims")
        _check(lib().vnr_volume_set_groundtruth_f32(self._h, _ptr(v)))

    VALUE_TYPES = {"uint8": 0, "int8": 1, "uint16": 2, "int16": 3, "uint32": 4, "int32": 5, "float32": 8, "float64": 12}

    def set_groundtruth_file(self, path, dtype, offset=0, big_endian=False, value_range=None):
        """raw structured volume file -> HBM, normalised to [0,1]; returns the unnormalised (min, max) used"""
        lo, hi = value_range if value_range else (1.0, 0.0)
        r = (C.c_float * 2)()
        _check(lib().vnr_volume_set_groundtruth_file(self._h, str(path).encode(), C.c_int(self.VALUE_TYPES[str(np.dtype(dtype))]), C.c_uint64(offset),
                                                     C.c_int(1 if big_endian else 0), C.c_float(lo), C.c_float(hi), r))
        return float(r[0]), float(r[1])

    def set_groundtruth_outofcore(self, path, dtype, value_range, offset=0, num_concurrent_blocks=0, num_blocks=0):
        """OutOfCoreSampler: the volume stays in the file; training draws from a pool of random slabs kept in HBM"""
        _check(lib().vnr_volume_set_groundtruth_outofcore(self._h, str(path).encode(), C.c_int(self.VALUE_TYPES[str(np.dtype(dtype))]), C.c_uint64(offset),
                                                          C.c_float(value_range[0]), C.c_float(value_range[1]), C.c_uint32(num_concurrent_blocks),
                                                          C.c_uint32(num_blocks)))
        self.uniforms_per_sample = 5

    def outofcore_info(self, table=False):
        ns, nr, sb, up = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(lib().vnr_volume_outofcore_info(self._h, C.byref(ns), C.byref(nr), C.byref(sb), None, None, C.byref(up)))
        info = {"n_slots": ns.value, "n_refresh": nr.value, "slot_bytes": sb.value, "bytes_uploaded": up.value}
        if table:
            first = np.empty(ns.value, dtype=np.uint64); length = np.empty(ns.value, dtype=np.uint32)
            _check(lib().vnr_volume_outofcore_info(self._h, None, None, None, _ptr(first), _ptr(length), C.byref(up)))
            info.update(first_voxel=first, length=length, bytes_uploaded=up.value)
        return info

    def set_groundtruth_device(self, d_volume):
        _check(lib().vnr_volume_set_groundtruth_device(self._h, _ptr(d_volume)))

    def macrocell_from_groundtruth(self):
        _check(lib().vnr_volume_macrocell_from_groundtruth(self._h))

    def get_macrocell(self):
        md = np.zeros(3, dtype=np.int32)
        _check(lib().vnr_volume_get_macrocell(self._h, _ptr(md), None, None))
        cells = int(md[0]) * int(md[1]) * int(md[2])
        vr = np.empty(2 * cells, dtype=np.float32)
        mo = np.empty(cells, dtype=np.float32)
        _check(lib().vnr_volume_get_macrocell(self._h, _ptr(md), _ptr(vr), _ptr(mo)))
        return tuple(int(x) for x in md), vr, mo

    def set_macrocell(self, value_range):
        vr = _f32(value_range)
        _check(lib().vnr_volume_set_macrocell(self._h, _ptr(vr)))

    def set_transfer_function(self, rgb, alpha, value_range=(0.0, 1.0)):
        rgb = _f32(rgb).reshape(-1, 3)
        alpha = _f32(alpha)
        _check(lib().vnr_volume_set_tfn(self._h, _ptr(rgb), C.c_int(rgb.shape[0]), _ptr(alpha), C.c_int(alpha.size),
                                        C.c_float(value_range[0]), C.c_float(value_range[1])))

    # -- training
    def train(self, steps, batch=0, fast_mode=True, stream=None):
        _check(lib().vnr_volume_train(self._h, C.c_int(steps), C.c_int(batch), C.c_int(1 if fast_mode else 0), _stream(stream)))

    def train_on(self, d_xyz, d_target, n, stream=None):
        _check(lib().vnr_volume_train_on(self._h, _ptr(d_xyz), _ptr(d_target), C.c_size_t(n), _stream(stream)))

    def train_grads(self, d_xyz, d_target, n, n_global=None, stream=None):
        _check(lib().vnr_volume_train_grads(self._h, _ptr(d_xyz), _ptr(d_target), C.c_size_t(n), C.c_size_t(n_global or n),
                                            _stream(stream)))

    def optimizer_step(self, stream=None):
        _check(lib().vnr_volume_optimizer_step(self._h, _stream(stream)))

    # -- communicator (data-parallel training through train())
    def attach_comm(self, comm):
        _check(lib().vnr_volume_attach_comm(self._h, comm._h))

    def detach_comm(self):
        _check(lib().vnr_volume_detach_comm(self._h))

    # -- measurement taps of the fused training kernel (train.cu)
    def train_debug(self, variant=1, flags=0, profile=False):
        _check(lib().vnr_volume_train_debug(self._h, C.c_int(variant), C.c_uint32(flags), C.c_int(1 if profile else 0)))

    def train_profile(self):
        """per-CTA role timers of the last training kernel: uint32 array [n_ctas][words]"""
        n, w = C.c_int(), C.c_int()
        _check(lib().vnr_volume_train_profile(self._h, None, C.c_size_t(0), C.byref(n), C.byref(w)))
        out = np.zeros((max(n.value, 0), w.value), dtype=np.uint32)
        if out.size:
            _check(lib().vnr_volume_train_profile(self._h, _ptr(out), C.c_size_t(out.size), None, None))
        return out

    def grad_buffer(self, which):
        """(device pointer, n_elements, is_f32) of the MLP (which=0) or grid (which=1) gradients."""
        p, n, f = C.c_void_p(), C.c_size_t(), C.c_int()
        _check(lib().vnr_volume_grad_buffer(self._h, C.c_int(which), C.byref(p), C.byref(n), C.byref(f)))
        return p.value, n.value, bool(f.value)

    def get_grads(self):
        gm = np.empty(self.n_mlp_params, dtype=np.float32)
        gg = np.empty(self.n_params - self.n_mlp_params, dtype=np.uint16)
        _check(lib().vnr_volume_get_grads(self._h, _ptr(gm), _ptr(gg)))
        return gm, gg

    def sample_at(self, xyz, hw_texture=False):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.empty(xyz.shape[0], dtype=np.float32)
        _check(lib().vnr_volume_sample_at(self._h, _ptr(xyz), _ptr(out), C.c_size_t(xyz.shape[0]), C.c_int(1 if hw_texture else 0)))
        return out

    def last_loss(self):
        loss = C.c_double()
        _check(lib().vnr_volume_last_loss(self._h, C.byref(loss)))
        return loss.value

    def sample(self, d_xyz, d_target, n, stream=None):
        _check(lib().vnr_volume_sample(self._h, _ptr(d_xyz), _ptr(d_target), C.c_size_t(n), _stream(stream)))

    def stream(self):
        p = C.c_void_p()
        _check(lib().vnr_volume_stream(self._h, C.byref(p)))
        return p.value or 0

    def macrocell_update(self, d_xyz, d_values, n, stream=None):
        _check(lib().vnr_volume_macrocell_update(self._h, _ptr(d_xyz), _ptr(d_values), C.c_size_t(n), _stream(stream)))

    def macrocell_buffer(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().vnr_volume_macrocell_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def macrocell_refresh(self, stream=None):
        _check(lib().vnr_volume_macrocell_refresh(self._h, _stream(stream)))

    def dp_export(self):
        h = C.create_string_buffer(192)
        _check(lib().vnr_volume_dp_export(self._h, h))
        return h.raw

    def dp_attach(self, rank, world, all_handles):
        _check(lib().vnr_volume_dp_attach(self._h, int(rank), int(world), C.c_char_p(all_handles) if all_handles else None))

    def dp_detach(self):
        _check(lib().vnr_volume_dp_detach(self._h))

    def dp_optimizer_step(self, stream=None):
        _check(lib().vnr_volume_dp_optimizer_step(self._h, _stream(stream)))

    def dp_finish_step(self, stream=None):
        _check(lib().vnr_volume_dp_finish_step(self._h, _stream(stream)))

    def sampler_skip(self, n_floats):
        _check(lib().vnr_volume_sampler_skip(self._h, C.c_uint64(n_floats)))

    def decode_progressive(self, stream=None):
        _check(lib().vnr_volume_decode_progressive(self._h, _stream(stream)))

    def num_blobs(self):
        n = C.c_int()
        _check(lib().vnr_volume_num_blobs(self._h, C.byref(n)))
        return n.value

    def get_decoded(self):
        out = np.empty(self.dims[2] * self.dims[1] * self.dims[0], dtype=np.float32)
        _check(lib().vnr_volume_get_decoded(self._h, _ptr(out)))
        return out.reshape(self.dims[2], self.dims[1], self.dims[0])

    def psnr(self):
        v = C.c_double()
        _check(lib().vnr_volume_psnr(self._h, C.byref(v)))
        return v.value

    def ssim(self, return_map=False):
        """vnrNeuralVolumeGetSSIM; return_map: also the per-window values [(dz-6), (dy-6), (dx-6)]."""
        v = C.c_double()
        m = None
        if return_map:
            m = np.empty((self.dims[2] - 6, self.dims[1] - 6, self.dims[0] - 6), dtype=np.float32)
        _check(lib().vnr_volume_ssim(self._h, C.byref(v), _ptr(m) if m is not None else None))
        return (v.value, m) if return_map else v.value

    def test_loss(self, batch=0):
        v = C.c_double()
        _check(lib().vnr_volume_test_loss(self._h, C.c_int(batch), C.byref(v)))
        return v.value

    def export(self, path, which=0):
        """vnrNeuralVolumeDecodeInference (which=0) / DecodeReference (which=1); returns (min, max) written."""
        r = (C.c_float * 2)()
        _check(lib().vnr_volume_export(self._h, os.fsencode(path), C.c_int(which), r))
        return float(r[0]), float(r[1])

    def stats(self):
        step, loss = C.c_uint64(), C.c_double()
        _check(lib().vnr_volume_stats(self._h, C.byref(step), C.byref(loss)))
        return step.value, loss.value


class Renderer:
    """vnrRenderer -- api.h:168-178."""

    def __init__(self, volume):
        self._h = C.c_void_p()
        self.volume = volume           # keeps the volume alive (RendererContext, api_internal.h:41-45)
        _check(lib().vnr_renderer_create(volume._h, C.byref(self._h)))
        self.size = (0, 0)

    def close(self):
        if self._h:
            lib().vnr_renderer_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_size(self, w, h):
        _check(lib().vnr_renderer_set_size(self._h, int(w), int(h)))
        self.size = (int(w), int(h))

    def set_camera(self, cam_from, cam_at, cam_up, fovy=60.0):
        _check(lib().vnr_renderer_set_camera(self._h, _ptr(_f32(cam_from)), _ptr(_f32(cam_at)), _ptr(_f32(cam_up)), C.c_float(fovy)))

    def set_mode(self, mode):
        _check(lib().vnr_renderer_set_mode(self._h, int(mode)))

    def set_groundtruth_source(self, on):
        _check(lib().vnr_renderer_set_groundtruth_source(self._h, C.c_int(1 if on else 0)))

    def set_sampling_rate(self, r):
        _check(lib().vnr_renderer_set_sampling_rate(self._h, C.c_float(r)))

    def set_density_scale(self, s):
        _check(lib().vnr_renderer_set_density_scale(self._h, C.c_float(s)))

    def set_clipping_box(self, lower, upper):
        _check(lib().vnr_renderer_set_clipping_box(self._h, _ptr(_f32(lower)), _ptr(_f32(upper))))

    def set_layout(self, tiled=True, transpose=True):
        _check(lib().vnr_renderer_set_layout(self._h, C.c_int(1 if tiled else 0), C.c_int(1 if transpose else 0)))

    def set_scaling(self, scale):
        _check(lib().vnr_renderer_set_scaling(self._h, _ptr(_f32(scale))))

    def reset_accumulation(self):
        _check(lib().vnr_renderer_reset_accumulation(self._h))

    def set_partition(self, rank, world):
        _check(lib().vnr_renderer_set_partition(self._h, int(rank), int(world)))

    def set_jitter_mode(self, mode):
        _check(lib().vnr_renderer_set_jitter_mode(self._h, int(mode)))

    def render(self):
        _check(lib().vnr_render(self._h))

    def map_frame(self, copy=True):
        """vnrRendererMapFrame: syncs and returns the host frame (h, w, 4).  copy=False returns a view of the
        library's pinned buffer, valid until the second-next map_frame (the reference's double-buffer contract)."""
        p = lib().vnr_map_frame(self._h)
        if not p:
            raise VnrError(-4, lib().vnr_last_error().decode("utf-8", "replace"))
        w, h = self.size
        a = np.ctypeslib.as_array(p, shape=(h, w, 4))
        return a.copy() if copy else a

    def device_frame(self):
        p = C.c_void_p()
        _check(lib().vnr_renderer_device_frame(self._h, C.byref(p), None))
        return p.value

    def set_download(self, on):
        _check(lib().vnr_renderer_set_download(self._h, C.c_int(1 if on else 0)))

    def round_counts(self, max_rounds=64):
        out = (C.c_uint32 * max_rounds)(); n = C.c_int()
        _check(lib().vnr_renderer_round_counts(self._h, out, C.c_int(max_rounds), C.byref(n)))
        return [int(out[k]) for k in range(n.value)]

    def set_zero_copy(self, on):
        _check(lib().vnr_renderer_set_zero_copy(self._h, C.c_int(1 if on else 0)))

    def download(self):
        _check(lib().vnr_renderer_download(self._h))

    def set_profiling(self, on):
        _check(lib().vnr_renderer_set_profiling(self._h, C.c_int(1 if on else 0)))

    def set_graph(self, on):
        _check(lib().vnr_renderer_set_graph(self._h, C.c_int(1 if on else 0)))

    def set_frame_target(self, d_ptr):
        _check(lib().vnr_renderer_set_frame_target(self._h, C.c_void_p(d_ptr) if d_ptr else None))

    def attach_comm(self, comm):
        """tile-parallel rendering over the communicator: render() on every rank, map_frame() on rank 0"""
        _check(lib().vnr_renderer_attach_comm(self._h, comm._h))

    def detach_comm(self):
        _check(lib().vnr_renderer_detach_comm(self._h))

    def set_frames_in_flight(self, n):
        """depth of the renderer's frame ring (vnr_renderer_set_frames_in_flight); map_frame then returns the oldest unmapped frame"""
        _check(lib().vnr_renderer_set_frames_in_flight(self._h, C.c_int(n)))

    def streams(self):
        """the cudaStream_t of every frame slot"""
        n = C.c_int()
        _check(lib().vnr_renderer_streams(self._h, None, C.c_int(0), C.byref(n)))
        arr = (C.c_void_p * n.value)()
        _check(lib().vnr_renderer_streams(self._h, arr, C.c_int(n.value), None))
        return [int(a or 0) for a in arr]

    def set_n_iters(self, n):
        _check(lib().vnr_renderer_set_n_iters(self._h, C.c_int(n)))

    def profile(self):
        ms, n, k = C.c_float(), C.c_int(), C.c_uint64()
        _check(lib().vnr_renderer_profile(self._h, C.byref(ms), C.byref(n), C.byref(k)))
        return {"decode_ms": ms.value, "decode_launches": n.value, "kernel_launches": k.value}

    def stream(self):
        p = C.c_void_p()
        _check(lib().vnr_renderer_stream(self._h, C.byref(p)))
        return p.value or 0

    def stats(self):
        s = np.zeros(4, dtype=np.uint64)
        _check(lib().vnr_renderer_stats(self._h, _ptr(s)))
        return {"rays_hit": int(s[0]), "samples_decoded": int(s[1]), "samples_composited": int(s[2]), "rounds": int(s[3])}


class Scene:
    """Scene description (VIDI3D / DIVA JSON) as the reference's apps pass it to vnrCreateSimpleVolume /
    vnrCreateCamera / vnrCreateTransferFunction (serializer.cpp:138-477).  Host-only."""

    VALUE_TYPE_NAMES = {0: "uint8", 1: "int8", 2: "uint16", 3: "int16", 4: "uint32", 5: "int32", 8: "float32", 12: "float64"}

    def __init__(self, text=None, path=None):
        self._h = C.c_void_p()
        arg = os.fsencode(path) if path is not None else text.encode("utf-8")
        _check(lib().vnr_scene_create(arg, C.c_int(1 if path is not None else 0), C.byref(self._h)))
        dims, vt, nt, rg, hr = (C.c_int * 3)(), C.c_int(), C.c_int(), (C.c_float * 2)(), C.c_int()
        _check(lib().vnr_scene_volume(self._h, dims, C.byref(vt), C.byref(nt), rg, C.byref(hr)))
        self.dims, self.value_type, self.n_timesteps = tuple(dims), vt.value, nt.value
        self.dtype = self.VALUE_TYPE_NAMES.get(vt.value)
        self.value_range = (float(rg[0]), float(rg[1])) if hr.value else None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vnr_scene_release.restype = None
            lib().vnr_scene_release(self._h)
            self._h = None

    def timestep(self, t=0):
        name, off, big = C.c_char_p(), C.c_uint64(), C.c_int()
        _check(lib().vnr_scene_timestep(self._h, C.c_int(t), C.byref(name), C.byref(off), C.byref(big)))
        return os.fsdecode(name.value), off.value, bool(big.value)

    def camera(self):
        f, a, u, fov = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)(), C.c_float()
        _check(lib().vnr_scene_camera(self._h, f, a, u, C.byref(fov)))
        return tuple(f), tuple(a), tuple(u), fov.value

    def tfn(self):
        """(rgb [n,3] or None, alpha pairs [m,2] or None, value range or None); raises VnrError(-3...) when the scene
        carries a transfer function in the (unrestated) OVR tfn-module format."""
        rgb, a = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        n, m, rg, hr = C.c_int(), C.c_int(), (C.c_float * 2)(), C.c_int()
        _check(lib().vnr_scene_tfn(self._h, C.byref(rgb), C.byref(n), C.byref(a), C.byref(m), rg, C.byref(hr)))
        col = np.ctypeslib.as_array(rgb, (n.value, 3)).copy() if n.value else None
        alp = np.ctypeslib.as_array(a, (m.value, 2)).copy() if m.value else None
        return col, alp, ((float(rg[0]), float(rg[1])) if hr.value else None)

    def load_groundtruth(self, volume, t=0):
        """vnrCreateSimpleVolume(scene, "GPU") + vnrCreateNeuralVolume(config, simple): time step t -> HBM, normalised
        with the scene's range (or the data's when it has none); returns the unnormalised range used."""
        name, off, big = self.timestep(t)
        return volume.set_groundtruth_file(name, self.dtype, offset=off, big_endian=big, value_range=self.value_range)


def ipc_export(d_ptr):
    h = C.create_string_buffer(64)
    _check(lib().vnr_ipc_export(C.c_void_p(d_ptr), h))
    return h.raw


def ipc_open(handle):
    p = C.c_void_p()
    _check(lib().vnr_ipc_open(C.c_char_p(handle), C.byref(p)))
    return p.value


def ipc_close(d_ptr):
    _check(lib().vnr_ipc_close(C.c_void_p(d_ptr)))


class PeerBarrier:
    """vnr_peer_barrier_*: stream-ordered cross-rank barrier over NVLink peer memory."""

    def __init__(self):
        self._h = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(lib().vnr_peer_barrier_create(C.byref(self._h), h))
        self.handle = h.raw

    def attach(self, rank, world, all_handles):
        _check(lib().vnr_peer_barrier_attach(self._h, int(rank), int(world), C.c_char_p(all_handles) if all_handles else None))

    def sync(self, stream):
        _check(lib().vnr_peer_barrier_sync(self._h, _stream(stream)))

    def timed_out(self):
        v = C.c_uint64()
        _check(lib().vnr_peer_barrier_check(self._h, C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            lib().vnr_peer_barrier_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


VNR_RAYMARCHING_NO_SHADING_DECODING = 4
VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING = 5
VNR_RAYMARCHING_NO_SHADING_IN_SHADER = 6
VNR_RAYMARCHING_GRADIENT_SHADING_DECODING = 7
VNR_RAYMARCHING_GRADIENT_SHADING_SAMPLE_STREAMING = 8
VNR_RAYMARCHING_GRADIENT_SHADING_IN_SHADER = 9
VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_DECODING = 10
VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_SAMPLE_STREAMING = 11
VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_IN_SHADER = 12
VNR_PATHTRACING_DECODING = 13
VNR_PATHTRACING_SAMPLE_STREAMING = 14
VNR_PATHTRACING_IN_SHADER = 15
