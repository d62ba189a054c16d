"""ctypes front-end of oracle/_ref/libvnr_marcher_ref.so: the reference's OWN ray marcher / path tracer / macrocell / transfer-
function sources (core/renderer/method_raymarching.cu, method_pathtracing.cu, core/macrocell.cu, core/instantvnr_types.cu),
compiled unmodified in place from /root/reference by oracle/ref_marcher/Makefile against oracle/ovr_shim (stand-ins for the
un-vendored OVR headers).  TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by instantvnr_b200.  Needs a GPU."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvnr_marcher_ref.so")
_lib = None
DECODE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.refm_last_error.restype = C.c_char_p
        _lib.refm_release.restype = None
    return _lib


class RefMarcherError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise RefMarcherError(lib().refm_last_error().decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefMarcher:
    """One scene of the reference renderer: a normalised float volume in a 3-D texture, its macrocells (MacroCell, the
    reference's code), a transfer function (TransferFunctionObject, the reference's code) and the two render methods."""

    def __init__(self, dims, volume):
        v = _f32(volume)
        assert v.size == int(np.prod(dims))
        self.dims = tuple(int(d) for d in dims)
        self._h = C.c_void_p()
        d = np.array(self.dims, dtype=np.int32)
        _chk(lib().refm_create(d.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.byref(self._h)))
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            lib().refm_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_volume(self, volume):
        v = _f32(volume)
        _chk(lib().refm_update_volume(self._h, v.ctypes.data_as(C.c_void_p)))

    def set_transfer_function(self, rgb, alpha, value_range=(0.0, 1.0)):
        """rgb [n,3], alpha [m] (positions are not used by the reference: only the .y of its vec2f pairs, api.cpp:485-498)"""
        rgb, alpha = _f32(rgb).reshape(-1, 3), _f32(alpha).ravel()
        pairs = _f32(np.stack([np.linspace(0, 1, alpha.size, dtype=np.float32), alpha], 1))
        _chk(lib().refm_set_transfer_function(self._h, rgb.ctypes.data_as(C.c_void_p), C.c_int(rgb.shape[0]), pairs.ctypes.data_as(C.c_void_p),
                                              C.c_int(alpha.size), C.c_float(value_range[0]), C.c_float(value_range[1])))

    def set_macrocell_value_range(self, value_range):
        vr = _f32(value_range)
        _chk(lib().refm_set_macrocell_value_range(self._h, vr.ctypes.data_as(C.c_void_p)))

    def macrocell_reset(self):
        _chk(lib().refm_macrocell_reset(self._h))

    def macrocell_update_explicit(self, d_xyz_ptr, d_values_ptr, n):
        """device pointers: n x (x, y, z) coordinates and n values (MacroCell::update_explicit, core/macrocell.cu:42-73)"""
        _chk(lib().refm_macrocell_update_explicit(self._h, C.c_void_p(d_xyz_ptr), C.c_void_p(d_values_ptr), C.c_size_t(n)))

    def get_macrocell(self):
        d = np.zeros(3, dtype=np.int32)
        _chk(lib().refm_get_macrocell(self._h, d.ctypes.data_as(C.c_void_p), None, None))
        cells = int(np.prod(d))
        vr, mo = np.zeros((cells, 2), np.float32), np.zeros(cells, np.float32)
        _chk(lib().refm_get_macrocell(self._h, d.ctypes.data_as(C.c_void_p), vr.ctypes.data_as(C.c_void_p), mo.ctypes.data_as(C.c_void_p)))
        return tuple(int(x) for x in d), vr, mo

    def set_decoder(self, fn_address, user):
        """NeuralVolume::inference of this scene = fn(user, d_xyz, d_out, n, stream): e.g. ref_inference of the reference's
        tiny-cuda-nn build, or vnr_volume_decode of the library under test (same argument order)."""
        self._keep = (fn_address, user)
        _chk(lib().refm_set_decoder(self._h, C.c_void_p(fn_address), user))

    def set_sampling(self, sampling_rate=1.0, density_scale=1.0):
        _chk(lib().refm_set_sampling(self._h, C.c_float(sampling_rate), C.c_float(density_scale)))

    def set_clipbox(self, lower, upper):
        lo, hi = _f32(lower), _f32(upper)
        _chk(lib().refm_set_clipbox(self._h, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)))

    def reset_accumulation(self):
        _chk(lib().refm_reset_accumulation(self._h))

    def render(self, mode, size, cam_from, cam_at, cam_up, fovy=60.0, neural=False):
        w, h = size
        out = np.zeros((h, w, 4), dtype=np.float32)
        st = np.zeros(2, dtype=np.uint64)
        f, a, u = _f32(cam_from), _f32(cam_at), _f32(cam_up)
        _chk(lib().refm_render(self._h, C.c_int(mode), C.c_int(1 if neural else 0), C.c_int(w), C.c_int(h), f.ctypes.data_as(C.c_void_p),
                               a.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), C.c_float(fovy), out.ctypes.data_as(C.c_void_p),
                               st.ctypes.data_as(C.c_void_p)))
        return out, {"decode_calls": int(st[0]), "decode_coords": int(st[1])}


def function_address(cdll, name):
    return C.cast(getattr(cdll, name), C.c_void_p).value
