"""ctypes front-end of oracle/_ref/libvnr_tcnn_ref.so: the reference's OWN tiny-cuda-nn build
(compiled in place from /root/reference/tcnn by oracle/ref_driver/Makefile) behind a small driver.
TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by instantvnr_b200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvnr_tcnn_ref.so")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_n_params.restype = C.c_uint64
    return _lib


class RefError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise RefError(lib().ref_last_error().decode())


class RefNetwork:
    """NetworkWithInputEncoding<3,1> + Trainer as core/networks/tcnn_network.h:200-209 builds them."""

    def __init__(self, model_json_text, seed=1337):
        import json
        txt = "\n".join(l for l in model_json_text.splitlines() if not l.strip().startswith("//"))
        json.loads(txt)
        h = lib().ref_create(txt.encode(), C.c_uint32(seed))
        if not h:
            raise RefError(lib().ref_last_error().decode())
        self.h = C.c_void_p(h)
        self.n_params = int(lib().ref_n_params(self.h))

    def set_params_f16(self, p16):
        p = np.ascontiguousarray(p16, dtype=np.uint16)
        _chk(lib().ref_set_params_f16(self.h, p.ctypes.data_as(C.c_void_p), C.c_uint64(p.size)))

    def get_params_f16(self):
        p = np.empty(self.n_params, dtype=np.uint16)
        _chk(lib().ref_get_params_f16(self.h, p.ctypes.data_as(C.c_void_p), C.c_uint64(p.size)))
        return p

    def inference(self, d_xyz, d_out, n, stream=0):
        """torch tensors (device); n multiple of 256."""
        _chk(lib().ref_inference(self.h, C.c_void_p(d_xyz.data_ptr()), C.c_void_p(d_out.data_ptr()), C.c_uint32(n), C.c_void_p(stream)))

    def training_step(self, d_xyz, d_target, n, stream=0, want_loss=True):
        loss = C.c_float(0)
        _chk(lib().ref_training_step(self.h, C.c_void_p(d_xyz.data_ptr()), C.c_void_p(d_target.data_ptr()), C.c_uint32(n), C.c_void_p(stream),
                                     C.byref(loss) if want_loss else None))
        return loss.value

    def __del__(self):
        try:
            lib().ref_destroy(self.h)
        except Exception:
            pass
