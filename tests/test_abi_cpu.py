"""The C-ABI library loads without a GPU and exports every symbol include/vnr_c.h declares;
host-side logic (config parsing, error codes) works; compute calls fail loudly without a device."""
import ctypes as C
import os
import re
import subprocess

import pytest

import instantvnr_b200 as vnr


def _declared_symbols():
    txt = open(vnr.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vnr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = vnr.lib()
    syms = _declared_symbols()
    assert len(syms) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", vnr.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    for s in syms:
        getattr(lib, s)


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", vnr.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_reference_the_oracle():
    src = os.path.join(os.path.dirname(vnr.__file__))
    for root, _, files in os.walk(src):
        if "_build" in root:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inl", ".py", "Makefile")):
                body = open(os.path.join(root, f)).read()
                code = "\n".join(l for l in body.splitlines() if "oracle" in l and not l.strip().startswith(("//", "#", '"', "*")))
                assert "import oracle" not in code and "oracle/" not in code.replace("oracle/.", ""), (f, code)
    ldd = subprocess.check_output(["ldd", vnr.LIB_PATH]).decode()
    assert "oracle" not in ldd


def test_config_errors_are_reported_not_thrown():
    h = C.c_void_p()
    lib = vnr.lib()
    assert lib.vnr_volume_create(b"{ not json", 8, 8, 8, C.byref(h)) == -1
    assert b"json" in lib.vnr_last_error()
    assert lib.vnr_volume_create(b'{"encoding": {"otype": "Frequency"}}', 8, 8, 8, C.byref(h)) == -3
    assert lib.vnr_volume_create(vnr.example_model_json().encode(), 0, 8, 8, C.byref(h)) == -1
    assert lib.vnr_volume_create(b'{"encoding": {"n_features_per_level": 3}}', 8, 8, 8, C.byref(h)) == -1
    assert lib.vnr_volume_create(None, 8, 8, 8, C.byref(h)) == -1


def test_compute_fails_loudly_without_device():
    if vnr.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vnr.VnrError) as e:
        vnr.NeuralVolume(vnr.example_model_json(), (8, 8, 8))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_level_table_of_the_probe_matches_the_oracle():
    """bench.py's level-structured gather probe (vnr_probe_levels) is sized by instantvnr_b200.hash_grid_level_entries: the same
    per-level entry counts as the oracle's restatement of tcnn's grid layout (encodings/grid.h:591-611)"""
    import numpy as np
    import instantvnr_b200 as vnr
    import oracle as O
    for kw in (dict(), dict(n_levels=16, log2_hashmap=19, base_res=16), dict(n_levels=8, log2_hashmap=22, base_res=16), dict(n_levels=4, log2_hashmap=12, base_res=8)):
        n_levels, log2, base = kw.get("n_levels", 8), kw.get("log2_hashmap", 19), kw.get("base_res", 16)
        m = O.ModelCfg(n_levels, 8 if n_levels <= 8 else 2, log2, base, 2.0, 4)
        want = np.diff(np.asarray(m.offsets[:n_levels + 1], dtype=np.int64))
        got = np.asarray(vnr.hash_grid_level_entries(n_levels=n_levels, log2_hashmap=log2, base_res=base), dtype=np.int64)
        assert np.array_equal(got, want), (kw, got, want)
