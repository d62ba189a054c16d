"""Scene descriptions (scene.cpp) against the reference's serializer (serializer.cpp:138-477): the VIDI3D and DIVA
layouts, value-type names, ranges and the camera shift.  Host-only: runs without a GPU."""
import json

import numpy as np
import pytest

import instantvnr_b200 as vnr


def _vidi(tmp_path, type_name="UNSIGNED_SHORT", extra_volume=None, files=1, **data):
    names = []
    for i in range(files):
        p = tmp_path / f"t{i}.raw"
        p.write_bytes(b"\0" * 16)
        names.append(str(p))
    ds = [dict({"format": "REGULAR_GRID_RAW_BINARY", "fileName": n, "dimensions": {"x": 40, "y": 30, "z": 20}, "type": type_name}, **data) for n in names]
    vol = {"transferFunction": {"whatever": 1}}
    vol.update(extra_volume or {})
    return {"version": "VIDI3D", "dataSource": ds,
            "view": {"volume": vol, "camera": {"eye": {"x": 20.0, "y": 15.0, "z": -90.0}, "center": {"x": 20.0, "y": 15.0, "z": 10.0},
                                                 "up": {"x": 0.0, "y": 1.0, "z": 0.0}, "fovy": 45.0}}}


def test_vidi_scene_volume_camera_and_ranges(tmp_path):
    root = _vidi(tmp_path, extra_volume={"scalarMappingRange": {"minimum": 0.25, "maximum": 0.5}}, offset=128, endian="BIG_ENDIAN", files=3)
    sc = vnr.Scene(text=json.dumps(root))
    assert sc.dims == (40, 30, 20) and sc.dtype == "uint16" and sc.n_timesteps == 3
    assert sc.timestep(2) == (str(tmp_path / "t2.raw"), 128, True)
    # scalarMappingRange x numeric_limits<uint16_t>::max() in float (serializer.cpp:229-232)
    assert sc.value_range == (float(np.float32(65535.0 * np.float32(0.25))), float(np.float32(65535.0 * np.float32(0.5))))
    frm, at, up, fovy = sc.camera()
    assert frm == (0.0, 0.0, -100.0) and at == (0.0, 0.0, 0.0) and up == (0.0, 1.0, 0.0) and fovy == 45.0   # eye/center - dims/2 (:367-369)
    # the unnormalised range wins (:213-217); the version key is optional (:428)
    root["view"]["volume"]["scalarMappingRangeUnnormalized"] = {"minimum": -3.5, "maximum": 900.0}
    del root["version"]
    assert vnr.Scene(text=json.dumps(root)).value_range == (-3.5, 900.0)
    # no range keys at all: taken from the data
    root["view"]["volume"] = {}
    sc2 = vnr.Scene(text=json.dumps(root))
    assert sc2.value_range is None and sc2.tfn() == (None, None, None)
    with pytest.raises(vnr.VnrError):
        sc2.timestep(3)


@pytest.mark.parametrize("name,dtype", [("BYTE", "int8"), ("UNSIGNED_BYTE", "uint8"), ("SHORT", "int16"), ("UNSIGNED_SHORT", "uint16"),
                                          ("INT", "int32"), ("UNSIGNED_INT", "uint32"), ("FLOAT", "float32"), ("DOUBLE", "float64"),
                                          ("HALF", "int8")])       # unknown names map to the first enum pair (NLOHMANN_JSON_SERIALIZE_ENUM)
def test_value_type_names(tmp_path, name, dtype):
    assert vnr.Scene(text=json.dumps(_vidi(tmp_path, type_name=name))).dtype == dtype


def test_float_range_is_not_scaled_and_the_tfn_table_is_explicit_only(tmp_path):
    root = _vidi(tmp_path, type_name="FLOAT", extra_volume={"scalarMappingRange": {"minimum": 0.1, "maximum": 0.9}})
    sc = vnr.Scene(text=json.dumps(root))
    assert sc.value_range == (float(np.float32(0.1)), float(np.float32(0.9)))
    with pytest.raises(vnr.VnrError) as e:          # OVR tfn-module format: not restated, reported
        sc.tfn()
    assert e.value.code == -3
    root["view"]["volume"]["transferFunction"] = {"colors": [[0, 0, 1], [1, 0, 0]], "alphas": [[0.0, 0.0], [0.5, 0.25], [1.0, 1.0]]}
    col, alp, rg = vnr.Scene(text=json.dumps(root)).tfn()
    assert col.tolist() == [[0, 0, 1], [1, 0, 0]] and alp.tolist() == [[0.0, 0.0], [0.5, 0.25], [1.0, 1.0]] and rg == sc.value_range


def test_candidate_file_names_and_errors(tmp_path):
    root = _vidi(tmp_path)
    real = root["dataSource"][0]["fileName"]
    root["dataSource"][0]["fileName"] = [str(tmp_path / "missing.raw"), real]       # valid_filename: first existing candidate
    assert vnr.Scene(text=json.dumps(root)).timestep(0)[0] == real
    root["dataSource"][0]["fileName"] = [str(tmp_path / "missing.raw")]
    with pytest.raises(vnr.VnrError, match="Cannot find volume file"):
        vnr.Scene(text=json.dumps(root))
    bad = _vidi(tmp_path); bad["dataSource"][0]["format"] = "VDB"
    with pytest.raises(vnr.VnrError, match="data type unimplemented"):
        vnr.Scene(text=json.dumps(bad))
    with pytest.raises(vnr.VnrError, match="unknown JSON configuration format"):
        vnr.Scene(text=json.dumps(dict(_vidi(tmp_path), version="X")))
    with pytest.raises(vnr.VnrError, match="expected to be an array"):
        vnr.Scene(text=json.dumps(dict(_vidi(tmp_path), dataSource={})))
    with pytest.raises(vnr.VnrError):
        vnr.Scene(text="{ not json")


def test_diva_scene_and_scene_files_with_comments(tmp_path):
    root = {"version": "DIVA", "volume": {"dims": {"x": 8, "y": 9, "z": 10}, "type": "FLOAT", "range": {"x": -1.0, "y": 2.0},
                                          "filename": ["a.raw", "b.raw"], "bigendian": True}}
    sc = vnr.Scene(text=json.dumps(root))
    assert sc.dims == (8, 9, 10) and sc.dtype == "float32" and sc.value_range == (-1.0, 2.0) and sc.n_timesteps == 2
    assert sc.timestep(1) == ("b.raw", 0, True)
    with pytest.raises(vnr.VnrError):               # camera / tfn of DIVA scenes are "TODO" in the reference (serializer.cpp:174)
        sc.camera()
    # a scene FILE (vnrJson that is_string()): parsed with comments allowed (json::parse(file, nullptr, true, true), serializer.h:20);
    # relative file names fall back to the scene's directory
    (tmp_path / "vol.raw").write_bytes(b"\0" * 4)
    text = "// scene\n" + json.dumps(dict(root, volume=dict(root["volume"], filename="vol.raw"))) + "\n/* end */\n"
    path = tmp_path / "scene.json"
    path.write_text(text)
    assert vnr.Scene(path=path).timestep(0)[0] == str(tmp_path / "vol.raw")
    with pytest.raises(vnr.VnrError):
        vnr.Scene(path=tmp_path / "nope.json")
