"""Row N3 (host side, CPU): the DGDE -> GMW wire format.  The compatibility reader must reproduce the unmodified
reference reader (`GMW/utilities/dataset_utilities.py:11-56`; fixture `wire_reference_load_data.npz` generated from it by
oracle/make_golden.py on the two committed JSON files); the binary form must round-trip bit-exactly, map without parsing
and be several times smaller than the indent=4 JSON."""
import os

import numpy as np
import pytest
import torch

from dcd_b200 import synth, wire
from conftest import GOLDEN

TRAIN_JSON = os.path.join(GOLDEN, "gen_data_train_small.json")
INFER_JSON = os.path.join(GOLDEN, "gen_data_infer_small.json")


@pytest.mark.parametrize("split,path", [("train", TRAIN_JSON), ("valid", INFER_JSON)])
def test_json_reader_equals_reference_load_data(split, path):
    ref = np.load(os.path.join(GOLDEN, "wire_reference_load_data.npz"))
    mine = wire.load_reference_json(path, split)
    for key in ("kpts_2d", "kpts_3d", "pred_rot", "gt_location", "img_idx", "dim"):
        want = ref["%s_%s" % (split, key)]
        assert mine[key].dtype == np.float32 and mine[key].shape == want.shape and np.array_equal(mine[key], want), key
    if split == "valid":
        assert mine["image_ids"] == ["000007", "000123"]                       # the float cast of the reference loses "000007"
        assert mine["img_ref"].tolist() == [[0, 0], [0, 1], [1, 0], [1, 1], [1, 2]]
        assert mine["kpts_2d"].shape[1] == 73                                   # 83 regressed, the first 73 kept


@pytest.mark.parametrize("split,path", [("train", TRAIN_JSON), ("valid", INFER_JSON)])
def test_binary_round_trip_and_size(tmp_path, split, path):
    out = str(tmp_path / ("gen_%s.dcdw" % split))
    sizes = wire.convert_json(path, out, split)
    assert sizes["binary_bytes"] == os.path.getsize(out) and sizes["json_bytes"] > 5 * sizes["binary_bytes"]
    ref = wire.load_reference_json(path, split)
    for mmap in (True, False):
        got = wire.load(out, mmap=mmap)
        assert set(got) == set(ref)
        for k, v in ref.items():
            if k == "image_ids":
                assert got[k] == v
            else:
                assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k
                assert not mmap or got[k].ctypes.data % 64 == 0 or got[k].size == 0   # mapped payloads are 64-byte aligned
    ds_json, ds_bin = wire.Dataset(split, path), wire.Dataset(split, out)
    assert len(ds_json) == len(ds_bin) == 5
    for i in range(len(ds_bin)):
        for a, b in zip(ds_json[i], ds_bin[i]):
            assert np.array_equal(a, b)
    assert len(ds_bin[0]) == (6 if split == "valid" else 5)                    # (…, dim, img_idx) only for validation


def test_binary_rejects_foreign_and_truncated_files(tmp_path):
    p = tmp_path / "x.dcdw"
    p.write_bytes(b"not a wire file at all")
    with pytest.raises(ValueError):
        wire.load(str(p))
    out = str(tmp_path / "t.dcdw")
    wire.convert_json(TRAIN_JSON, out, "train")
    blob = open(out, "rb").read()
    open(out, "wb").write(blob[:len(blob) // 2])
    with pytest.raises(ValueError):
        wire.load(out)
    with pytest.raises(ValueError):
        wire.load_reference_json(TRAIN_JSON, "test")


def test_from_detector_matches_the_json_route(tmp_path):
    """Records built straight from detector tensors equal the ones that went through the JSON files."""
    ob = synth.make_objects(N=6, n=83, seed=5)
    P = np.array(synth.P2, dtype=np.float64)
    loc = torch.stack((0.3 * ob.gt_depth, torch.full((6,), 1.6), ob.gt_depth), dim=1)
    dim = torch.tensor([[1.5, 1.6, 3.9]]).expand(6, 3)
    ids = ["000031", "000032"]
    ref = np.array([[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2]], dtype=np.int32)
    rec = wire.from_detector(ob.kps, ob.kps_3d, ob.rot_y, loc, P, dim=dim, image_ids=ids, img_ref=ref)
    assert rec["kpts_2d"].shape == (6, 73, 2) and rec["kpts_3d"].shape == (6, 73, 3)
    assert np.allclose(rec["kpts_2d"], ob.kps_norm[:, :73].numpy(), rtol=0, atol=1e-6)   # detector_loss.py:149-155
    assert rec["img_idx"].tolist()[3] == [32.0, 0.0]
    out = str(tmp_path / "d.dcdw")
    wire.save(out, rec)
    back = wire.load(out)
    assert back["image_ids"] == ids and np.array_equal(back["kpts_3d"], rec["kpts_3d"])
    ds = wire.Dataset("valid", out)
    k2, k3, rot, gl, dm, ii = ds[4]
    assert np.array_equal(k3, ob.kps_3d[4, :73].numpy()) and ii.tolist() == [32.0, 1.0] and dm.tolist() == [1.5, 1.600000023841858, 3.9000000953674316]
