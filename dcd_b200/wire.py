"""DGDE -> GMW wire format: a binary tensor file instead of the reference's indent=4 JSON.

The reference hands the detector's per-object keypoints to GMW through `gen_data/gen_data_train.json`
(writer `DGDE/model/head/detector_loss.py:148-173` + `DGDE/engine/trainer.py:208-215`) and `gen_data_infer.json`
(writer `DGDE/engine/inference.py:59-84`): nested Python float lists, one number per line, parsed by
`GMW/utilities/dataset_utilities.py:11-56` (`load_data`) object by object.  Here:

  * `load_reference_json(path, split)`  reads those JSON files and returns exactly what the reference's `load_data`
    returns (same keys, float32 arrays, same object order) — the compatibility path;
  * `save(path, data)` / `load(path)`   one little-endian file: magic, JSON header (array table, image ids), then the
    raw float32 / int32 arrays, 64-byte aligned; `load(..., mmap=True)` maps it read-only without parsing anything;
  * `convert_json(json_path, out_path, split)`  JSON -> binary once;
  * `Dataset` mirrors the reference's `Dataset` (`dataset_utilities.py:58-73`) on either form;
  * `from_detector(...)`  builds the same records straight from DGDE tensors (the normalisation of
    `detector_loss.py:149-155`), so a joint pipeline never touches the disk.

`img_idx` needs care: the reference stores `(0, 0)` per training object and `(image id string, object index)` per
validation object and then casts the whole list with `np.array(..., dtype=np.float32)`, i.e. image id "000123" becomes
123.0.  The binary file keeps that float pair (`img_idx` [M,2]) and, for validation data, the image id strings
themselves in the header (`image_ids`, with `img_ref` [M,2] int32 = (index into image_ids, object index)).
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional

import numpy as np

MAGIC = b"DCDW1\x00\x00\x00"
ALIGN = 64
NUM_KPTS = 73                      # GMW reads the first 73 keypoints of every object (dataset_utilities.py:45-46)
TRAIN_KEYS = ("kpts_2d", "kpts_3d", "pred_rot", "gt_location", "img_idx", "dim")


def load_reference_json(path: str, split: str) -> Dict[str, np.ndarray]:
    """`load_data(args, dataset_split)` of the reference on one JSON file (`split` = "train" or "valid")."""
    out: Dict[str, list] = {k: [] for k in TRAIN_KEYS}
    with open(path, "r") as f:
        data = json.load(f)
    image_ids: List[str] = []
    img_ref: List[tuple] = []
    if split == "train":
        for i in range(len(data["kpts_2d"])):
            for j in range(len(data["kpts_2d"][i])):
                out["kpts_2d"].append(np.array(data["kpts_2d"][i][j]))
                out["kpts_3d"].append(np.array(data["kpts_3d"][i][j]))
                out["pred_rot"].append([data["pred_rot"][i][j]])
                out["gt_location"].append(np.array(data["gt_location"][i][j]))
                out["img_idx"].append((0, 0))
    elif split == "valid":
        for img in data.keys():
            image_ids.append(img)
            for i, a in enumerate(data[img]):
                out["kpts_2d"].append(np.array(a["kpts_2d"]).astype(np.float32).reshape((-1, 2))[:NUM_KPTS, :])
                out["kpts_3d"].append(np.array(a["kpts_3d"]).astype(np.float32).reshape((-1, 3))[:NUM_KPTS, :])
                out["pred_rot"].append(np.array(a["pred_rot"]).astype(np.float32))
                out["dim"].append(a["dim"])
                out["gt_location"].append(np.array(a["pred_location"]).astype(np.float32))
                out["img_idx"].append((img, int(i)))
                img_ref.append((len(image_ids) - 1, int(i)))
    else:
        raise ValueError("split must be 'train' or 'valid'")
    res = {k: np.array(v, dtype=np.float32) for k, v in out.items()}
    if split == "valid":
        res["image_ids"] = image_ids                      # extras (not in the reference's dict): lossless image ids
        res["img_ref"] = np.array(img_ref, dtype=np.int32).reshape(-1, 2)
    return res


def save(path: str, data: Dict[str, object]) -> int:
    """Write `data` (arrays + optional `image_ids` list) as one binary file; returns the number of bytes written."""
    arrays = {}
    for k, v in data.items():
        if k == "image_ids":
            continue
        a = np.ascontiguousarray(v)
        if a.dtype.kind == "f":
            a = a.astype("<f4", copy=False)
        elif a.dtype.kind in "iu":
            a = a.astype("<i4", copy=False)
        else:
            raise TypeError("unsupported dtype for %s: %s" % (k, a.dtype))
        arrays[k] = a
    table = []
    offset = 0
    for k, a in arrays.items():
        offset = (offset + ALIGN - 1) // ALIGN * ALIGN
        table.append({"name": k, "dtype": a.dtype.str, "shape": list(a.shape), "offset": offset, "nbytes": int(a.nbytes)})
        offset += a.nbytes
    header = json.dumps({"version": 1, "arrays": table, "image_ids": list(data.get("image_ids", []))}).encode("utf-8")
    pre = len(MAGIC) + 8 + len(header)
    base = (pre + ALIGN - 1) // ALIGN * ALIGN
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", len(header), base))
        f.write(header)
        f.write(b"\x00" * (base - pre))
        pos = 0
        for ent, a in zip(table, arrays.values()):
            f.write(b"\x00" * (ent["offset"] - pos))
            f.write(a.tobytes())
            pos = ent["offset"] + a.nbytes
        total = base + pos
    os.replace(tmp, path)
    return total


def load(path: str, mmap: bool = True) -> Dict[str, object]:
    """Read a file written by `save`.  mmap=True: the arrays are read-only views of the mapped file (zero parse, zero copy)."""
    with open(path, "rb") as f:
        head = f.read(len(MAGIC) + 8)
        if len(head) < len(MAGIC) + 8 or head[:len(MAGIC)] != MAGIC:
            raise ValueError("%s is not a DCDW1 file" % path)
        hlen, base = struct.unpack("<II", head[len(MAGIC):])
        meta = json.loads(f.read(hlen).decode("utf-8"))
    if meta.get("version") != 1:
        raise ValueError("unsupported DCDW version %r" % meta.get("version"))
    size = os.path.getsize(path)
    out: Dict[str, object] = {}
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    for ent in meta["arrays"]:
        lo = base + ent["offset"]
        if lo + ent["nbytes"] > size:
            raise ValueError("truncated DCDW1 file: %s" % path)
        a = np.frombuffer(buf, dtype=np.dtype(ent["dtype"]), count=ent["nbytes"] // np.dtype(ent["dtype"]).itemsize, offset=lo)
        out[ent["name"]] = a.reshape(ent["shape"])
    if meta.get("image_ids"):
        out["image_ids"] = meta["image_ids"]
    return out


def convert_json(json_path: str, out_path: str, split: str) -> Dict[str, float]:
    """JSON -> binary once; returns the two file sizes."""
    data = load_reference_json(json_path, split)
    nbytes = save(out_path, data)
    return {"json_bytes": float(os.path.getsize(json_path)), "binary_bytes": float(nbytes)}


def from_detector(kps_img, kps_3d, pred_rot, locations, P, dim=None, image_ids: Optional[List[str]] = None,
                  img_ref=None) -> Dict[str, object]:
    """The record set of `generate_data` (detector_loss.py:148-173) / `inference.py:59-84` straight from DGDE tensors:
    kps_img [M,n,2] image-space keypoints, P the image's 3x4 projection matrix (or [M,3,4]); the 2D keypoints are
    normalised as in detector_loss.py:149-155 ((u - cx) / fx, (v - cy) / fy).  Accepts torch tensors (any device) or
    numpy arrays; returns numpy float32 arrays in the layout of `load_reference_json`."""
    def np32(x):
        if hasattr(x, "detach"):
            x = x.detach().cpu().numpy()
        return np.asarray(x, dtype=np.float32)
    kps = np32(kps_img)[:, :NUM_KPTS, :]
    Pm = np.asarray(P.detach().cpu().numpy() if hasattr(P, "detach") else P)
    Pm = np.broadcast_to(Pm.reshape((-1, 3, 4)), (kps.shape[0], 3, 4)).astype(np.float32)
    norm = np.empty_like(kps)
    norm[:, :, 0] = (kps[:, :, 0] - Pm[:, None, 0, 2]) / Pm[:, None, 0, 0]
    norm[:, :, 1] = (kps[:, :, 1] - Pm[:, None, 1, 2]) / Pm[:, None, 1, 1]
    M = kps.shape[0]
    out: Dict[str, object] = {
        "kpts_2d": norm, "kpts_3d": np32(kps_3d)[:, :NUM_KPTS, :], "pred_rot": np32(pred_rot).reshape(M, 1),
        "gt_location": np32(locations).reshape(M, 3),
        "dim": np32(dim).reshape(M, 3) if dim is not None else np.zeros((0,), dtype=np.float32),
    }
    if image_ids is not None:
        ref = np.asarray(img_ref, dtype=np.int32).reshape(M, 2)
        out["image_ids"] = list(image_ids)
        out["img_ref"] = ref
        out["img_idx"] = np.stack([np.array([float(image_ids[i]) for i in ref[:, 0]], dtype=np.float32),
                                   ref[:, 1].astype(np.float32)], axis=1)
    else:
        out["img_idx"] = np.zeros((M, 2), dtype=np.float32)
    return out


class Dataset:
    """Mirror of `GMW/utilities/dataset_utilities.py:58-73` on a binary file, a reference JSON file or a record dict."""

    def __init__(self, dataset_split: str, source, batch_size: int = 1, mmap: bool = True):
        self.batch_size = batch_size
        self.dataset_split = dataset_split
        if isinstance(source, dict):
            self.data = source
        elif str(source).endswith(".json"):
            self.data = load_reference_json(source, dataset_split)
        else:
            self.data = load(source, mmap=mmap)
        self.len = len(self.data["kpts_2d"])

    def __getitem__(self, index):
        d = self.data
        if self.dataset_split == "valid":
            return d["kpts_2d"][index], d["kpts_3d"][index], d["pred_rot"][index], d["gt_location"][index], d["dim"][index], d["img_idx"][index]
        return d["kpts_2d"][index], d["kpts_3d"][index], d["pred_rot"][index], d["gt_location"][index], d["img_idx"][index]

    def __len__(self):
        return self.len
