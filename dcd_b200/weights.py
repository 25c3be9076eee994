"""Parameter-blob <-> reference state_dict conversion for the GMW edge nets.

The reference checkpoint contract is the GMW `state_dict`:
`FeatureExtractor{4,6}d.conv_in.0.{weight,bias}` and
`FeatureExtractor{4,6}d.conv_<k>.{preconv,conv1,conv2}.0.{weight,bias}` with Conv1d weights
[out, in, 1] (GMW/model/yi2018cvpr/model.py:21-47, ops.py:59).  The kernels read one flat FP32
blob per net (layout in include/dcd_b200.h): every matrix transposed to [in][out].
"""
from __future__ import annotations

from typing import Dict

import torch

NET_CH = 128
NET_DEPTH = 12
NET_NAMES = (("FeatureExtractor4d", 4), ("FeatureExtractor6d", 6))
_SUBS = ("preconv", "conv1", "conv2")


def blob_size(cin: int, depth: int = NET_DEPTH) -> int:
    return cin * NET_CH + NET_CH + depth * 3 * (NET_CH * NET_CH + NET_CH)


def _entries(prefix: str, cin: int, depth: int):
    """(state_dict key stem, offset of W^T, offset of b, in_features) in blob order."""
    off = 0
    yield prefix + ".conv_in", off, off + cin * NET_CH, cin
    off += cin * NET_CH + NET_CH
    for k in range(depth):
        for sub in _SUBS:
            yield "%s.conv_%d.%s" % (prefix, k, sub), off, off + NET_CH * NET_CH, NET_CH
            off += NET_CH * NET_CH + NET_CH


def pack_state_dict(sd: Dict[str, torch.Tensor], prefix: str, cin: int, depth: int = NET_DEPTH) -> torch.Tensor:
    """Reference state_dict -> flat FP32 blob (CPU or wherever `sd` lives)."""
    any_t = next(iter(sd.values()))
    blob = torch.empty(blob_size(cin, depth), dtype=torch.float32, device=any_t.device)
    for stem, ow, ob, fin in _entries(prefix, cin, depth):
        w = sd[stem + ".0.weight"].detach().to(torch.float32).reshape(NET_CH, fin)
        blob[ow:ow + fin * NET_CH] = w.t().reshape(-1)
        blob[ob:ob + NET_CH] = sd[stem + ".0.bias"].detach().to(torch.float32)
    return blob


def pack_parameters(params: Dict[str, torch.Tensor], prefix: str, cin: int, depth: int = NET_DEPTH) -> torch.Tensor:
    """Differentiable form of pack_state_dict: `params` maps the reference names to the LIVE parameters
    (dict(model.named_parameters())); the blob is one torch.cat, so autograd carries the blob gradient back to
    every reference parameter's .grad (transposes and splits) without any bookkeeping by the caller."""
    parts = []
    for stem, _, _, fin in _entries(prefix, cin, depth):
        parts.append(params[stem + ".0.weight"].to(torch.float32).reshape(NET_CH, fin).t().reshape(-1))
        parts.append(params[stem + ".0.bias"].to(torch.float32).reshape(-1))
    return torch.cat(parts)


def unpack_blob(blob: torch.Tensor, prefix: str, cin: int, depth: int = NET_DEPTH) -> Dict[str, torch.Tensor]:
    """Flat blob (parameters or gradients) -> tensors named and shaped like the reference state_dict."""
    out = {}
    for stem, ow, ob, fin in _entries(prefix, cin, depth):
        out[stem + ".0.weight"] = blob[ow:ow + fin * NET_CH].reshape(fin, NET_CH).t().reshape(NET_CH, fin, 1)
        out[stem + ".0.bias"] = blob[ob:ob + NET_CH]
    return out
