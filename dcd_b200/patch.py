"""Attribute-patch the reference so that its pipelines call the B200 kernels.

The reference has no plugin registry for this path: `Anno_Encoder.decode_pairs_kpts_depth` is a
plain method and `compute_z` / `compute_reg_loss` are module-level functions of GMW/main.py
(SURVEY.md section 8b), so the drop-in is attribute assignment on the objects a maintainer passes in.

    import dcd_b200.patch as patch
    patch.install(anno_encoder_cls=Anno_Encoder)                       # DGDE (train + inference)
    patch.install(gmw_main=main_module, gmw_model=model.module)        # GMW  (train + validation)
    patch.uninstall()
"""
from __future__ import annotations

import types
from typing import List, Tuple

import torch

from . import ops

_undo: List[Tuple[object, str, object, bool]] = []


def _set(obj, name, value):
    had = name in vars(obj) if not isinstance(obj, type) else name in obj.__dict__
    _undo.append((obj, name, getattr(obj, name, None), had))
    if isinstance(obj, (type, types.ModuleType)):
        setattr(obj, name, value)
    else:
        object.__setattr__(obj, name, value)      # keeps nn.Module from registering a sub-module


def _decode_method(self, kps, kps_3d, rot_y, K, training=False, kpts_2d_mask=None, gt_depth=None, weight=None):
    return ops.decode_pairs_kpts_depth(kps, kps_3d, rot_y, K, training=training, kpts_2d_mask=kpts_2d_mask,
                                       gt_depth=gt_depth, weight=weight)


def install(anno_encoder_cls=None, gmw_main=None, gmw_model=None) -> None:
    """Patch whichever reference objects are given.

    anno_encoder_cls: the reference `Anno_Encoder` class (or an instance).
    gmw_main:         the imported GMW/main.py module (compute_z, compute_reg_loss are replaced).
    gmw_model:        a reference `GMW` nn.Module instance; its weights are copied into a
                      dcd_b200.ops.GMW and its forward is redirected (edge_P is returned as None).
    """
    if anno_encoder_cls is not None:
        target = anno_encoder_cls
        fn = _decode_method if isinstance(target, type) else types.MethodType(_decode_method, target)
        _set(target, "decode_pairs_kpts_depth", fn)
    if gmw_main is not None:
        _set(gmw_main, "compute_z", ops.compute_z)
        _set(gmw_main, "compute_reg_loss", ops.compute_reg_loss)
    if gmw_model is not None:
        dev = next(gmw_model.parameters()).device
        fast = ops.GMW().to(dev).load_reference_state_dict(gmw_model.state_dict())
        _set(gmw_model, "_dcd_b200", fast)

        def forward(kpts_2d, kpts_3d, pred_rot=None, args=None, _fast=fast):
            return _fast(kpts_2d, kpts_3d, pred_rot, args)

        _set(gmw_model, "forward", forward)


def uninstall() -> None:
    while _undo:
        obj, name, old, had = _undo.pop()
        if had:
            setattr(obj, name, old)
        else:
            try:
                object.__delattr__(obj, name) if not isinstance(obj, (type, types.ModuleType)) else delattr(obj, name)
            except AttributeError:
                pass


def sync_gradients_to_reference(gmw_model) -> None:
    """Copy the blob gradients of the patched model into the reference module's .grad fields so
    that the reference's optimizer (GMW/main.py:255,466) keeps working unchanged."""
    fast = gmw_model._dcd_b200
    grads = fast.reference_grads()
    with torch.no_grad():
        for k, p in gmw_model.named_parameters():
            p.grad = grads[k].reshape(p.shape).clone()
