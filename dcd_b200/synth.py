"""Seeded synthetic KITTI-shaped objects for tests and benchmarks.

Geometry-consistent: a 3D keypoint template in the object frame (n-10 points inside the
box, then the 8 box corners and the bottom/top centres in the order of
DGDE/data/datasets/kitti_utils.py:136-147) is rotated by the yaw, translated, projected
through KITTI's P2 and perturbed by pixel noise, so the per-edge depths scatter around the
true depth with realistic conditioning (a few % of the edges hit a clamp).
Everything is generated on the CPU with a seeded torch.Generator; no dataset is read.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

BASE_SEED = 20220710
# KITTI P2 (fx=fy, cx, cy and the 4th column), SURVEY 8d
P2 = ((721.5377, 0.0, 609.5593, 44.85728),
      (0.0, 721.5377, 172.854, 0.2163791),
      (0.0, 0.0, 1.0, 0.002745884))
# car dimension prior (l, h, w): DGDE/config/defaults.py:229-236
DIM_MEAN = (3.884, 1.5261, 1.6286)
DIM_STD = (0.4259, 0.1367, 0.1022)
IMG_W, IMG_H = 1242.0, 375.0


@dataclass
class Objects:
    """A batch of N objects with n keypoints each (all CPU float32 unless noted)."""
    kps: torch.Tensor          # [N,n,2] pixel coordinates (DGDE form)
    kps_norm: torch.Tensor     # [N,n,2] ((u-cx)/fx, (v-cy)/fy) (GMW form, detector_loss.py:149-155)
    kps_3d: torch.Tensor       # [N,n,3] object-frame template (bottom-centre origin)
    rot_y: torch.Tensor        # [N,1]
    K: torch.Tensor            # [N,3,4] projection matrix of the object's frame
    mask: torch.Tensor         # [N,n] bool keypoint visibility
    gt_depth: torch.Tensor     # [N]
    frame_id: torch.Tensor     # [N] int64
    counts: torch.Tensor       # [F] int64 objects per frame

    @property
    def N(self) -> int:
        return self.kps.shape[0]

    @property
    def n(self) -> int:
        return self.kps.shape[1]

    def slice(self, lo: int, hi: int) -> "Objects":
        return Objects(self.kps[lo:hi], self.kps_norm[lo:hi], self.kps_3d[lo:hi], self.rot_y[lo:hi],
                       self.K[lo:hi], self.mask[lo:hi], self.gt_depth[lo:hi], self.frame_id[lo:hi],
                       self.counts)


def frame_counts(frames: int, max_objects: int, ragged: bool, seed: int) -> torch.Tensor:
    """Objects per frame: `max_objects` each, or U{1..max_objects} when ragged."""
    if not ragged:
        return torch.full((frames,), max_objects, dtype=torch.int64)
    g = torch.Generator().manual_seed(seed ^ 0x5EED)
    return torch.randint(1, max_objects + 1, (frames,), generator=g, dtype=torch.int64)


def make_objects(N: Optional[int] = None, n: int = 73, seed: int = BASE_SEED, counts: Optional[torch.Tensor] = None,
                 noise_px: float = 0.5, jitter_fx: float = 0.0, dtype=torch.float32) -> Objects:
    """Generate objects.  Give either N (frames of 50 objects, last one short) or `counts`."""
    if counts is None:
        assert N is not None
        full, rem = divmod(N, 50)
        counts = torch.tensor([50] * full + ([rem] if rem else []), dtype=torch.int64)
    N = int(counts.sum())
    F = counts.numel()
    g = torch.Generator().manual_seed(seed)
    f64 = torch.float64

    def rand(*shape):
        return torch.rand(*shape, generator=g, dtype=f64)

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=f64)

    frame_id = torch.repeat_interleave(torch.arange(F, dtype=torch.int64), counts)
    # per-frame intrinsics
    P = torch.tensor(P2, dtype=f64).unsqueeze(0).repeat(F, 1, 1)
    if jitter_fx > 0:
        s = 1.0 + jitter_fx * (2 * rand(F) - 1)
        P[:, 0, 0] *= s
        P[:, 1, 1] *= s
    K = P[frame_id]                                                  # [N,3,4]

    dims = torch.tensor(DIM_MEAN, dtype=f64) + torch.tensor(DIM_STD, dtype=f64) * randn(N, 3)
    dims = dims.clamp_min(0.5)
    l, h, w = dims[:, 0], dims[:, 1], dims[:, 2]
    tz = 5.0 + 55.0 * rand(N)
    tx = (1.2 * rand(N) - 0.6) * tz
    ty = 1.65 + 0.2 * randn(N)
    ry = (2 * rand(N) - 1) * math.pi

    n_free = max(n - 10, 0)
    pts = torch.empty(N, n, 3, dtype=f64)
    if n_free:
        r = rand(N, n_free, 3)
        pts[:, :n_free, 0] = (r[:, :, 0] - 0.5) * l[:, None]
        pts[:, :n_free, 1] = -r[:, :, 1] * h[:, None]
        pts[:, :n_free, 2] = (r[:, :, 2] - 0.5) * w[:, None]
    sx = torch.tensor([1, 1, -1, -1, 1, 1, -1, -1], dtype=f64) * 0.5
    sy = torch.tensor([0, 0, 0, 0, -1, -1, -1, -1], dtype=f64)
    sz = torch.tensor([1, -1, -1, 1, 1, -1, -1, 1], dtype=f64) * 0.5
    fixed = torch.zeros(N, 10, 3, dtype=f64)
    fixed[:, :8, 0] = sx[None] * l[:, None]
    fixed[:, :8, 1] = sy[None] * h[:, None]
    fixed[:, :8, 2] = sz[None] * w[:, None]
    fixed[:, 9, 1] = -h
    k_fixed = min(10, n)
    pts[:, n - k_fixed:] = fixed[:, 10 - k_fixed:]

    c, s_ = torch.cos(ry)[:, None], torch.sin(ry)[:, None]
    xc = c * pts[:, :, 0] + s_ * pts[:, :, 2] + tx[:, None]
    yc = pts[:, :, 1] + ty[:, None]
    zc = -s_ * pts[:, :, 0] + c * pts[:, :, 2] + tz[:, None]
    den = zc + K[:, 2, 3][:, None]
    u = (K[:, 0, 0][:, None] * xc + K[:, 0, 2][:, None] * zc + K[:, 0, 3][:, None]) / den
    v = (K[:, 1, 1][:, None] * yc + K[:, 1, 2][:, None] * zc + K[:, 1, 3][:, None]) / den
    u = u + noise_px * randn(N, n)
    v = v + noise_px * randn(N, n)
    kps = torch.stack((u, v), dim=-1)
    inside = (u >= 0) & (u < IMG_W) & (v >= 0) & (v < IMG_H)
    mask = (rand(N, n) < 0.85) & inside

    kps32 = kps.to(dtype)
    K32 = K.to(dtype)
    kps_norm = torch.stack(((kps32[:, :, 0] - K32[:, None, 0, 2]) / K32[:, None, 0, 0],
                            (kps32[:, :, 1] - K32[:, None, 1, 2]) / K32[:, None, 1, 1]), dim=-1)
    return Objects(kps=kps32.contiguous(), kps_norm=kps_norm.contiguous(), kps_3d=pts.to(dtype).contiguous(),
                   rot_y=ry.to(dtype).unsqueeze(-1).contiguous(), K=K32.contiguous(), mask=mask.contiguous(),
                   gt_depth=tz.to(dtype), frame_id=frame_id, counts=counts)


def kitti_val_batch(ragged: bool = True, frames: int = 3769, max_objects: int = 50, n: int = 73,
                    seed: int = BASE_SEED + 1) -> Objects:
    """BASELINE.json configs[1]: 3769 frames, <=50 objects per frame, 73 keypoints."""
    return make_objects(n=n, seed=seed, counts=frame_counts(frames, max_objects, ragged, seed))


def random_state_dict(seed: int, depth: int = 12, channels: int = 128):
    """Seeded random GMW weights with torch's Conv1d default init bounds, keyed like the reference
    state_dict (FeatureExtractor{4,6}d.conv_in / conv_<k>.{preconv,conv1,conv2}); for benchmarks."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cin):
        bound = 1.0 / math.sqrt(cin)
        sd[key + ".0.weight"] = (torch.rand((channels, cin, 1), generator=g) * 2 - 1) * bound
        sd[key + ".0.bias"] = (torch.rand((channels,), generator=g) * 2 - 1) * bound

    for name, cin in (("FeatureExtractor4d", 4), ("FeatureExtractor6d", 6)):
        conv(name + ".conv_in", cin)
        for k in range(depth):
            for sub in ("preconv", "conv1", "conv2"):
                conv("%s.conv_%d.%s" % (name, k, sub), channels)
    return sd
