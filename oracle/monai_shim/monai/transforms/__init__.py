import torch
import torch.nn.functional as F


class Transform:
    pass


class SpatialPad(Transform):
    """Symmetric zero-pad (channel-first, no batch dim) up to at least `spatial_size`."""

    def __init__(self, spatial_size, method="symmetric", mode="constant", **kwargs):
        self.spatial_size = spatial_size

    def __call__(self, img):
        sp = img.shape[1:]
        size = list(self.spatial_size) if isinstance(self.spatial_size, (list, tuple)) else [self.spatial_size] * len(sp)
        pads = []
        for d in reversed(range(len(sp))):
            tot = max(size[d] - sp[d], 0) if size[d] > 0 else 0
            pads += [tot // 2, tot - tot // 2]
        return F.pad(img, pads)


class CenterSpatialCrop(Transform):
    """Centre crop (channel-first, no batch dim); non-positive roi entries keep the full dim."""

    def __init__(self, roi_size, lazy=False):
        self.roi_size = roi_size

    def __call__(self, img):
        sp = img.shape[1:]
        roi = list(self.roi_size) if isinstance(self.roi_size, (list, tuple)) else [self.roi_size] * len(sp)
        sl = [slice(None)]
        for d in range(len(sp)):
            r = sp[d] if roi[d] <= 0 else min(roi[d], sp[d])
            c = sp[d] // 2
            start = max(c - r // 2, 0)
            sl.append(slice(start, start + r))
        return img[tuple(sl)]
