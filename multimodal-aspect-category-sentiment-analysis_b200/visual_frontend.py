"""Visual front-end of the training / inference loops: the ResNet-152 grid and ROI features that
feed the fusion path. Reference: ``myResNetImg`` / ``myResNetRoI`` (fcmf_framework/resnet_utils.py:6-55) called ONE IMAGE
INDEX AT A TIME from the loops (run_multimodal_fcmf.py:449-460, 516-527, 619-629): 7 grid forwards + 7 x NR ROI forwards of
batch B each per step, outputs stacked on the host.

Here the same convolutional trunk (any module with torchvision's ResNet attributes: conv1, bn1, relu, maxpool, layer1-4 -- the
convolutions themselves are library code, cuDNN) is called ONCE per branch on all B x NI images (resp. B x NI x NR crops),
channels-last, and the adaptive 7x7 pooling / global mean and the [B, NI, 49, 2048] / [B, NI, NR, 2048] layouts the fusion
path reads are produced in one pass. Results are identical to the reference loop whenever BatchNorm uses running statistics
(eval(), inference, or a frozen trunk: ``if_fine_tune=False`` features are detached anyway, resnet_utils.py:26-28); with
BatchNorm in train() mode the batch statistics of one big call would differ from those of 7 calls of B images, so that case
keeps the reference's grouping (one call per image index) -- same results, still no Python loop over ROIs inside a group.
An optional feature cache (frozen trunk only) returns stored features for image keys seen before.
"""
from __future__ import annotations

from typing import Dict, Hashable, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor


def _trunk(resnet: nn.Module, x: Tensor) -> Tensor:
    x = resnet.maxpool(resnet.relu(resnet.bn1(resnet.conv1(x))))
    return resnet.layer4(resnet.layer3(resnet.layer2(resnet.layer1(x))))


def _bn_uses_batch_stats(m: nn.Module) -> bool:
    return any(isinstance(s, nn.modules.batchnorm._BatchNorm) and s.training and s.track_running_stats for s in m.modules()) or \
        any(isinstance(s, nn.modules.batchnorm._BatchNorm) and not s.track_running_stats for s in m.modules())


class VisualFrontEnd(nn.Module):
    """resnet_img / resnet_roi: the two trunks of run_multimodal_fcmf.py:224-227 (state_dict keys ``resnet_img.resnet.*``,
    ``resnet_roi.resnet.*`` as saved by the reference, :558-563)."""

    class _Wrap(nn.Module):                      # keeps the reference's ``.resnet`` attribute in the state_dict keys
        def __init__(self, resnet):
            super().__init__()
            self.resnet = resnet

    def __init__(self, resnet_img: nn.Module, resnet_roi: nn.Module, if_fine_tune: bool = False, att_size: int = 7,
                 channels_last: bool = True, cache: bool = False):
        super().__init__()
        self.resnet_img, self.resnet_roi = self._Wrap(resnet_img), self._Wrap(resnet_roi)
        self.if_fine_tune, self.att_size, self.channels_last = if_fine_tune, att_size, channels_last
        self._cache: Optional[Dict[Hashable, tuple]] = {} if cache else None

    def _run(self, resnet: nn.Module, x: Tensor, group: int) -> Tensor:
        """x [N, 3, h, w] -> trunk output; `group` images per call when BatchNorm needs the reference's batch statistics."""
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        ctx = torch.enable_grad() if self.if_fine_tune else torch.no_grad()
        with ctx:
            if _bn_uses_batch_stats(resnet) and group < x.shape[0]:
                return torch.cat([_trunk(resnet, x[i:i + group]) for i in range(0, x.shape[0], group)], 0)
            return _trunk(resnet, x)

    def grid_features(self, t_img_features: Tensor) -> Tensor:
        """[B, NI, 3, h, w] -> visual_embeds_att [B, NI, att*att, C] (run_multimodal_fcmf.py:449-452, 459)."""
        B, NI = t_img_features.shape[:2]
        # image-index-major order = the reference's call grouping (one call per image index, B images each)
        x = t_img_features.transpose(0, 1).reshape(NI * B, *t_img_features.shape[2:])
        f = F.adaptive_avg_pool2d(self._run(self.resnet_img.resnet, x, B), [self.att_size, self.att_size])
        C = f.shape[1]
        out = f.reshape(NI, B, C, self.att_size * self.att_size).permute(1, 0, 3, 2)
        return out if self.if_fine_tune else out.detach()

    def roi_features(self, roi_img_features: Tensor) -> Tensor:
        """[B, NI, NR, 3, h, w] (any float dtype; the loader yields float64, cast as at :445) -> roi_embeds_att [B, NI, NR, C]."""
        B, NI, NR = roi_img_features.shape[:3]
        x = roi_img_features.float().permute(1, 2, 0, 3, 4, 5).reshape(NI * NR * B, *roi_img_features.shape[3:])
        f = self._run(self.resnet_roi.resnet, x, B).mean(3).mean(2)
        out = f.reshape(NI, NR, B, -1).permute(2, 0, 1, 3)
        return out if self.if_fine_tune else out.detach()

    def forward(self, t_img_features: Tensor, roi_img_features: Tensor, keys: Optional[Sequence[Hashable]] = None):
        """-> (visual_embeds_att [B, NI, 49, C], roi_embeds_att [B, NI, NR, C]). keys: one hashable per sample; with the
        cache enabled (frozen trunk, eval-mode BatchNorm) samples seen before skip both trunks."""
        use_cache = (self._cache is not None and keys is not None and not self.if_fine_tune
                     and not _bn_uses_batch_stats(self.resnet_img) and not _bn_uses_batch_stats(self.resnet_roi))
        if not use_cache:
            return self.grid_features(t_img_features), self.roi_features(roi_img_features)
        miss = [i for i, k in enumerate(keys) if k not in self._cache]
        if miss:
            idx = torch.as_tensor(miss, device=t_img_features.device)
            g, r = self.grid_features(t_img_features.index_select(0, idx)), self.roi_features(roi_img_features.index_select(0, idx))
            for j, i in enumerate(miss):
                self._cache[keys[i]] = (g[j].clone(), r[j].clone())
        vis = torch.stack([self._cache[k][0] for k in keys], 0)
        roi = torch.stack([self._cache[k][1] for k in keys], 0)
        return vis, roi
