"""Input side of the path (SURVEY.md section 8 f row 4): the 2-D feature pyramids NeuConNet back-projects.

Drop-in for models/backbone.py:22-77 (`MnasMulti`: same constructor, submodule names and state-dict keys) and for the two
list comprehensions of models/neuralrecon.py:53-54 that call each backbone once per view (2 x 9 calls of batch 1).  The
convolutions stay library calls (cuDNN through torch -- like the dense 2-D fusion convs, outside section 8's hand-written
scope); what this module changes is the CALL PATTERN:

  * `MnasMulti.forward_views(imgs)` runs all V views of a fragment batch as ONE call per backbone.  The reference evaluates
    in train() mode (main.py:357), so every BatchNorm2d normalises with the statistics of ITS call, i.e. of one view; a plain
    V-image batch would mix the views' statistics and change the features.  `ViewBatchNorm2d` keeps the per-view statistics
    inside the batched call (views folded into the channel axis for the normalisation: [V*B, C, H, W] -> [B, V*C, H, W], free
    for B == 1) and applies the V running-statistics updates in view order, so outputs AND buffers equal V separate calls.
  * `FeatureExtractor` is the image half of NeuralRecon.forward: normalise, both backbones batched, and the per-view /
    per-level lists NeuConNet.forward expects (views of the batched outputs, no copies).

The architecture comes from torchvision's MNASNet; weights come from the checkpoint's state dict (the reference downloads
ImageNet weights at construction when alpha == 1.0 -- there is no network here, and a checkpoint overwrites them anyway).
Device-agnostic torch modules (this is host-side glue + library convs, not a CUDA kernel of this package).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _get_depths(alpha):
    """Channel widths of the MNASNet stages (torchvision's own rounding rule; the reference restates it, backbone.py:6-19)."""
    from torchvision.models.mnasnet import _get_depths as tv_depths
    return tv_depths(alpha)


class ViewBatchNorm2d(nn.BatchNorm2d):
    """BatchNorm2d whose training-mode statistics are taken per VIEW of a view-major batch ([V*B, C, H, W], view v = rows
    v*B .. v*B+B-1) when `views` > 1 -- bit-for-bit the arithmetic of V separate calls up to the reduction order."""

    views = 1

    def forward(self, x):
        V = self.views
        if V <= 1 or not (self.training or not self.track_running_stats):
            return super().forward(x)
        vb, c, h, w = x.shape
        assert vb % V == 0, "the batch is not a whole number of views"
        b = vb // V
        folded = x.view(V, b, c, h, w).transpose(0, 1).reshape(b, V * c, h, w)
        track = self.training and self.track_running_stats and self.running_mean is not None
        stat_m = stat_v = None
        if track:       # momentum 1 on scratch buffers: they receive exactly this call's mean / unbiased variance per (view, channel)
            stat_m = torch.zeros(V * c, dtype=self.running_mean.dtype, device=x.device)
            stat_v = torch.ones(V * c, dtype=self.running_var.dtype, device=x.device)
        y = F.batch_norm(folded, stat_m, stat_v, self.weight.repeat(V) if self.weight is not None else None,
                         self.bias.repeat(V) if self.bias is not None else None, True, 1.0, self.eps)
        if track:
            with torch.no_grad():
                stat_m, stat_v = stat_m.view(V, c), stat_v.view(V, c)
                for v in range(V):      # the reference's V calls update the buffers one after the other
                    self.num_batches_tracked += 1
                    m = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                    self.running_mean.mul_(1 - m).add_(stat_m[v], alpha=m)
                    self.running_var.mul_(1 - m).add_(stat_v[v], alpha=m)
        return y.view(b, V, c, h, w).transpose(0, 1).reshape(vb, c, h, w)


class MnasMulti(nn.Module):
    """MNASNet trunk + 3-level FPN head -> [1/4 (depths[2] ch), 1/8 (depths[3]), 1/16 (depths[4])] feature maps."""

    def __init__(self, alpha=1.0):
        super().__init__()
        import torchvision
        depths = _get_depths(alpha)
        trunk = torchvision.models.MNASNet(alpha=alpha).layers
        self.conv0 = nn.Sequential(*[trunk[i] for i in range(9)])
        self.conv1 = trunk[9]
        self.conv2 = trunk[10]
        self.out1 = nn.Conv2d(depths[4], depths[4], 1, bias=False)
        self.out_channels = [depths[4]]
        final_chs = depths[4]
        self.inner1 = nn.Conv2d(depths[3], final_chs, 1, bias=True)
        self.inner2 = nn.Conv2d(depths[2], final_chs, 1, bias=True)
        self.out2 = nn.Conv2d(final_chs, depths[3], 3, padding=1, bias=False)
        self.out3 = nn.Conv2d(final_chs, depths[2], 3, padding=1, bias=False)
        self.out_channels += [depths[3], depths[2]]
        for m in self.modules():
            if type(m) is nn.BatchNorm2d:
                m.__class__ = ViewBatchNorm2d       # same parameters / buffers / state-dict keys

    def forward(self, x):
        conv0 = self.conv0(x)
        conv1 = self.conv1(conv0)
        conv2 = self.conv2(conv1)
        out16 = self.out1(conv2)
        top = F.interpolate(conv2, scale_factor=2, mode="nearest") + self.inner1(conv1)
        out8 = self.out2(top)
        top = F.interpolate(top, scale_factor=2, mode="nearest") + self.inner2(conv0)
        out4 = self.out3(top)
        return [out4, out8, out16]

    def forward_views(self, imgs):
        """imgs: [V, B, 3, H, W] (or a list of V [B, 3, H, W] tensors) -> list over views of [1/4, 1/8, 1/16] maps, each
        [B, C, h, w]: views of ONE batched call with per-view BatchNorm statistics (= V calls of forward)."""
        if isinstance(imgs, (list, tuple)):
            imgs = torch.stack(list(imgs))
        V, B = imgs.shape[:2]
        bns = [m for m in self.modules() if isinstance(m, ViewBatchNorm2d)]
        for m in bns:
            m.views = V
        try:
            levels = self.forward(imgs.reshape(V * B, *imgs.shape[2:]))
        finally:
            for m in bns:
                m.views = 1
        levels = [lv.view(V, B, *lv.shape[1:]) for lv in levels]
        return [[lv[v] for lv in levels] for v in range(V)]


class FeatureExtractor(nn.Module):
    """The image half of NeuralRecon.forward (models/neuralrecon.py:25-54): `backbone2d` (TSDF branch) and `backbone_occ_pano`
    (occupancy / panoptic branch) under the reference's attribute names, two batched calls instead of 2 x V."""

    def __init__(self, alpha=1.0, pixel_mean=(103.53, 116.28, 123.675), pixel_std=(1.0, 1.0, 1.0)):
        super().__init__()
        self.backbone2d = MnasMulti(alpha)
        self.backbone_occ_pano = MnasMulti(alpha)
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean, dtype=torch.float32).view(-1, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std, dtype=torch.float32).view(-1, 1, 1), persistent=False)

    def normalizer(self, x):
        return (x - self.pixel_mean.type_as(x)) / self.pixel_std.type_as(x)

    def forward(self, imgs):
        """imgs [B, V, 3, H, W] (inputs['imgs']) -> (features_backbone2d, features_backbone2d_occ_pano): lists over the V views
        of [1/4, 1/8, 1/16] maps -- the first two arguments of NeuConNet.forward."""
        x = self.normalizer(imgs).transpose(0, 1)          # view-major
        return self.backbone2d.forward_views(x), self.backbone_occ_pano.forward_views(x)
