"""TEST INFRASTRUCTURE: a loss with the SHAPE of the reference's inversion loss (exp/cips3d/models/projector_v9.py:1106-1137:
VGG16 perceptual distance of the decoder's image and -- weighted 50x -- of the 64x64 thumb, plus an MSE term), built from
seeded random-init networks because no checkpoint can be downloaded here:
  * `vgg`     : the conv1_1 .. conv3_3 stack of VGG16 (3x3 convs + ReLU, two 2x2 max-pools), seeded Kaiming init;
  * `decoder` : a stand-in for the reference decoder's entry -- a seeded 1x1 conv 256 -> 3 on the feature map, tanh, bilinear
                x2 -- so that the loss sends a real cotangent through `feature_map`, as `renderer_detach=False` does.
Used identically by both arms of the loss-curve parity test (libc3dpp vs torch autograd of the reference restatement)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class RefShapedLoss(nn.Module):
    def __init__(self, seed=0, rgb_weight=1.0, thumb_weight=50.0, mse_weight=1.0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        chans = [(3, 64), (64, 64), "M", (64, 128), (128, 128), "M", (128, 256), (256, 256), (256, 256)]
        layers = []
        for c in chans:
            if c == "M":
                layers.append(nn.MaxPool2d(2))
                continue
            conv = nn.Conv2d(c[0], c[1], 3, padding=1)
            with torch.no_grad():
                conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (c[0] * 9)) ** 0.5)
                conv.bias.zero_()
            layers += [conv, nn.ReLU()]
        self.vgg = nn.Sequential(*layers)
        self.dec = nn.Conv2d(256, 3, 1)
        with torch.no_grad():
            self.dec.weight.copy_(torch.randn(self.dec.weight.shape, generator=g) * 0.5)
            self.dec.bias.zero_()
        self.rgb_weight, self.thumb_weight, self.mse_weight = rgb_weight, thumb_weight, mse_weight
        self.requires_grad_(False)

    def features(self, x):
        f = self.vgg(x)
        return f / (f.shape[1] * f.shape[2] * f.shape[3]) ** 0.5

    def forward(self, thumbs, targets, feats):
        img = F.interpolate(torch.tanh(self.dec(feats)), scale_factor=2, mode="bilinear", align_corners=False)
        tgt_img = F.interpolate(targets, scale_factor=2, mode="bilinear", align_corners=False)
        percep = (self.features(tgt_img) - self.features(img)).square().sum() * self.rgb_weight + \
                 (self.features(targets) - self.features(thumbs)).square().sum() * self.thumb_weight
        return percep + F.mse_loss(img, tgt_img) * self.mse_weight
