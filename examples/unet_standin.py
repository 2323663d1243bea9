"""Stand-in for the reference's single-view generator (development/multiImage_pytorch/models.py:208-346), used by
the configs[4] measurements (bench.py ``train_step_c5``, examples/train_step_c5.py).

The reference's model is outside this repository's scope and is used unchanged by its own training script; what the
full-step measurement needs is a network of the SAME SIZE AND SHAPE in front of the loss: 8 stride-2 encoders and 8
up-sampling decoders (nearest up-sampling followed by two 4x4 convolutions) with skip connections, instance
normalisation, and a fully-connected "global track" that exchanges channel means with the convolutional track
(merge layers) - 79.99 M parameters in fp32 = 320 MB of gradients per all-reduce, like the reference's
``SingleViewModel`` (80.0 M).  Written from the layer table below, not from the reference's code.
The output is the 9-channel encoding in [-1, 1] after tanh, which goes straight into ``MixedLoss.forward_encoded``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

WIDTHS = (64, 128, 256, 512, 512, 512, 512, 512)        # encoder output channels at 128, 64, ... 1 pixels (256x256 input)


class _Stage(nn.Module):
    """conv block -> channel means (to the global track) -> instance norm -> + linear(global features)."""

    def __init__(self, conv, channels, norm, act):
        super().__init__()
        self.conv, self.act = conv, act
        self.norm = nn.InstanceNorm2d(channels, affine=True) if norm else None
        self.merge = nn.Linear(channels, channels, bias=False)

    def forward(self, x, g):
        if self.act:
            x = F.leaky_relu(x, 0.2)
        x = self.conv(x)
        mean = x.mean(dim=(2, 3))
        if self.norm is not None:
            x = self.norm(x)
        if g is not None:
            x = x + self.merge(g)[:, :, None, None]
        return x, mean


def _down(cin, cout):
    return nn.Conv2d(cin, cout, 4, stride=2, padding=1, bias=False)


def _up(cin, cout):
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(cin, cout, 4, bias=False),
                         nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(cout, cout, 4, bias=False))


class _Global(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.fc = nn.Linear(cin, cout)

    def forward(self, mean, g):
        return F.selu(self.fc(mean if g is None else torch.cat((g, mean), dim=1)))


class UNetStandIn(nn.Module):
    """[B,3,256,256] -> [B,9,256,256] in [-1,1]."""

    def __init__(self, out_channels=9, widths=WIDTHS, in_channels=3):
        super().__init__()
        n = len(widths)
        enc_in = (in_channels,) + tuple(widths[:-1])
        self.enc = nn.ModuleList(_Stage(_down(enc_in[i], widths[i]), widths[i], norm=0 < i < n - 1, act=i > 0) for i in range(n))
        # global track after encoder i feeds encoder i+1 (so it has that stage's width); the last one feeds the first decoder
        g_out = tuple(widths[1:]) + (widths[-1],)
        g_in = (in_channels,) + tuple(2 * w for w in widths[1:])
        self.genc = nn.ModuleList(_Global(g_in[i], g_out[i]) for i in range(n))
        dec_out = tuple(widths[-2::-1]) + (out_channels,)                                  # 512,512,512,512,256,128,64,9
        dec_in = (widths[-1],) + tuple(2 * c for c in dec_out[:-1])
        self.dec = nn.ModuleList(_Stage(_up(dec_in[i], dec_out[i]), dec_out[i], norm=i < n - 1, act=True) for i in range(n))
        self.drop = nn.ModuleList(nn.Dropout(0.5) if i < 3 else nn.Identity() for i in range(n))
        gd_out = dec_out[1:] + (out_channels,)
        self.gdec = nn.ModuleList(_Global(2 * dec_out[i], gd_out[i]) for i in range(n))
        # Two layers exist for the shape's sake only, as in the reference: the first encoder has no global features to
        # merge and nothing consumes the global track after the last decoder.  They never receive a gradient, so they are
        # frozen (4,267 of the 79.99 M parameters) - DistributedDataParallel then needs no unused-parameter search.
        for p in list(self.enc[0].merge.parameters()) + list(self.gdec[-1].parameters()):
            p.requires_grad_(False)

    def forward(self, x):
        g = self.genc[0](x.mean(dim=(2, 3)), None)
        skips = []
        for i, stage in enumerate(self.enc):
            x, mean = stage(x, None if i == 0 else g)
            if i > 0:
                g = self.genc[i](mean, g)
            skips.append(x)
        for i, stage in enumerate(self.dec):
            if i > 0:
                x = torch.cat((x, skips[-1 - i]), dim=1)
            x, mean = stage(x, g)
            x = self.drop[i](x)
            g = self.gdec[i](mean, g)
        return torch.tanh(x)


if __name__ == "__main__":
    m = UNetStandIn()
    print("%.3f M parameters" % (sum(p.numel() for p in m.parameters()) / 1e6))
    with torch.no_grad():
        print(m(torch.rand(1, 3, 256, 256)).shape)
