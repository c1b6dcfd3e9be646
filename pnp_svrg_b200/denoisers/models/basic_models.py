"""Class needed to unpickle the reference's MMO checkpoints, which are whole pickled
``DataParallel(models.basic_models.simple_CNN)`` modules (reference denoisers/models/basic_models.py:8-38,
MMODenoise.py:68-71).  Only the attribute layout matters here; inference runs in csrc/cnn_*.cuh."""
import torch.nn as nn


class simple_CNN(nn.Module):
    def __init__(self, n_ch_in=3, n_ch_out=3, n_ch=64, nl_type='relu', depth=5, bn=False):
        super().__init__()
        self.nl_type, self.depth, self.bn = nl_type, depth, bn
        self.in_conv = nn.Conv2d(n_ch_in, n_ch, 3, 1, 1, bias=True)
        self.conv_list = nn.ModuleList([nn.Conv2d(n_ch, n_ch, 3, 1, 1, bias=True) for _ in range(depth - 2)])
        self.out_conv = nn.Conv2d(n_ch, n_ch_out, 3, 1, 1, bias=True)
        if nl_type == 'relu':
            self.nl_list = nn.ModuleList([nn.LeakyReLU() for _ in range(depth - 1)])
        if bn:
            self.bn_list = nn.ModuleList([nn.BatchNorm2d(n_ch) for _ in range(depth - 2)])

    def forward(self, x_in):
        x = self.nl_list[0](self.in_conv(x_in))
        for i in range(self.depth - 2):
            x = self.conv_list[i](x)
            if self.bn:
                x = self.bn_list[i](x)
            x = self.nl_list[i + 1](x)
        return self.out_conv(x) + x_in
