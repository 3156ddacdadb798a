"""Encoder with the reference's topology and parameter names (reference:
src/models/resnet_modified.py): a BatchNorm-free ResNet-18 layout with a 3x3 stem, circular
padding along the image width (the LiDAR image wraps around) and zero padding along the height,
strides (1,2) x3 then (2,2), tanh or relu.  `state_dict()` keys match the reference
(`conv1.weight`, `layerL.B.conv{1,2}.weight`, `layerL.0.downsample.0.weight`, `fc.{weight,bias}`)
so checkpoints interchange.

Round-1 status: the 2-D convolutions still run through torch's cuDNN path (library baseline);
the tcgen05 implicit-GEMM replacement is §8 row a6 / DESIGN.md "next".  Everything around them
(wrap padding, activation, residual) is expressed so that a fused kernel can slot in per block.
"""
import torch
import torch.nn.functional as F


def _wrap_w(x):
    """Circular padding of one column on each side of W, none on H (CircularPad((1,1,0,0)))."""
    return F.pad(x, (1, 1, 0, 0), mode="circular")


class _Conv3x3Wrap(torch.nn.Conv2d):
    """3x3 conv, zero padding in H (padding=(1,0)), input pre-wrapped in W; no bias
    (src/models/resnet_modified.py:126-129)."""

    def __init__(self, cin, cout, stride=1):
        super().__init__(cin, cout, kernel_size=3, stride=stride, padding=(1, 0), bias=False)


class BasicBlock(torch.nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, activation_fct="relu"):
        super().__init__()
        self.conv1 = _Conv3x3Wrap(inplanes, planes, stride)
        self.conv2 = _Conv3x3Wrap(planes, planes)
        self.downsample = downsample
        self.stride = stride
        self.act = torch.relu if activation_fct == "relu" else torch.tanh

    def forward(self, x):
        out = self.act(self.conv1(_wrap_w(x)))
        out = self.conv2(_wrap_w(out))
        identity = x if self.downsample is None else self.downsample(x)       # 1x1, no padding needed
        return self.act(out + identity)


class ResNetModified(torch.nn.Module):
    def __init__(self, in_channels, num_outputs, use_dropout=False, layers=(2, 2, 2, 2),
                 factor_fewer_resnet_channels=1, activation_fct="relu"):
        super().__init__()
        self.activation_fct = activation_fct
        widths = [int(c / factor_fewer_resnet_channels) for c in (64, 128, 256, 512)]
        strides = [1, (1, 2), (1, 2), (2, 2)]                                # src/models/resnet_modified.py:49-62
        self.dropout_values = torch.nn.Dropout(p=0.2) if use_dropout else torch.nn.Identity()
        self.dropout_channels = torch.nn.Dropout2d(p=0.2) if use_dropout else torch.nn.Identity()
        self.conv1 = torch.nn.Conv2d(in_channels, widths[0], kernel_size=3, stride=(1, 2), padding=(1, 0), bias=False)
        self.maxpool = torch.nn.MaxPool2d(kernel_size=3, stride=(1, 2), padding=(1, 0))
        inplanes = widths[0]
        for li, (planes, stride, blocks) in enumerate(zip(widths, strides, layers), start=1):
            seq = []
            for bi in range(blocks):
                s = stride if bi == 0 else 1
                down = None
                if bi == 0 and (s != 1 or inplanes != planes):
                    down = torch.nn.Sequential(torch.nn.Conv2d(inplanes, planes, kernel_size=1, stride=s, bias=False))
                seq.append(BasicBlock(inplanes, planes, s, down, activation_fct))
                inplanes = planes
            setattr(self, f"layer{li}", torch.nn.Sequential(*seq))
        self.avgpool = torch.nn.AdaptiveAvgPool2d((1, 1))
        self.fc = torch.nn.Linear(widths[3], num_outputs)
        for mod in self.modules():
            if isinstance(mod, torch.nn.Conv2d):                              # :64-66
                torch.nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity=activation_fct)

    def forward(self, x):
        act = torch.relu if self.activation_fct == "relu" else torch.tanh
        x = self.dropout_values(x)
        x = act(self.conv1(_wrap_w(x)))
        x = self.maxpool(_wrap_w(x))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.dropout_channels(self.layer3(x2))
        x4 = self.layer4(x3)
        out = self.dropout_values(self.fc(torch.flatten(self.avgpool(x4), 1)))
        return [x1, x2, x3, x4, out]
