"""Structural lowering of bottleneck ResNets (runtime/fused_resnet.py) - CPU only."""
import torch
import torch.nn as nn

from sparktorch_b200.models import ResNet50
from sparktorch_b200.models.resnet import Bottleneck
from sparktorch_b200.runtime.fused_resnet import lower_resnet


def test_our_resnet50_lowers_with_all_parameters_covered():
    m = ResNet50(num_classes=16, image_size=64)
    low = lower_resnet(m, 3 * 64 * 64)
    assert low is not None and len(low.blocks) == 16 and low.pool == (3, 2, 1) and low.in_shape == (3, 64, 64)
    assert sum(b.down is not None for b in low.blocks) == 4          # one projection per stage
    assert [b.u2.conv.stride[0] for b in low.blocks].count(2) == 3   # v1.5: the 3x3 carries the stride
    assert lower_resnet(m, 3 * 32 * 32) is None                      # rows do not match the declared image size
    assert sum(p.numel() for p in ResNet50().parameters()) == 25_557_032


def test_torchvision_style_naming_is_accepted():
    class ResNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 16, 7, 2, 3, bias=False)
            self.bn1 = nn.BatchNorm2d(16)
            self.relu = nn.ReLU(inplace=True)
            self.maxpool = nn.MaxPool2d(3, 2, 1)
            self.layer1 = nn.Sequential(Bottleneck(16, 8, 1))
            self.layer2 = nn.Sequential(Bottleneck(32, 16, 2))
            self.layer3 = nn.Sequential(Bottleneck(64, 32, 2))
            self.layer4 = nn.Sequential(Bottleneck(128, 64, 2))
            self.avgpool = nn.AdaptiveAvgPool2d(1)
            self.fc = nn.Linear(256, 10)

    low = lower_resnet(ResNet(), 3 * 32 * 32)
    assert low is not None and len(low.blocks) == 4


def test_unsupported_variants_fall_back():
    m = ResNet50(num_classes=8, image_size=32)
    m.stem[0] = nn.Conv2d(3, 64, 7, 2, 3, bias=True)                 # conv bias: not a conv-bn unit
    assert lower_resnet(m, 3 * 32 * 32) is None
    m = ResNet50(num_classes=8, image_size=32)
    m.layers[0].bn1 = nn.BatchNorm2d(64, momentum=None)              # cumulative moving average
    assert lower_resnet(m, 3 * 32 * 32) is None
    m = ResNet50(num_classes=8, image_size=32)
    m.extra = nn.Linear(4, 4)                                        # a parameter outside the recognised layers
    assert lower_resnet(m, 3 * 32 * 32) is None

    class NotAResNet(nn.Module):
        def forward(self, x):
            return x

    assert lower_resnet(NotAResNet(), 3 * 32 * 32) is None
