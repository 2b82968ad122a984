"""ResNet-50 (He et al.) written for this package – Spark rows are flat vectors, so the
module starts with the reshape to (3, H, W).  Architecture: 7x7/2 stem, 3x3/2 max-pool,
bottleneck stages [3, 4, 6, 3] with widths 64/128/256/512 (x4 expansion), global average
pool, fc.  Parameter count 25 557 032 for 1000 classes."""
import torch
import torch.nn as nn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin: int, width: int, stride: int = 1):
        super().__init__()
        cout = width * self.expansion
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet50(nn.Module):
    def __init__(self, num_classes: int = 1000, image_size: int = 224):
        super().__init__()
        self.image_size = image_size
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, stride=2, padding=1))
        layers, cin = [], 64
        for width, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for b in range(blocks):
                layers.append(Bottleneck(cin, width, stride if b == 0 else 1))
                cin = width * Bottleneck.expansion
        self.layers = nn.Sequential(*layers)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = x.view(-1, 3, self.image_size, self.image_size).contiguous(memory_format=torch.channels_last)
        x = self.layers(self.stem(x))
        return self.fc(torch.flatten(self.pool(x), 1))
