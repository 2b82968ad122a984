"""Small networks used by the API scenario tests (shapes follow the reference's
test nets, sparktorch/tests/simple_net.py: 10→20→1 regressor, 10→5→2→5→10
auto-encoder, 10→20→2 log-softmax classifier, width-parameterised regressor)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Regressor(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(10, 20)
        self.fc2 = nn.Linear(20, 1)

    def forward(self, x):
        return self.fc2(F.relu(self.fc1(x)))


class AutoEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.Sequential(nn.Linear(10, 5), nn.ReLU(), nn.Linear(5, 2), nn.ReLU())
        self.dec = nn.Sequential(nn.Linear(2, 5), nn.ReLU(), nn.Linear(5, 10))

    def forward(self, x):
        return self.dec(self.enc(x))


class Classifier(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(10, 20)
        self.fc2 = nn.Linear(20, 2)

    def forward(self, x):
        return F.log_softmax(self.fc2(F.relu(self.fc1(x))), dim=1)


class WidthRegressor(nn.Module):
    def __init__(self, param):
        super().__init__()
        self.fc1 = nn.Linear(10, param)
        self.fc2 = nn.Linear(param, 1)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


class ResNet(nn.Module):
    """A small bottleneck ResNet (projection blocks with stride 1 and 2 plus one identity block): every code path of
    the fused residual-network plan, shallow enough for tight tolerances."""

    def __init__(self, num_classes=16, image_size=32, widths=(16, 32, 64, 128), stem=16):
        super().__init__()
        from sparktorch_b200.models.resnet import Bottleneck

        self.image_size = image_size
        self.stem = nn.Sequential(nn.Conv2d(3, stem, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(stem), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, stride=2, padding=1))
        layers, cin = [], stem
        for i, w in enumerate(widths):
            layers.append(Bottleneck(cin, w, 1 if i == 0 else 2))
            cin = w * 4
            if i == 1:
                layers.append(Bottleneck(cin, w, 1))
        self.layers = nn.Sequential(*layers)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = x.view(-1, 3, self.image_size, self.image_size)
        return self.fc(torch.flatten(self.pool(self.layers(self.stem(x))), 1))
