"""MNIST model family used by the examples and benchmarks.

Architectures are the ones the reference's examples train (examples/
simple_dnn.py:21-27: 784→256→256→10 MLP; examples/cnn_network.py:6-24:
conv5x5(1→16)→ReLU→conv3x3(16→32)→maxpool2→dropout2d(.25)→fc(3872→10))."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def mnist_dnn(hidden: int = 256, n_classes: int = 10, in_features: int = 784) -> nn.Sequential:
    return nn.Sequential(
        nn.Linear(in_features, hidden), nn.ReLU(),
        nn.Linear(hidden, hidden), nn.ReLU(),
        nn.Linear(hidden, n_classes),
    )


class MnistCNN(nn.Module):
    def __init__(self, n_classes: int = 10, p_drop: float = 0.25):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 16, kernel_size=5)
        self.conv2 = nn.Conv2d(16, 32, kernel_size=3)
        self.drop = nn.Dropout2d(p_drop)
        self.fc = nn.Linear(32 * 11 * 11, n_classes)

    def forward(self, x):
        x = x.view(-1, 1, 28, 28)
        x = F.relu(self.conv1(x))
        x = F.max_pool2d(self.conv2(x), 2)
        x = self.drop(x)
        return self.fc(torch.flatten(x, 1))
