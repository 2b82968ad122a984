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
