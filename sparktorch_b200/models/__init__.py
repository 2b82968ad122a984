from sparktorch_b200.models.mnist import MnistCNN, mnist_dnn
from sparktorch_b200.models.resnet import ResNet50

__all__ = ["MnistCNN", "mnist_dnn", "ResNet50"]
