from sparktorch_b200.models.mnist import MnistCNN, mnist_dnn

__all__ = ["MnistCNN", "mnist_dnn"]
