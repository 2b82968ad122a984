"""Fused residual-network plan (csrc/norm.cu + tcgen05 GEMMs) against PyTorch fp32 autograd."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _small_resnet(num_classes=16, image=64):
    from sparktorch_b200.models import ResNet50

    torch.manual_seed(0)
    return ResNet50(num_classes=num_classes, image_size=image)


class ResNet(nn.Module):
    """One bottleneck per stage, narrow widths: shallow enough for tight bf16-vs-fp32 tolerances."""

    def __init__(self, num_classes=16, image_size=32, widths=(16, 32, 64, 128)):
        super().__init__()
        from sparktorch_b200.models.resnet import Bottleneck

        self.image_size = image_size
        self.stem = nn.Sequential(nn.Conv2d(3, 16, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(16), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, stride=2, padding=1))
        layers, cin = [], 16
        for i, w in enumerate(widths):
            layers.append(Bottleneck(cin, w, 1 if i == 0 else 2))
            cin = w * 4
            if i == 1:
                layers.append(Bottleneck(cin, w, 1))      # an identity-skip block as well
        self.layers = nn.Sequential(*layers)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = x.view(-1, 3, self.image_size, self.image_size)
        return self.fc(torch.flatten(self.pool(self.layers(self.stem(x))), 1))


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("depth", ["shallow", "resnet50"])
def test_forward_backward_match_autograd(depth):
    batch = 16 if depth == "shallow" else 8
    from sparktorch_b200.runtime.fused_resnet import FusedResNet, lower_resnet
    from sparktorch_b200.runtime.optim import parse_optimizer

    dev = torch.device("cuda:0")
    image, ncls = 64, 16
    if depth == "shallow":
        torch.manual_seed(0)
        model = ResNet(ncls, image)
    else:
        model = _small_resnet(ncls, image)
    # Random-init batch-norm nets amplify rounding noise layer by layer (tiny batch statistics), so the yardstick is
    # what torch's own bf16 autocast run loses against fp32 on the same net: we must not be worse than that by more
    # than a margin, parameter by parameter.
    tol_logits, tol_fc = (0.995, 0.995) if depth == "shallow" else (0.93, 0.95)
    ref = copy.deepcopy(model).to(dev).float().train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3 * image * image, generator=g).to(torch.bfloat16)
    y = torch.randint(0, ncls, (batch,), generator=g)
    low = lower_resnet(model, 3 * image * image)
    spec = parse_optimizer(torch.optim.SGD, {"lr": 0.1, "momentum": 0.9})
    model = model.to(dev)
    plan = FusedResNet(model, low, "ce", spec, 1, 0, dev, batch, 0, 0, 4)
    plan.fu.updater.publish()
    plan.train["acts"][0].copy_(x.to(dev))
    plan.train["tgt"].copy_(y.to(dev).to(torch.int32))
    plan.forward(plan.train, slot=plan.loss_slot, train=True)
    plan.backward(plan.train)
    torch.cuda.synchronize()
    loss = float(plan.loss_slot[0])
    # reference
    out = ref(x.to(dev).float())
    ref_loss = nn.functional.cross_entropy(out, y.to(dev))
    ref_loss.backward()
    assert abs(loss - float(ref_loss.detach())) < 0.05 * max(1.0, abs(float(ref_loss.detach()))), (loss, float(ref_loss.detach()))
    logits = plan.train["logits"].float()
    assert _cos(logits, out.detach()) > tol_logits, _cos(logits, out.detach())
    got = dict(model.named_parameters())
    amp = copy.deepcopy(ref)
    amp.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        nn.functional.cross_entropy(amp(x.to(dev).float()).float(), y.to(dev)).backward()
    amp_grads = {n: p.grad for n, p in amp.named_parameters()}
    worst, bad = 1.0, []
    for name, p in ref.named_parameters():
        c = _cos(got[name].grad, p.grad)
        ratio = float(got[name].grad.norm() / (p.grad.norm() + 1e-30))
        tol_grad = min(0.97, _cos(amp_grads[name], p.grad) - 0.10)
        if depth != "shallow":
            # 50 random-init layers at batch 8 are chaotic (torch's own bf16 autocast run correlates ~0.15 with fp32 in
            # the early layers): directions are only checked near the loss, magnitudes everywhere
            tol_grad = 0.9 if name.startswith("fc.") else -1.0
            assert 0.5 < ratio < 2.0, (name, ratio)
        print(f"  {name:32s} cos {c:.4f} (autocast {_cos(amp_grads[name], p.grad):.4f})  |ours|/|ref| {ratio:.3f}")
        worst = min(worst, c)
        if c <= tol_grad:
            bad.append((name, round(c, 4)))
    assert not bad, bad
    # running statistics follow nn.BatchNorm2d
    for (n1, b1), (n2, b2) in zip(model.named_buffers(), ref.named_buffers()):
        if n1.endswith("running_mean") or n1.endswith("running_var"):
            torch.testing.assert_close(b1.float(), b2.float(), rtol=5e-2, atol=5e-2)
    # the big tensors should agree closely
    assert _cos(got["fc.weight"].grad, dict(ref.named_parameters())["fc.weight"].grad) > tol_fc
    print("worst gradient cosine", worst)


def test_fit_small_resnet_through_engine():
    """End to end through train_partition: a small bottleneck net memorises 64 random samples (the CPU emulation of the
    same plan reaches ~0 loss by iteration 12 in fp32 and in bf16 for several seeds, so the margin below is wide)."""
    from sparktorch_b200.runtime.engine import TrainConfig, train_partition

    dev = torch.device("cuda:0")
    image, ncls, n = 32, 8, 64
    torch.manual_seed(0)
    model = ResNet(ncls, image)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 3 * image * image)).astype(np.float32)
    Y = rng.integers(0, ncls, size=(n, 1)).astype(np.float32)
    cfg = TrainConfig(iters=40, mini_batch=32, device="cuda", seed=3, engine="fused", sampling="sweep")
    res = train_partition(0, 1, model, nn.CrossEntropyLoss(), torch.optim.SGD, {"lr": 0.05, "momentum": 0.9}, X, Y, cfg, dev)
    h = res.history
    assert h["engine"] == "fused_resnet", h["engine"]
    tl = h["train_loss"]
    assert len(tl) == 40 and all(np.isfinite(tl))
    assert np.mean(tl[-6:]) < 0.3 * np.mean(tl[:4]), tl
    sd = res.state_dict
    assert sd["stem.1.num_batches_tracked"].item() == 40
    torch.manual_seed(0)
    m2 = ResNet(ncls, image)
    m2.load_state_dict(sd)


def test_fused_resnet_predictor_uses_running_stats():
    from sparktorch_b200.runtime.predict import BatchPredictor

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = ResNet(16, 32)
    with torch.no_grad():                     # non-trivial running statistics / affine parameters
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    ref = copy.deepcopy(net).to(dev).eval()
    X = np.random.default_rng(1).normal(size=(100, 3 * 32 * 32)).astype(np.float32)
    bp = BatchPredictor(net.eval(), dev, batch_size=48)
    out = bp.predict_block(X)
    assert bp.fused is not None and bp.fused.engine == "fused_resnet"
    with torch.no_grad():
        want = ref(torch.from_numpy(X).to(dev)).cpu().numpy()
    assert _cos(torch.from_numpy(out), torch.from_numpy(want)) > 0.995
    assert (out.argmax(1) == want.argmax(1)).mean() > 0.9
