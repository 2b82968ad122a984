"""Fused residual-network plan (ResNet-50 style bottleneck nets with BatchNorm).

The config is named in BASELINE.json / SURVEY.md §2.7 (G6 at ResNet-50 shapes, G13 BatchNorm): the
reference itself would run it through cuDNN convolutions and ATen batch-norm kernels
(distributed.py:150,176 for any ``nn.Module``).  Here every convolution is a tcgen05 GEMM over
channels-last bf16 activations ``[B*H*W, C]`` and everything around it is a handful of vectorised kernels
(csrc/norm.cu):

    conv 1x1           Y = X W^T                                   (activations ARE the GEMM operand)
    conv kxk / strided Y = im2col_cl(X) W^T,  W kept [Cout, kh, kw, Cin] (channels_last parameter layout, so the
                       bf16 copy K-AR publishes and the fp32 gradient the wgrad GEMM writes need no permutation)
    BatchNorm (train)  bn_stats -> bn_finalize (+running stats) -> bn_apply (+residual, +ReLU)
    backward           bn_backward(reduce) writes dgamma/dbeta straight into the K-AR gradient arena and fuses the
                       ReLU mask and the sum of the two gradients meeting at a residual join; bn_backward(apply)
                       produces dY (+ the masked gradient that continues along the skip connection)
                       wgrad  dW = dY^T A   (split-K red.add into the fp32 arena)
                       dgrad  1x1: dX = dY W;  kxk: im2col_cl(zero-dilated, padded dY) x flipped taps
    head               global average pool, Linear (GEMM + bias), softmax cross-entropy kernel
    update             one K-AR launch: NVLink reduce-scatter + SGD/Adam + publish of the 25.6 M parameters

Interface-compatible with FusedMLP / FusedCNN so ``fused_mlp.try_fused_fit`` captures it in one CUDA graph.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import List, Optional

import torch
import torch.nn as nn

from sparktorch_b200.ops import native
from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.fused_mlp import _pad


@dataclasses.dataclass
class Unit:
    """conv -> batch-norm (-> +residual) (-> ReLU)"""
    conv: nn.Conv2d
    bn: nn.BatchNorm2d
    relu: bool
    g: dict = None           # geometry, filled by the plan


@dataclasses.dataclass
class Block:
    u1: Unit
    u2: Unit
    u3: Unit
    down: Optional[Unit]


@dataclasses.dataclass
class LoweredResNet:
    in_shape: tuple
    stem: Unit
    pool: tuple              # (k, s, p) of the stem max-pool
    blocks: List[Block]
    fc: nn.Linear
    final_log_softmax: bool = False


def _one(v):
    return v[0] if isinstance(v, (tuple, list)) else v


def _conv_ok(c) -> bool:
    if not isinstance(c, nn.Conv2d) or c.bias is not None or c.groups != 1 or c.padding_mode != "zeros":
        return False
    if isinstance(c.padding, str):
        return False
    for t in (c.kernel_size, c.stride, c.padding, c.dilation):
        if isinstance(t, (tuple, list)) and t[0] != t[1]:
            return False
    return _one(c.dilation) == 1


def _bn_ok(b) -> bool:
    return isinstance(b, nn.BatchNorm2d) and b.affine and b.track_running_stats and b.momentum is not None


def lower_resnet(model: nn.Module, in_features: int) -> Optional[LoweredResNet]:
    """Structural match of a bottleneck ResNet: stem (conv-bn-relu-maxpool), blocks with conv1/bn1/conv2/bn2/
    conv3/bn3 and an optional projection (``down`` / ``downsample``), adaptive average pool, fc."""
    if type(model).__name__ not in ("ResNet50", "ResNet"):
        return None
    stem = getattr(model, "stem", None)
    if isinstance(stem, nn.Sequential) and len(stem) == 4:
        conv, bn, act, mp = stem
    else:
        conv, bn, act, mp = (getattr(model, n, None) for n in ("conv1", "bn1", "relu", "maxpool"))
    if not (_conv_ok(conv) and _bn_ok(bn) and isinstance(act, nn.ReLU) and isinstance(mp, nn.MaxPool2d)):
        return None
    if mp.dilation != 1 or mp.ceil_mode or mp.return_indices:
        return None
    blocks_src = []
    if isinstance(getattr(model, "layers", None), nn.Sequential):
        blocks_src = list(model.layers)
    else:
        for n in ("layer1", "layer2", "layer3", "layer4"):
            layer = getattr(model, n, None)
            if not isinstance(layer, nn.Sequential):
                return None
            blocks_src += list(layer)
    blocks = []
    for b in blocks_src:
        if type(b).__name__ != "Bottleneck":
            return None
        cs = [getattr(b, f"conv{i}", None) for i in (1, 2, 3)]
        bs = [getattr(b, f"bn{i}", None) for i in (1, 2, 3)]
        if not all(_conv_ok(c) for c in cs) or not all(_bn_ok(x) for x in bs):
            return None
        down = getattr(b, "down", getattr(b, "downsample", None))
        du = None
        if down is not None:
            if not (isinstance(down, nn.Sequential) and len(down) == 2 and _conv_ok(down[0]) and _bn_ok(down[1])):
                return None
            du = Unit(down[0], down[1], False)
        blocks.append(Block(Unit(cs[0], bs[0], True), Unit(cs[1], bs[1], True), Unit(cs[2], bs[2], True), du))
    pool = getattr(model, "pool", getattr(model, "avgpool", None))
    fc = getattr(model, "fc", None)
    if not isinstance(pool, nn.AdaptiveAvgPool2d) or _one(pool.output_size) != 1 or not isinstance(fc, nn.Linear):
        return None
    if not blocks:
        return None
    cin = conv.in_channels
    side = int(round(math.sqrt(in_features / cin)))
    if side * side * cin != in_features or getattr(model, "image_size", side) != side:
        return None
    # every parameter must belong to a recognised layer; channel counts must be vector friendly
    covered = {id(conv.weight), id(bn.weight), id(bn.bias), id(fc.weight)}
    if fc.bias is not None:
        covered.add(id(fc.bias))
    for b in blocks:
        for u in (b.u1, b.u2, b.u3) + ((b.down,) if b.down else ()):
            covered |= {id(u.conv.weight), id(u.bn.weight), id(u.bn.bias)}
            if u.conv.in_channels % 8 or u.conv.out_channels % 8 or u.conv.out_channels > 2048:
                return None
    params = [p for p in model.parameters() if p.requires_grad]
    if len(params) != len(covered) or any(id(p) not in covered for p in params):
        return None
    if conv.out_channels % 8:
        return None
    return LoweredResNet((cin, side, side), Unit(conv, bn, True), (_one(mp.kernel_size), _one(mp.stride or mp.kernel_size),
                                                                    _one(mp.padding)), blocks, fc)


class FusedResNet:
    """Buffers + enqueue functions; same interface as FusedMLP/FusedCNN (forward / backward / update over ``buf``)."""

    fuse_loss = False          # the 1000-way loss is its own kernel, batches are gathered (no in-GEMM windows)

    def __init__(self, model: nn.Module, low: LoweredResNet, kind: str, spec, world: int, rank: int, device: torch.device,
                 batch: int, val_rows: int, patience: int, max_iters: int, seed: int = 0, backend_factory=None,
                 inference: bool = False, ops=None, act_dtype: torch.dtype = torch.bfloat16):
        """``backend_factory(numel, wbf16_elems)`` / ``ops`` / ``act_dtype`` exist for the CPU emulation of the plan
        (runtime/emulation.py + ops/reference.py); production uses FusedUpdate, the native extension and bf16."""
        assert kind == "ce"
        self.eval_running = inference        # inference: BatchNorm uses the running statistics (module.eval())
        self.act_dtype = act_dtype
        self.C = ops if ops is not None else native.load(required=True)
        self.device, self.world, self.rank = device, world, rank
        self.kind, self.low, self.lowered = kind, low, low
        self.B, self.nv = batch, val_rows
        C0, H0, W0 = low.in_shape
        self.d0 = C0 * H0 * W0
        self.ncls = low.fc.out_features
        # ---- geometry ---------------------------------------------------------------------------------
        def geo(u: Unit, cin, h, w):
            k, s, p = _one(u.conv.kernel_size), _one(u.conv.stride), _one(u.conv.padding)
            assert u.conv.in_channels == cin
            oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
            u.g = dict(Cin=cin, H=h, W=w, Cout=u.conv.out_channels, k=k, s=s, p=p, OH=oh, OW=ow, K=cin * k * k,
                       direct=(k == 1 and s == 1))
            return u.conv.out_channels, oh, ow

        c, h, w = geo(low.stem, C0, H0, W0)
        pk, ps, pp = low.pool
        self.pool_in = (h, w)
        h, w = (h + 2 * pp - pk) // ps + 1, (w + 2 * pp - pk) // ps + 1
        self.pool_out = (h, w)
        for b in low.blocks:
            cin, hin, win = c, h, w
            c1, h1, w1 = geo(b.u1, cin, hin, win)
            c2, h2, w2 = geo(b.u2, c1, h1, w1)
            c, h, w = geo(b.u3, c2, h2, w2)
            if b.down is not None:
                cd, hd, wd = geo(b.down, cin, hin, win)
                assert (cd, hd, wd) == (c, h, w)
            else:
                assert (cin, hin, win) == (c, h, w)
        self.feat_c, self.feat_hw = c, h * w
        assert low.fc.in_features == c
        self.units: List[Unit] = [low.stem]
        for b in low.blocks:
            self.units += [b.u1, b.u2, b.u3] + ([b.down] if b.down else [])
        # ---- arenas -----------------------------------------------------------------------------------
        numel = FlatArena.measure(model)
        w_off, off = {}, 0
        for m, rows, cols in [(u.conv, u.g["Cout"], u.g["K"]) for u in self.units] + [(low.fc, self.ncls, c)]:
            w_off[id(m)] = (off, rows, cols)
            off += _pad(rows * _pad(cols), 64)
        if backend_factory is not None:
            self.fu = backend_factory(numel, off)
        else:
            from sparktorch_b200.parallel.p2p import FusedUpdate

            self.fu = FusedUpdate(numel, spec, world, rank, device, patience, max_iters, wbf16_elems=off)
        for bn in [u.bn for u in self.units]:                       # buffers live on the device, kernels update them
            bn.running_mean.data = bn.running_mean.data.to(device=device, dtype=torch.float32).contiguous()
            bn.running_var.data = bn.running_var.data.to(device=device, dtype=torch.float32).contiguous()
            bn.num_batches_tracked.data = bn.num_batches_tracked.data.to(device)
        self.arena = FlatArena(model, device, grad_buffer=self.fu.grad, param_buffer=self.fu.param)
        pos = {id(p): o for p, o in zip(self.arena.params, self.arena.offsets)}
        # kxk weights (beyond the stem) are kept physically [Cout, kh, kw, Cin]: the module sees a channels_last view
        with torch.no_grad():
            for u in self.units[1:]:
                g = u.g
                if g["k"] > 1:
                    p_, o, n = u.conv.weight, pos[id(u.conv.weight)], u.conv.weight.numel()
                    vals = p_.detach().clone()
                    phys = self.fu.param[o: o + n].view(g["Cout"], g["k"], g["k"], g["Cin"])
                    phys.copy_(vals.permute(0, 2, 3, 1))
                    p_.data = phys.permute(0, 3, 1, 2)
                    p_.grad = self.fu.grad[o: o + n].view(g["Cout"], g["k"], g["k"], g["Cin"]).permute(0, 3, 1, 2)
        seg_of = {}
        self.Wb, self.gW = {}, {}
        for m in [u.conv for u in self.units] + [low.fc]:
            o, rows, cols = w_off[id(m)]
            ldw = _pad(cols)
            wo = pos[id(m.weight)]
            seg_of[wo] = [wo, rows, cols, o, ldw, -1, 0, 1]
            self.Wb[id(m)] = self.fu.wbf16[o: o + rows * ldw].view(rows, ldw)[:, :cols]
            self.gW[id(m)] = self.fu.grad[wo: wo + rows * cols].view(rows, cols)
        self.vec = {}                                   # fp32 vectors in the arenas: id(param) -> (param view, grad view)
        vec_params = [low.fc.bias] if low.fc.bias is not None else []
        for u in self.units:
            vec_params += [u.bn.weight, u.bn.bias]
        for p_ in vec_params:
            o, n = pos[id(p_)], p_.numel()
            seg_of[o] = [o, 1, n, -1, 0, -1, 0, 1]
            self.vec[id(p_)] = (self.fu.param[o: o + n], self.fu.grad[o: o + n])
        self.fu.build([seg_of[o] for o in sorted(seg_of)], zero_grads=True)
        self.loss_slot = self.fu.grad[numel: numel + 1]
        self.train_loss_slot = self.fu.grad[numel + 2: numel + 3]
        self.Wflip = {}
        for u in self.units[1:]:
            g = u.g
            if g["k"] > 1:
                self.Wflip[id(u)] = torch.zeros(g["Cin"], g["k"] * g["k"] * g["Cout"], dtype=self.fu.wbf16.dtype, device=device)
        self.window = None
        self.bn_in_epilogue = os.environ.get("SPARKTORCH_BN_EPILOGUE", "0") == "1"
        self.train = self._alloc(batch, backward=not inference)
        self.val = self._alloc(val_rows, backward=False) if val_rows > 0 else None
        self.n_launch = 0
        self._bn_saved = None

    # ------------------------------------------------------------------------------------------------
    def _alloc(self, rows: int, backward: bool):
        dev, bf = self.device, self.act_dtype
        z = lambda r, c_: torch.zeros(r, c_, dtype=bf, device=dev)          # noqa: E731
        buf = {"rows": rows, "acts": [torch.zeros(rows, _pad(self.d0), dtype=bf, device=dev)[:, : self.d0]], "u": {}}
        # batch-norm accumulators of all units in one slab: cleared by a single memset at the top of forward
        buf["sums_all"] = torch.zeros(sum(2 * u.g["Cout"] for u in self.units), dtype=torch.float32, device=dev)
        so = 0
        for i, u in enumerate(self.units):
            g = u.g
            M = rows * g["OH"] * g["OW"]
            ub = {"M": M, "Y": z(M, g["Cout"]), "Z": z(M, g["Cout"]),
                  "sums": buf["sums_all"][so: so + 2 * g["Cout"]],
                  "stats": torch.zeros(2 * g["Cout"], dtype=torch.float32, device=dev)}
            so += 2 * g["Cout"]
            if not g["direct"]:
                ub["col"] = z(M, _pad(g["K"]))[:, : g["K"]]
            if backward:
                ub["dY"] = z(M, g["Cout"])
                if i > 0:
                    Min = rows * g["H"] * g["W"]
                    ub["gX"] = z(Min, g["Cin"])                   # gradient wrt this unit's input
                    if g["k"] > 1:
                        ub["colT"] = z(Min, g["k"] * g["k"] * g["Cout"])
                    elif not g["direct"]:
                        ub["gXc"] = z(M, g["Cin"])                # compact dgrad of a strided 1x1 before zero insertion
            buf["u"][id(u)] = ub
        ph, pw = self.pool_out
        c_stem = self.low.stem.g["Cout"]
        buf["P"] = z(rows * ph * pw, c_stem)
        buf["P_idx"] = torch.zeros(rows * ph * pw * c_stem, dtype=torch.uint8, device=dev)
        buf["feat"] = z(rows, self.feat_c)
        buf["logits_full"] = torch.zeros(rows, _pad(self.ncls, 16), dtype=torch.float32, device=dev)
        buf["logits"] = buf["logits_full"][:, : self.ncls]
        buf["tgt"] = torch.zeros(rows, dtype=torch.int32, device=dev)
        if backward:
            buf["dlogits"] = z(rows, _pad(self.ncls))[:, : self.ncls]
            buf["dfeat"] = z(rows, self.feat_c)
            buf["dZs"] = z(rows * self.pool_in[0] * self.pool_in[1], c_stem)
            buf["G"] = {id(b): z(buf["u"][id(b.u3)]["M"], b.u3.g["Cout"]) for b in self.low.blocks}
            last = self.low.blocks[-1]
            buf["gLast"] = z(buf["u"][id(last.u3)]["M"], last.u3.g["Cout"])
        return buf

    def _gemm(self, *a, **kw):
        self.C.gemm_bf16(*a, **kw)
        self.n_launch += 1

    @staticmethod
    def _split_k(M: int, N: int, K: int) -> int:
        tiles = ((M + 127) // 128) * ((N + 63) // 64)
        kb = (K + 63) // 64
        if kb < 8 or tiles >= 148:
            return 1
        return max(1, min(kb // 4, 148 * 2 // tiles, 128))

    # ------------------------------------------------------------------------------------------------
    def _unit_fwd(self, buf, u: Unit, x: torch.Tensor, train: bool, residual=None, stem_src=None) -> torch.Tensor:
        C, g, ub, rows = self.C, u.g, buf["u"][id(u)], buf["rows"]
        M = ub["M"]
        if stem_src is not None:          # Spark rows: (c, h, w) order, generic-stride im2col with (c, i, j) columns
            C.im2col_nhwc(stem_src, rows, g["H"], g["W"], g["Cin"], stem_src.stride(0), g["W"], 1, g["H"] * g["W"], g["k"], g["k"],
                          g["s"], g["p"], ub["col"])
            self.n_launch += 1
            a = ub["col"]
        elif g["direct"]:
            a = x
        else:
            C.im2col_cl(x, rows, g["H"], g["W"], g["Cin"], g["k"], g["k"], g["s"], g["p"], 1, g["OH"], g["OW"], ub["col"])
            self.n_launch += 1
            a = ub["col"]
        bn = u.bn
        gamma, beta = self.vec[id(bn.weight)][0], self.vec[id(bn.bias)][0]
        Cout = g["Cout"]
        if self.eval_running and not train:
            self._gemm(a, False, self.Wb[id(u.conv)], False, M, Cout, g["K"], out_bf16=ub["Y"])
            C.bn_finalize(ub["sums"], M, Cout, bn.eps, bn.momentum, bn.running_mean, bn.running_var, ub["stats"],
                          from_running=True, update_running=False)
            C.bn_apply(ub["Y"], M, Cout, ub["stats"], gamma, beta, ub["Z"], relu=u.relu, residual=residual)
            self.n_launch += 2
            return ub["Z"]
        if self.bn_in_epilogue:
            # batch statistics are column sums of the GEMM output: reduced in its epilogue, finalised inside bn_apply
            self._gemm(a, False, self.Wb[id(u.conv)], False, M, Cout, g["K"], out_bf16=ub["Y"], colsum=ub["sums"][:Cout],
                       colsumsq=ub["sums"][Cout:])
            C.bn_apply(ub["Y"], M, Cout, ub["stats"], gamma, beta, ub["Z"], relu=u.relu, residual=residual, sums=ub["sums"],
                       eps=bn.eps, momentum=bn.momentum, running_mean=bn.running_mean if train else None,
                       running_var=bn.running_var if train else None)
            self.n_launch += 1
            return ub["Z"]
        self._gemm(a, False, self.Wb[id(u.conv)], False, M, Cout, g["K"], out_bf16=ub["Y"])
        C.bn_stats(ub["Y"], M, Cout, ub["sums"])
        C.bn_finalize(ub["sums"], M, Cout, bn.eps, bn.momentum, bn.running_mean, bn.running_var, ub["stats"],
                      from_running=False, update_running=train)
        C.bn_apply(ub["Y"], M, Cout, ub["stats"], gamma, beta, ub["Z"], relu=u.relu, residual=residual)
        self.n_launch += 3
        return ub["Z"]

    def forward(self, buf, slot=None, train: bool = False, windowed: bool = False) -> None:
        C, rows, low = self.C, buf["rows"], self.low
        if self.bn_in_epilogue:
            buf["sums_all"].zero_()
            self.n_launch += 1
        zs = self._unit_fwd(buf, low.stem, None, train, stem_src=buf["acts"][0])
        sg = low.stem.g
        pk, ps, pp = low.pool
        C.maxpool_cl(False, zs, buf["P"], buf["P_idx"], rows, sg["OH"], sg["OW"], sg["Cout"], pk, ps, pp)
        self.n_launch += 1
        cur = buf["P"]
        for b in low.blocks:
            res = self._unit_fwd(buf, b.down, cur, train) if b.down is not None else cur
            z1 = self._unit_fwd(buf, b.u1, cur, train)
            z2 = self._unit_fwd(buf, b.u2, z1, train)
            cur = self._unit_fwd(buf, b.u3, z2, train, residual=res)
        C.avgpool_cl(False, cur, buf["feat"], rows, self.feat_hw, self.feat_c)
        self.n_launch += 1
        fc = low.fc
        bias = self.vec[id(fc.bias)][0] if fc.bias is not None else None
        self._gemm(buf["feat"], False, self.Wb[id(fc)], False, rows, self.ncls, self.feat_c, bias=bias, out_f32=buf["logits"])
        if slot is not None:
            dz = buf["dlogits"] if train else None
            db = self.vec[id(fc.bias)][1] if (train and fc.bias is not None) else None
            C.softmax_ce(buf["logits"], buf["tgt"], self.ncls, slot, dz, None, db, _pad(self.ncls))
            self.n_launch += 1

    # ------------------------------------------------------------------------------------------------
    def _unit_bwd(self, buf, u: Unit, x_in: Optional[torch.Tensor], dz_a, dz_b=None, g_out=None, need_dx: bool = True):
        """Backward of conv-bn(-relu) given the gradient(s) wrt its output; returns the gradient wrt its input."""
        C, g, ub, rows = self.C, u.g, buf["u"][id(u)], buf["rows"]
        M, bn = ub["M"], u.bn
        gamma = self.vec[id(bn.weight)][0]
        dgamma, dbeta = self.vec[id(bn.weight)][1], self.vec[id(bn.bias)][1]
        C.bn_backward(True, ub["Y"], M, g["Cout"], ub["stats"], gamma, ub["Z"], u.relu, dz_a, dz_b=dz_b, dgamma=dgamma, dbeta=dbeta)
        C.bn_backward(False, ub["Y"], M, g["Cout"], ub["stats"], gamma, ub["Z"], u.relu, dz_a, dz_b=dz_b, dgamma=dgamma, dbeta=dbeta,
                      dy=ub["dY"], g_out=g_out)
        self.n_launch += 2
        dY = ub["dY"]
        a = x_in if g["direct"] else ub["col"]
        sk = self._split_k(g["Cout"], g["K"], M)
        self._gemm(dY, True, a, True, g["Cout"], g["K"], M, out_f32=self.gW[id(u.conv)], atomic_f32=sk > 1, split_k=sk)
        if not need_dx:
            return None
        Wb = self.Wb[id(u.conv)]
        if g["direct"]:
            self._gemm(dY, False, Wb, True, M, g["Cin"], g["Cout"], out_bf16=ub["gX"])
        elif g["k"] == 1:
            self._gemm(dY, False, Wb, True, M, g["Cin"], g["Cout"], out_bf16=ub["gXc"])
            C.im2col_cl(ub["gXc"], rows, g["OH"], g["OW"], g["Cin"], 1, 1, 1, 0, g["s"], g["H"], g["W"], ub["gX"])
            self.n_launch += 1
        else:
            kk = g["k"] * g["k"]
            C.flip_conv_weight_cl(Wb, g["Cout"], g["Cin"], g["k"], g["k"], self.Wflip[id(u)])
            C.im2col_cl(dY, rows, g["OH"], g["OW"], g["Cout"], g["k"], g["k"], 1, g["k"] - 1 - g["p"], g["s"], g["H"], g["W"],
                        ub["colT"])
            self.n_launch += 2
            self._gemm(ub["colT"], False, self.Wflip[id(u)], False, rows * g["H"] * g["W"], g["Cin"], kk * g["Cout"],
                       out_bf16=ub["gX"])
        return ub["gX"]

    def backward(self, buf, windowed: bool = False) -> None:
        C, rows, low = self.C, buf["rows"], self.low
        fc = low.fc
        sk = self._split_k(self.ncls, self.feat_c, rows)
        self._gemm(buf["dlogits"], True, buf["feat"], True, self.ncls, self.feat_c, rows, out_f32=self.gW[id(fc)], atomic_f32=sk > 1,
                   split_k=sk)
        self._gemm(buf["dlogits"], False, self.Wb[id(fc)], True, rows, self.feat_c, self.ncls, out_bf16=buf["dfeat"])
        C.avgpool_cl(True, buf["gLast"], buf["dfeat"], rows, self.feat_hw, self.feat_c)
        self.n_launch += 1
        ga, gb = buf["gLast"], None
        for bi in range(len(low.blocks) - 1, -1, -1):
            b = low.blocks[bi]
            x_in = buf["u"][id(low.blocks[bi - 1].u3)]["Z"] if bi > 0 else buf["P"]
            z1, z2 = buf["u"][id(b.u1)]["Z"], buf["u"][id(b.u2)]["Z"]
            G = buf["G"][id(b)]
            g2 = self._unit_bwd(buf, b.u3, z2, ga, dz_b=gb, g_out=G)
            g1 = self._unit_bwd(buf, b.u2, z1, g2)
            ga = self._unit_bwd(buf, b.u1, x_in, g1)
            gb = self._unit_bwd(buf, b.down, x_in, G) if b.down is not None else G
        sg = low.stem.g
        pk, ps, pp = low.pool
        C.maxpool_cl(True, buf["dZs"], ga, buf["P_idx"], rows, sg["OH"], sg["OW"], sg["Cout"], pk, ps, pp, dy_b=gb)
        self.n_launch += 1
        self._unit_bwd(buf, low.stem, None, buf["dZs"], need_dx=False)

    def update(self) -> None:
        self.fu.updater.step()
        self.n_launch += 1

    # warm-up launches must not leave a trace in the running statistics
    def save_extra(self):
        self._bn_saved = [(u.bn.running_mean.clone(), u.bn.running_var.clone()) for u in self.units]

    def restore_extra(self):
        for u, (rm, rv) in zip(self.units, self._bn_saved):
            u.bn.running_mean.copy_(rm)
            u.bn.running_var.copy_(rv)
        self._bn_saved = None

    def finish(self, iters_run: int) -> None:
        for u in self.units:
            u.bn.num_batches_tracked += int(iters_run)
