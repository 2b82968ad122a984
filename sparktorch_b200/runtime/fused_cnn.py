"""Fused CNN plan: conv stacks (+ReLU, +max-pool, +dropout2d) followed by a
Linear head, e.g. the reference's MNIST CNN (examples/cnn_network.py:6-24).

Convolutions run on the same tcgen05 GEMM as the Linear layers (implicit GEMM over
channels-last activations, see csrc/conv.cu for the operand builders):

    forward   im2col -> GEMM(+bias,+ReLU)            per conv;  maxpool(+dropout2d) writes torch's flatten order
    backward  wgrad  GEMM(dY^T, col) split-K into the fp32 gradient arena
              dgrad  im2col(pad k-1-p) of dY -> GEMM with flipped taps (+ReLU mask, +bias-grad column sums)
              pool   scatter through the stored argmax (+ReLU mask, +bias gradient of the producing conv)

The object offers the same interface as :class:`FusedMLP` (forward/backward/update over a
``buf`` dict whose ``acts[0]`` is the input batch), so the graph-capture driver in
``fused_mlp.try_fused_fit`` runs both.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from sparktorch_b200.ops import native
from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.fused_mlp import LinearSpec, _pad


@dataclasses.dataclass
class ConvSpec:
    module: nn.Conv2d
    relu: bool = False
    pool: int = 0            # k of a following k x k max-pool (0: none)
    p_drop: float = 0.0      # dropout2d after the pool


@dataclasses.dataclass
class LoweredCNN:
    in_shape: Tuple[int, int, int]      # (C, H, W) of the input image
    convs: List[ConvSpec]
    linears: List[LinearSpec]
    final_log_softmax: bool


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def lower_cnn(model: nn.Module, in_features: int) -> Optional[LoweredCNN]:
    """FX pattern: [view(-1,C,H,W)] (Conv2d [relu] [max_pool2d(k) [dropout/dropout2d]])+ flatten (Linear [relu])+ [log_softmax]."""
    import torch.fx as fx

    try:
        gm = fx.symbolic_trace(model)
    except Exception:
        return None
    mods = dict(gm.named_modules())
    convs: List[ConvSpec] = []
    linears: List[LinearSpec] = []
    in_shape = None
    flattened = False
    final_lsm = False
    prev = None
    for node in gm.graph.nodes:
        if node.op == "placeholder":
            if prev is not None:
                return None
            prev = node
            continue
        if node.op == "output":
            if node.args[0] is not prev:
                return None
            break
        if final_lsm:
            return None
        ins = [a for a in node.args if isinstance(a, fx.Node)]
        if len(ins) != 1 or ins[0] is not prev or len(prev.users) != 1:
            return None
        tgt = mods.get(node.target) if node.op == "call_module" else None
        is_relu = (isinstance(tgt, nn.ReLU) or (node.op == "call_function" and node.target in (F.relu, torch.relu))
                   or (node.op == "call_method" and node.target == "relu"))
        if node.op == "call_method" and node.target in ("view", "reshape") and not convs and not linears:
            shp = node.args[1:]
            if len(shp) == 1 and isinstance(shp[0], (tuple, list)):
                shp = tuple(shp[0])
            if len(shp) != 4 or shp[0] != -1 or not all(isinstance(v, int) for v in shp[1:]):
                return None
            in_shape = tuple(int(v) for v in shp[1:])
        elif isinstance(tgt, nn.Conv2d):
            if flattened or linears:
                return None
            if tgt.groups != 1 or _pair(tgt.dilation) != (1, 1) or tgt.padding_mode != "zeros" or isinstance(tgt.padding, str):
                return None
            if _pair(tgt.stride)[0] != _pair(tgt.stride)[1] or _pair(tgt.padding)[0] != _pair(tgt.padding)[1]:
                return None
            if convs and _pair(tgt.stride) != (1, 1):
                return None             # dgrad-as-correlation needs stride 1 beyond the first conv
            convs.append(ConvSpec(tgt))
        elif is_relu:
            if linears:
                if linears[-1].relu:
                    return None
                linears[-1].relu = True
            elif convs and not flattened and not convs[-1].relu and not convs[-1].pool:
                convs[-1].relu = True
            else:
                return None
        elif isinstance(tgt, nn.MaxPool2d) or (node.op == "call_function" and node.target is F.max_pool2d):
            if not convs or flattened or convs[-1].pool:
                return None
            if tgt is not None:
                k, s, p = _pair(tgt.kernel_size), _pair(tgt.stride or tgt.kernel_size), _pair(tgt.padding)
            else:
                k = _pair(node.args[1] if len(node.args) > 1 else node.kwargs.get("kernel_size"))
                s = node.args[2] if len(node.args) > 2 else node.kwargs.get("stride")
                s = k if s is None else _pair(s)
                p = _pair(node.args[3] if len(node.args) > 3 else node.kwargs.get("padding", 0))
            if k[0] != k[1] or s != k or p != (0, 0):
                return None
            convs[-1].pool = int(k[0])
        elif isinstance(tgt, (nn.Dropout2d, nn.Dropout)):
            # dropout2d right after the pool is fused; element-wise dropout is only accepted as a no-op (p=0)
            if isinstance(tgt, nn.Dropout2d) and convs and convs[-1].pool and not flattened and convs[-1].p_drop == 0.0:
                convs[-1].p_drop = float(tgt.p)
            elif tgt.p != 0.0:
                return None
        elif isinstance(tgt, nn.Flatten) or (node.op == "call_function" and node.target is torch.flatten) or \
                (node.op == "call_method" and node.target == "flatten"):
            start = node.args[1] if len(node.args) > 1 else node.kwargs.get("start_dim", 1 if tgt is not None else 0)
            if tgt is not None:
                start = tgt.start_dim
            if start != 1 or not convs or flattened:
                return None
            flattened = True
        elif isinstance(tgt, nn.Linear):
            if not flattened:
                return None
            linears.append(LinearSpec(tgt, False))
        elif isinstance(tgt, nn.LogSoftmax) and tgt.dim in (1, -1):
            final_lsm = True
        elif node.op == "call_function" and node.target in (F.log_softmax, torch.log_softmax):
            dim = node.kwargs.get("dim", node.args[1] if len(node.args) > 1 else None)
            if dim not in (1, -1):
                return None
            final_lsm = True
        elif isinstance(tgt, nn.Identity):
            pass
        else:
            return None
        prev = node
    if not convs or not linears or linears[-1].relu:
        return None
    if in_shape is None:
        return None                       # we need the image geometry (view(-1, C, H, W) at the top)
    C, H, W = in_shape
    if C * H * W != in_features or convs[0].module.in_channels != C:
        return None
    covered = set()
    for m in [c.module for c in convs] + [l.module for l in linears]:
        covered.add(id(m.weight))
        if m.bias is not None:
            covered.add(id(m.bias))
    params = [p for p in model.parameters() if p.requires_grad]
    if len(covered) != len(params) or any(id(p) not in covered for p in params) or list(model.buffers()):
        return None
    # geometry must chain up to the first Linear
    c, h, w = C, H, W
    for cs in convs:
        m = cs.module
        if m.in_channels != c:
            return None
        k, s, p = _pair(m.kernel_size), _pair(m.stride)[0], _pair(m.padding)[0]
        h, w, c = (h + 2 * p - k[0]) // s + 1, (w + 2 * p - k[1]) // s + 1, m.out_channels
        if cs.pool:
            h, w = h // cs.pool, w // cs.pool
        if h <= 0 or w <= 0:
            return None
    if linears[0].module.in_features != c * h * w:
        return None
    for a, b in zip(linears[:-1], linears[1:]):
        if a.module.out_features != b.module.in_features:
            return None
    return LoweredCNN(in_shape, convs, linears, final_lsm)


class FusedCNN:
    """Buffers + enqueue functions; interface-compatible with FusedMLP."""

    def __init__(self, model: nn.Module, low: LoweredCNN, kind: str, spec, world: int, rank: int, device: torch.device,
                 batch: int, val_rows: int, patience: int, max_iters: int, seed: int = 0, backend_factory=None,
                 inference: bool = False, ops=None, act_dtype: torch.dtype = torch.bfloat16):
        """``ops`` / ``act_dtype`` (and a ``backend_factory`` whose product sets ``wants_segments``): CPU emulation hooks."""
        from sparktorch_b200.parallel.p2p import FusedUpdate

        self.C = ops if ops is not None else native.load(required=True)
        self.act_dtype = act_dtype
        self.device, self.world, self.rank = device, world, rank
        self.kind, self.low = kind, low
        self.lowered = low
        self.B, self.nv = batch, val_rows
        self.seed = seed
        C0, H0, W0 = low.in_shape
        self.d0 = C0 * H0 * W0
        # ---- layer geometry -------------------------------------------------------------------------
        self.geo = []          # per conv: dict(Cin,H,W,Cout,kh,kw,s,p,OH,OW,K,pool,PH,PW)
        c, h, w = C0, H0, W0
        for cs in low.convs:
            m = cs.module
            kh, kw = _pair(m.kernel_size)
            s, p = _pair(m.stride)[0], _pair(m.padding)[0]
            oh, ow = (h + 2 * p - kh) // s + 1, (w + 2 * p - kw) // s + 1
            g = dict(Cin=c, H=h, W=w, Cout=m.out_channels, kh=kh, kw=kw, s=s, p=p, OH=oh, OW=ow, K=c * kh * kw,
                     pool=cs.pool, PH=oh // cs.pool if cs.pool else oh, PW=ow // cs.pool if cs.pool else ow)
            self.geo.append(g)
            c, h, w = m.out_channels, g["PH"], g["PW"]
        self.feat = c * h * w
        self.dl = [self.feat] + [ls.module.out_features for ls in low.linears]     # linear dims
        self.L = len(low.linears)
        dL = self.dl[-1]
        # ---- arenas ------------------------------------------------------------------------------------
        numel = FlatArena.measure(model)
        w_off, off = {}, 0
        mods = [cs.module for cs in low.convs] + [ls.module for ls in low.linears]
        for m in mods:
            rows = m.weight.shape[0]
            cols = m.weight[0].numel()
            w_off[id(m)] = (off, rows, cols)
            off += _pad(rows * _pad(cols), 64)
        if backend_factory is not None:
            self.fu = backend_factory(numel, off)
        else:
            self.fu = FusedUpdate(numel, spec, world, rank, device, patience, max_iters, wbf16_elems=off)
        self.arena = FlatArena(model, device, grad_buffer=self.fu.grad, param_buffer=self.fu.param)
        pos = {id(p): o for p, o in zip(self.arena.params, self.arena.offsets)}
        seg_of = {}
        self.Wb, self.bias, self.gW, self.gb, self.peer_kw = {}, {}, {}, {}, {}
        for m in mods:
            o, rows, cols = w_off[id(m)]
            ldw = _pad(cols)
            wo = pos[id(m.weight)]
            seg_of[wo] = [wo, rows, cols, o, ldw, -1, 0, 1]
            self.Wb[id(m)] = self.fu.wbf16[o: o + rows * ldw].view(rows, ldw)[:, :cols]
            self.gW[id(m)] = self.fu.grad[wo: wo + rows * cols].view(rows, cols)
            bo = -1
            if m.bias is not None:
                bo = pos[id(m.bias)]
                seg_of[bo] = [bo, 1, rows, -1, 0, -1, 0, 1]
                self.bias[id(m)] = self.fu.param[bo: bo + rows]
                self.gb[id(m)] = self.fu.grad[bo: bo + rows]
            else:
                self.bias[id(m)] = None
                self.gb[id(m)] = None
            peer = getattr(self.fu, "peer_table", None)
            self.peer_kw[id(m)] = dict(peer_table=peer, peer_shard_elems=self.fu.shard_elems, peer_w_off=wo, peer_bias_off=bo,
                                       peer_fill_local=True) if peer is not None else {}
        if backend_factory is None or getattr(self.fu, "wants_segments", False):
            self.fu.build([seg_of[o] for o in sorted(seg_of)], zero_grads=True)
        self.loss_slot = self.fu.grad[numel: numel + 1]
        self.train_loss_slot = self.fu.grad[numel + 2: numel + 3]
        # flipped-tap weights for the dgrad GEMMs of conv 2..n
        self.Wflip = {}
        for i, cs in enumerate(low.convs):
            if i > 0:
                g = self.geo[i]
                kflip = g["Cout"] * g["kh"] * g["kw"]
                self.Wflip[i] = torch.zeros(g["Cin"], _pad(kflip), dtype=self.fu.wbf16.dtype, device=device)[:, :kflip]
        self.fuse_loss = dL <= 32 and os.environ.get("SPARKTORCH_FUSE_LOSS", "1") == "1"
        self.window = None
        self.side = (torch.cuda.Stream(device=device)
                     if device.type == "cuda" and os.environ.get("SPARKTORCH_SIDE_STREAM", "1") == "1" else None)
        self.train = self._alloc(batch, backward=not inference)
        self.val = self._alloc(val_rows, backward=False) if val_rows > 0 else None
        self.n_launch = 0

    # ------------------------------------------------------------------------------------------------
    def _alloc(self, rows: int, backward: bool):
        dev, bf = self.device, self.act_dtype
        buf = {"rows": rows, "acts": [torch.zeros(rows, _pad(self.d0), dtype=bf, device=dev)[:, : self.d0]]}
        buf["col"], buf["conv_out"], buf["pool_out"], buf["pool_idx"] = [], [], [], []
        buf["dconv"], buf["colT"], buf["dpool"] = [], [], []
        for i, g in enumerate(self.geo):
            M = rows * g["OH"] * g["OW"]
            buf["col"].append(torch.zeros(M, _pad(g["K"]), dtype=bf, device=dev)[:, : g["K"]])
            buf["conv_out"].append(torch.zeros(M, _pad(g["Cout"]), dtype=bf, device=dev)[:, : g["Cout"]])
            last = i == len(self.geo) - 1
            if g["pool"] or last:
                if last:      # feeds the Linear head: one row per sample, torch flatten order (c, ph, pw)
                    n_out = g["Cout"] * g["PH"] * g["PW"]
                    buf["pool_out"].append(torch.zeros(rows, _pad(n_out), dtype=bf, device=dev)[:, :n_out])
                else:         # feeds another conv: NHWC, one row per pooled pixel
                    buf["pool_out"].append(torch.zeros(rows * g["PH"] * g["PW"], _pad(g["Cout"]), dtype=bf, device=dev)[:, : g["Cout"]])
                buf["pool_idx"].append(torch.zeros(rows * g["PH"] * g["PW"] * g["Cout"], dtype=torch.uint8, device=dev))
            else:
                buf["pool_out"].append(None)
                buf["pool_idx"].append(None)
            if backward:
                buf["dconv"].append(torch.zeros(M, _pad(g["Cout"]), dtype=bf, device=dev)[:, : g["Cout"]])
                if i > 0:
                    Min = rows * g["H"] * g["W"]
                    kflip = g["Cout"] * g["kh"] * g["kw"]
                    buf["colT"].append(torch.zeros(Min, _pad(kflip), dtype=bf, device=dev)[:, :kflip])
                else:
                    buf["colT"].append(None)
                # gradient wrt an NHWC pooled map (written by the dgrad GEMM of the next conv)
                buf["dpool"].append(torch.zeros(rows * g["PH"] * g["PW"], _pad(g["Cout"]), dtype=bf, device=dev)[:, : g["Cout"]]
                                    if (g["pool"] and not last) else None)
        # linear head
        buf["lin"] = [None] + [torch.zeros(rows, _pad(self.dl[l]), dtype=bf, device=dev)[:, : self.dl[l]] for l in range(1, self.L)]
        buf["logits_full"] = torch.zeros(rows, _pad(self.dl[-1], 16), dtype=torch.float32, device=dev)
        buf["logits"] = buf["logits_full"][:, : self.dl[-1]]
        if backward:
            buf["dz"] = [torch.zeros(rows, _pad(self.dl[l]), dtype=bf, device=dev)[:, : self.dl[l]] for l in range(0, self.L + 1)]
        buf["tgt"] = (torch.zeros(rows, dtype=torch.int32, device=dev) if self.kind == "ce"
                      else torch.zeros(rows, self.dl[-1], dtype=torch.float32, device=dev))
        return buf

    def _gemm(self, *a, **kw):
        self.C.gemm_bf16(*a, **kw)
        self.n_launch += 1

    def _win(self, flags: int) -> dict:
        if self.window is None or flags == 0:
            return {}
        return dict(win_state=self.fu.state, window_rows=self.window[1], n_windows=self.window[0], win_flags=flags)

    # ------------------------------------------------------------------------------------------------
    def forward(self, buf, slot=None, train: bool = False, windowed: bool = False) -> None:
        C, rows, low = self.C, buf["rows"], self.low
        x = buf["acts"][0]
        # source strides of the current feature map: start with torch's (c, h, w) row layout
        g0 = self.geo[0]
        src, strides = x, (x.stride(0), g0["W"], 1, g0["H"] * g0["W"])           # (sb, sh, sw, sc)
        for i, (cs, g) in enumerate(zip(low.convs, self.geo)):
            m = cs.module
            wkw = {}
            if windowed and i == 0 and self.window is not None:
                wkw = dict(state=self.fu.state, window_rows=self.window[1], n_windows=self.window[0])
            C.im2col_nhwc(src, rows, g["H"], g["W"], g["Cin"], strides[0], strides[1], strides[2], strides[3], g["kh"], g["kw"],
                          g["s"], g["p"], buf["col"][i], **wkw)
            self.n_launch += 1
            M = rows * g["OH"] * g["OW"]
            ckw = dict(bias=self.bias[id(m)], relu=cs.relu, out_bf16=buf["conv_out"][i])
            if self.peer_kw[id(m)] and buf is self.train:      # K-AG: conv weights gathered from their owners in the GEMM
                ckw["bias"] = None
                ckw.update(self.peer_kw[id(m)])
            self._gemm(buf["col"][i], False, self.Wb[id(m)], False, M, g["Cout"], g["K"], **ckw)
            out = buf["conv_out"][i]
            ldo = out.stride(0)
            last = i == len(self.geo) - 1
            if g["pool"] or last:
                k = g["pool"] or 1
                C.maxpool_dropout(False, out, ldo, buf["pool_out"][i], buf["pool_idx"][i], rows, g["OH"], g["OW"], g["Cout"], k,
                                  p_drop=cs.p_drop, train=train, seed=self.seed + 1000003 * i, state=self.fu.state,
                                  y_nhwc=not last)
                self.n_launch += 1
                po = buf["pool_out"][i]
                src = po
                ldp = po.stride(0)
                strides = (g["PH"] * g["PW"] * ldp, g["PW"] * ldp, ldp, 1)        # NHWC pooled map feeding the next conv
            else:
                src = out
                strides = (g["OH"] * g["OW"] * ldo, g["OW"] * ldo, ldo, 1)
        a = buf["pool_out"][len(self.geo) - 1]
        for l in range(1, self.L + 1):
            ls = low.linears[l - 1]
            m = ls.module
            last = l == self.L
            kw = dict(bias=self.bias[id(m)], relu=ls.relu)
            if self.peer_kw[id(m)] and buf is self.train:
                kw["bias"] = None
                kw.update(self.peer_kw[id(m)])
            if not last:
                kw["out_bf16"] = buf["lin"][l]
            elif slot is not None and self.fuse_loss and not self._deep_head(rows):
                kw.update(loss_kind=1 if self.kind == "ce" else 2, loss_slot=slot,
                          loss_scale=(1.0 / rows) if self.kind == "ce" else 1.0 / (rows * self.dl[-1]))
                kw["labels" if self.kind == "ce" else "target"] = buf["tgt"]
                if train:
                    kw.update(out_bf16=buf["dz"][l], colsum=self.gb[id(m)])
                kw.update(self._win(4 if windowed else 0))
            else:
                kw["out_f32"] = buf["logits"]
                if last and slot is not None and self._deep_head(rows):
                    # a [rows<=256, K>=2048] x [K, <=32] head is ONE tile: split K over many SMs (fp32 red.add into
                    # the zeroed logits) and run the loss as its own small kernel instead of serialising on one SM
                    buf["logits_full"].zero_()
                    self.n_launch += 1
                    kb = (self.dl[l - 1] + 63) // 64
                    kw.update(atomic_f32=True, split_k=max(2, min(32, kb // 4)))
            self._gemm(a, False, self.Wb[id(m)], False, rows, self.dl[l], self.dl[l - 1], **kw)
            a = buf["lin"][l] if not last else None
        if slot is not None and (not self.fuse_loss or self._deep_head(rows)):
            dz = buf["dz"][self.L] if train else None
            db = self.gb[id(low.linears[-1].module)] if train else None
            fn = self.C.softmax_ce if self.kind == "ce" else self.C.mse_loss
            lw = {}
            if windowed and self.window is not None:      # labels of the resident partition: same window as the images
                lw = dict(win_state=self.fu.state, window_rows=self.window[1], n_windows=self.window[0])
            fn(buf["logits"], buf["tgt"], self.dl[-1], slot, dz, None, db, _pad(self.dl[-1]), **lw)
            self.n_launch += 1

    def _deep_head(self, rows: int) -> bool:
        return rows <= 256 and self.dl[-2] >= 2048 and os.environ.get("SPARKTORCH_SPLIT_HEAD", "1") == "1"

    def _split_k(self, M: int, N: int, K: int) -> int:
        tiles = ((M + 127) // 128) * ((N + 63) // 64)
        kb = (K + 63) // 64
        if kb < 8 or tiles >= 148:
            return 1
        return max(1, min(kb // 4, 148 // tiles))

    def backward(self, buf, windowed: bool = False) -> None:
        C, rows, low = self.C, buf["rows"], self.low
        dz = buf["dz"]
        n_conv = len(self.geo)
        feat_in = buf["pool_out"][n_conv - 1]
        # ---- linear head ------------------------------------------------------------------------------
        for l in range(self.L, 0, -1):
            m = low.linears[l - 1].module
            a_prev = feat_in if l == 1 else buf["lin"][l - 1]
            sk = self._split_k(self.dl[l], self.dl[l - 1], rows)
            self._gemm(dz[l], True, a_prev, True, self.dl[l], self.dl[l - 1], rows, out_f32=self.gW[id(m)], atomic_f32=sk > 1, split_k=sk)
            relu_prev = low.linears[l - 2].relu if l > 1 else False
            colsum = self.gb[id(low.linears[l - 2].module)] if l > 1 else None
            self._gemm(dz[l], False, self.Wb[id(m)], True, rows, self.dl[l - 1], self.dl[l],
                       mask_src=a_prev if relu_prev else None, out_bf16=dz[l - 1], colsum=colsum)
        # dz[0] = gradient wrt the flattened feature map, (c, ph, pw) order
        grad_in = dz[0]
        for i in range(n_conv - 1, -1, -1):
            cs, g = low.convs[i], self.geo[i]
            m = cs.module
            M = rows * g["OH"] * g["OW"]
            dconv = buf["dconv"][i]
            last = i == n_conv - 1
            if g["pool"] or last:
                k = g["pool"] or 1
                if g["OH"] % k or g["OW"] % k:
                    dconv.zero_()
                C.maxpool_dropout(True, dconv, dconv.stride(0), grad_in, buf["pool_idx"][i], rows, g["OH"], g["OW"], g["Cout"], k,
                                  p_drop=cs.p_drop, train=True, seed=self.seed + 1000003 * i, state=self.fu.state,
                                  y_nhwc=not last, relu_x=buf["conv_out"][i] if cs.relu else None, db=self.gb[id(m)])
                self.n_launch += 1
            # else: dconv was produced (masked, with its bias gradient) by the dgrad GEMM of conv i+1
            sk = self._split_k(g["Cout"], g["K"], M)
            self._gemm(dconv, True, buf["col"][i], True, g["Cout"], g["K"], M, out_f32=self.gW[id(m)], atomic_f32=sk > 1, split_k=sk)
            if i > 0:
                gp = self.geo[i - 1]
                prev = low.convs[i - 1]
                C.flip_conv_weight(self.Wb[id(m)], g["Cout"], g["Cin"], g["kh"], g["kw"], self.Wflip[i])
                ldd = dconv.stride(0)
                C.im2col_nhwc(dconv, rows, g["OH"], g["OW"], g["Cout"], g["OH"] * g["OW"] * ldd, g["OW"] * ldd, ldd, 1,
                              g["kh"], g["kw"], 1, g["kh"] - 1 - g["p"], buf["colT"][i])
                self.n_launch += 2
                Min = rows * g["H"] * g["W"]
                kflip = g["Cout"] * g["kh"] * g["kw"]
                if gp["pool"]:
                    # gradient wrt the pooled map of conv i-1 (NHWC); its pool backward applies mask + bias gradient
                    self._gemm(buf["colT"][i], False, self.Wflip[i], False, Min, g["Cin"], kflip, out_bf16=buf["dpool"][i - 1])
                    grad_in = buf["dpool"][i - 1]
                else:
                    pm = prev.module
                    self._gemm(buf["colT"][i], False, self.Wflip[i], False, Min, g["Cin"], kflip,
                               mask_src=buf["conv_out"][i - 1] if prev.relu else None, out_bf16=buf["dconv"][i - 1],
                               colsum=self.gb[id(pm)])
                    grad_in = None

    def update(self) -> None:
        self.fu.updater.step()
        self.n_launch += 1
