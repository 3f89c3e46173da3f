"""Execution engine of the PeCLR step on one B200: owns the flat fp32 parameter / gradient buffers and the
bf16 operand copies, and sequences the CUDA kernels (through the C ABI) for

    two-view batch -> stem -> ResNet blocks -> avg-pool -> projection MLP -> fused equivariance + NT-Xent
    (forward and backward), i.e. Hybrid2Model.training_step + loss.backward() of the reference
    (src/models/unsupervised/hybrid2_model.py:27-90, simclr_model.py:59-67).

Memory layout (HBM): one flat fp32 master buffer with every trainable tensor in ``named_parameters()`` order
(conv weights channels-last = [Cout][k*k][Cin]), a same-shaped fp32 gradient buffer, Adam moments, a bf16 copy
of the master (fprop / wgrad operand), a bf16 buffer of transposed conv weights [Cin][k*k][Cout] (dgrad
operand) and the packed stem weights.  Activations are NHWC bf16; per layer the raw conv output y and the
post-BN activation a are kept for the backward pass (RN50 @ 224^2, 256 images: ~11 GB of 180 GB).
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .resnet_model import BNHolder, Block, ConvHolder

bf16 = torch.bfloat16


class _Seg:
    __slots__ = ("name", "begin", "size", "shape", "is_conv", "module", "pname")


class StepEngine:
    def __init__(self, model: nn.Module):
        self.model = model
        self.encoder = model.encoder
        self.head = getattr(model, "projection_head", None)
        self.device = torch.device("cpu")
        self.segs = []
        self._by_param = {}
        self._flatten()
        self.weights_dirty = True
        self.ctx = None
        self.world, self.rank = 1, 0
        self._sym = None  # symmetric-memory handles for the fused embedding all-gather
        self._epoch = 0
        self._ws = {}
        self.overlap_wgrad = os.environ.get("PECLR_OVERLAP_WGRAD", "1") != "0"
        self.fuse_bn_reduce = os.environ.get("PECLR_FUSE_BN_REDUCE", "1") != "0"
        # block-output BatchNorm backward: the dgrad that completes a block input's gradient also masks it with the
        # previous block's ReLU bits and accumulates that block's BN-backward sums (peclr_conv2d_dgrad_finish)
        self.fuse_block_bn = os.environ.get("PECLR_FUSE_BLOCK_BN", "1") != "0"
        # stride-2 shortcut dgrad scattered into an unzeroed buffer, read on its pixel lattice by the finishing dgrad
        self.finish_lattice = os.environ.get("PECLR_FINISH_LATTICE", "1") != "0"
        # weight gradients, optional: every convolution of a ResNet stage writes its pixel-split slabs into its own
        # region of one workspace and ONE ordered reduction per stage adds them to the gradients (instead of one per
        # convolution).  Measured SLOWER on B200 (same box: ResNet-50 17.55 vs 17.28 ms, ResNet-152 39.67 vs 38.85 ms):
        # the slabs of a whole stage (0.3-2 GB) fall out of the L2 before they are read back, the per-convolution
        # reduction finds much of its 30-40 MB still there.  Off by default; kept as a switch.
        self.batch_wgrad_reduce = os.environ.get("PECLR_BATCH_WGRAD_REDUCE", "0") != "0"
        self._wgrad_plans = {}
        self._side = None
        # workspaces of the split reductions (ordered, atomics-free: see include/peclr_b200.h).  The conv weight
        # gradients share one (they run in order on one stream), the stem's runs on the main stream and has its own;
        # the head's holds zero-at-rest tile counters.
        self._wgrad_ws, self._stem_ws, self._head_ws = ops.Workspace(), ops.Workspace(), ops.Workspace(zero=True)

    # ------------------------------------------------------------------ parameter arena
    def _trainable(self):
        for mod_name, mod in self.model.named_modules():
            if mod_name.startswith("encoder.final_layer"):
                continue
            for pname, p in mod._parameters.items():
                if p is not None:
                    yield (mod_name + "." + pname if mod_name else pname), mod, pname, p

    def _flatten(self, device=None):
        """(Re)build the flat buffers on `device` and rebind every parameter (and its .grad) as a view."""
        device = torch.device(device) if device is not None else self.device
        items = list(self._trainable())
        total = 0
        segs = []
        for name, mod, pname, p in items:
            s = _Seg()
            s.name, s.begin, s.size, s.shape = name, total, p.numel(), tuple(p.shape)
            s.is_conv, s.module, s.pname = p.dim() == 4, mod, pname
            assert s.begin % 8 == 0, (name, s.begin)  # 16-byte alignment of every bf16 operand for TMA
            total += s.size
            segs.append(s)
        flat = torch.empty(total, dtype=torch.float32, device=device)
        grads = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for s, (_, _, _, p) in zip(segs, items):
                src = p.detach().permute(0, 2, 3, 1) if s.is_conv else p.detach()
                flat[s.begin:s.begin + s.size].copy_(src.reshape(-1))
        for s in segs:
            s.module._parameters[s.pname] = nn.Parameter(self._view(flat, s))
            s.module._parameters[s.pname].grad = self._view(grads, s)
        self.flat, self.grads, self.segs, self.total = flat, grads, segs, total
        self._by_param = {s.name: s for s in segs}
        self._seg_index = {(id(s.module), s.pname): s for s in segs}
        self.device = device
        self.exp_avg = self.exp_avg_sq = None
        self.w_bf16 = self.wt_bf16 = self.w_stem = None
        # every BatchNorm shares ONE num_batches_tracked counter (they always advance together)
        nbt = None
        for m in self.model.modules():
            if isinstance(m, (BNHolder, nn.BatchNorm1d)):
                for k in ("running_mean", "running_var"):
                    m._buffers[k] = m._buffers[k].to(device)
                if nbt is None:
                    nbt = m._buffers["num_batches_tracked"].to(device)
                m._buffers["num_batches_tracked"] = nbt
        self.nbt = nbt
        fl = self.encoder.final_layer[0]
        for k in ("weight", "bias"):
            fl._parameters[k] = nn.Parameter(fl._parameters[k].detach().to(device))
        self._build_plan()
        self.weights_dirty = True

    @staticmethod
    def _view(buf, s):
        v = buf[s.begin:s.begin + s.size]
        if s.is_conv:
            co, ci, kh, kw = s.shape
            return v.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return v.view(s.shape)

    def to(self, device):
        if torch.device(device) != self.device:
            self._flatten(device)
        return self

    # ------------------------------------------------------------------ static plan
    def _seg(self, module, pname="weight"):
        return self._seg_index[(id(module), pname)]

    def _build_plan(self):
        enc = self.encoder
        f = enc.features
        self.stem_conv, self.stem_bn = f[0], f[1]
        self.blocks = [b for layer in (f[4], f[5], f[6], f[7]) for b in layer]
        # transposed (dgrad) weights for every conv except the stem
        self.t_off = {}
        entries, off = [], 0
        for m in enc.modules():
            if isinstance(m, ConvHolder) and m is not self.stem_conv:
                s = self._seg(m)
                self.t_off[id(m)] = off
                entries.append((s.begin, off, m.cout, m.k * m.k, m.cin))
                off += s.size
        self.t_total = off
        self._t_entries = entries
        self._t_table = None
        self._opt_tables = None

    def weight_decay_table(self, weight_decay, skip=("bias", "bn")):
        """exclude_from_wt_decay (base_model.py:30-51) applied to the flat segments."""
        return [0.0 if any(k in s.name for k in skip) else weight_decay for s in self.segs]

    # ------------------------------------------------------------------ bf16 operand copies
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise _lib.PeclrKernelError(
                "the PeCLR step runs only on a CUDA (sm_100a) device: move the model with .cuda(); "
                "there is no CPU path")

    def sync_weights(self):
        """Refresh the bf16 / transposed / packed weight copies from the fp32 master."""
        self._require_cuda()
        dev = self.device
        if self.w_bf16 is None:
            self.w_bf16 = torch.empty(self.total, dtype=bf16, device=dev)
            self.wt_bf16 = torch.empty(self.t_total, dtype=bf16, device=dev)
            self.w_stem = torch.empty((64, 4, 64), dtype=bf16, device=dev)
            self._t_table = ops.build_transpose_table(self._t_entries, dev)
        _lib.call("peclr_cast_bf16", self.flat, self.w_bf16, self.total, ops._s())
        self.refresh_derived_weights()
        self.weights_dirty = False

    def refresh_derived_weights(self):
        ops.weight_transpose(self.flat, self.wt_bf16, self._t_table)
        s = self._seg(self.stem_conv)
        _lib.call("peclr_stem_pack", self.flat[s.begin:], self.w_stem, ops._s())

    def _w(self, conv):  # [Cout, k*k, Cin] bf16
        s = self._seg(conv)
        return self.w_bf16[s.begin:s.begin + s.size].view(conv.cout, conv.k * conv.k, conv.cin)

    def _wt(self, conv):  # [Cin, k*k, Cout] bf16
        o = self.t_off[id(conv)]
        return self.wt_bf16[o:o + conv.cout * conv.k * conv.k * conv.cin].view(conv.cin, conv.k * conv.k, conv.cout)

    def _p(self, module, pname):  # fp32 master slice
        s = self._seg(module, pname)
        return self.flat[s.begin:s.begin + s.size]

    def _g(self, module, pname):
        s = self._seg(module, pname)
        return self.grads[s.begin:s.begin + s.size]

    # ------------------------------------------------------------------ forward
    def _bn_stats_for(self, bn, m, training, stats):
        """Training: the conv epilogue's sums.  Eval: sums synthesised from the running statistics."""
        if training:
            return stats
        mean, var = bn.running_mean.double(), bn.running_var.double()
        return torch.stack([mean * m, (var + mean * mean) * m])

    def forward_trunk(self, img1, img2, training=True):
        self._require_cuda()
        if self.weights_dirty:
            self.sync_weights()
        b, _, h, w = img1.shape
        n = 2 * b if img2 is not None else b  # img2 None: img1 is the whole batch (any size)
        dev = self.device
        ctx = {"n": n, "hw": (h, w), "blocks": []}
        convs = [self.stem_conv] + [c for blk in self.blocks for c, _ in blk.convs()] + \
                [blk.downsample[0] for blk in self.blocks if blk.downsample is not None]
        tot = sum(c.cout for c in convs)
        stats_all = torch.zeros((2 * tot,), dtype=torch.float64, device=dev)  # fp64 cross-CTA accumulators
        saved_all = torch.empty((2 * tot,), dtype=torch.float32, device=dev)
        cur = [0]

        def slab(c):
            o = cur[0]
            cur[0] += 2 * c
            return stats_all[o:o + 2 * c].view(2, c), saved_all[o:o + 2 * c].view(2, c)

        def run(bn):
            return (bn.running_mean, bn.running_var) if training else None

        xpad = ops.stem_input(img1, img2)
        st, sv = slab(64)
        y0 = ops.stem_fprop(xpad, self.w_stem, h, w, stats=st if training else None)
        bn0 = self.stem_bn
        a, _, pool_idx = ops.stem_bn_relu_pool(y0, self._bn_stats_for(bn0, n * (h // 2) * (w // 2), training, st),
                                               self._p(bn0, "weight"), self._p(bn0, "bias"), run(bn0), bn0.eps,
                                               bn0.momentum, saved=sv, want_idx=training)
        ctx["stem"] = (xpad, y0, sv, pool_idx)
        hh, ww = h // 4, w // 4
        for blk in self.blocks:
            rec = {"a_in": a, "in_hw": (hh, ww), "convs": []}
            x = a
            pairs = blk.convs()
            for i, (conv, bn) in enumerate(pairs):
                st, sv = slab(conv.cout)
                y = ops.conv2d_fprop(x, self._w(conv), conv.k, conv.stride, stats=st if training else None)
                m = y.numel() // conv.cout
                last = i == len(pairs) - 1
                if not last:
                    act, _ = ops.bn_apply(y, self._bn_stats_for(bn, m, training, st), self._p(bn, "weight"),
                                          self._p(bn, "bias"), relu=True, running=run(bn), eps=bn.eps,
                                          momentum=bn.momentum, saved=sv)
                    rec["convs"].append((conv, bn, x, y, act, sv))
                    x = act
                else:
                    # ReLU mask of the block output as bits (1 byte per 8 channels) for the backward pass
                    mbits = torch.empty((m, conv.cout // 8), dtype=torch.uint8, device=dev) if training else None
                    if blk.downsample is not None:
                        dconv, dbn = blk.downsample[0], blk.downsample[1]
                        dst, dsv = slab(dconv.cout)
                        yd = ops.conv2d_fprop(rec["a_in"], self._w(dconv), 1, dconv.stride,
                                              stats=dst if training else None)
                        out, _, _ = ops.bn_apply(
                            y, self._bn_stats_for(bn, m, training, st), self._p(bn, "weight"), self._p(bn, "bias"),
                            relu=True, res=yd,
                            res_bn=(self._bn_stats_for(dbn, m, training, dst), self._p(dbn, "weight"),
                                    self._p(dbn, "bias"), run(dbn)),
                            running=run(bn), eps=bn.eps, momentum=bn.momentum, saved=sv, rsaved=dsv, mask_out=mbits)
                        rec["down"] = (dconv, dbn, yd, dsv)
                    else:
                        out, _ = ops.bn_apply(y, self._bn_stats_for(bn, m, training, st), self._p(bn, "weight"),
                                              self._p(bn, "bias"), relu=True, res=rec["a_in"], running=run(bn),
                                              eps=bn.eps, momentum=bn.momentum, saved=sv, mask_out=mbits)
                    rec["mask"] = mbits
                    rec["convs"].append((conv, bn, x, y, out, sv))
                    a = out
            hh, ww = a.shape[1], a.shape[2]
            ctx["blocks"].append(rec)
        enc = ops.avgpool_fwd(a)
        ctx["final_shape"] = tuple(a.shape)
        if training:
            self.nbt.add_(1)
        return enc, ctx

    def forward_head(self, enc, training=True):
        lin1, bn, _, lin2 = self.head
        h1 = ops.linear_fwd(enc, self._p(lin1, "weight").view(lin1.out_features, lin1.in_features),
                            self._p(lin1, "bias"), ws=self._head_ws)
        if training:
            a1, saved = ops.bn1d_relu_fwd(h1, self._p(bn, "weight"), self._p(bn, "bias"),
                                          (bn.running_mean, bn.running_var), bn.eps, bn.momentum)
        else:
            # eval-mode BatchNorm1d: x * scale + shift from the running statistics (tiny: 2N x 512)
            scale = self._p(bn, "weight") * torch.rsqrt(bn.running_var + bn.eps)
            a1 = torch.relu(h1 * scale + (self._p(bn, "bias") - bn.running_mean * scale))
            saved = None
        p = ops.linear_fwd(a1, self._p(lin2, "weight").view(lin2.out_features, lin2.in_features), ws=self._head_ws)
        return p, (enc, h1, a1, saved)

    def encode(self, x):
        """Inference-style trunk forward on fp32 NCHW images, any batch size -- as the reference's ``encoder(x)``
        (uses the module's train/eval mode)."""
        enc, _ = self.forward_trunk(x.contiguous(), None, training=self.model.training)
        return enc

    # ------------------------------------------------------------------ loss
    def loss_workspace(self, b):
        key = (b, self.world)
        if key not in self._ws:
            self._ws[key] = ops.ntxent_workspace(b, self.world, self.device)
        return self._ws[key]

    def forward_loss(self, p, angle, jx, jy, hw, crop, rotate, temperature=0.5, want_grad=True):
        b = p.shape[0] // 2
        kw = {}
        if self.world > 1:
            sym = self._symmetric(b)
            kw = dict(workspace=sym["ws"], world=self.world, rank=self.rank, z_peers=sym["z_ptrs"],
                      flag_peers=sym["flag_ptrs"])
        else:
            kw = dict(workspace=self.loss_workspace(b))
        return ops.ntxent_fused(p, angle, jx, jy, hw, crop, rotate, temperature, want_grad=want_grad, **kw)

    def _symmetric(self, b):
        """Peer-mapped workspaces for the fused embedding all-gather (torch symmetric memory over NVLink)."""
        if self._sym is not None and self._sym["b"] == b:
            return self._sym
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        nbytes = _lib.call("peclr_ntxent_workspace_bytes", b, self.world)
        nflt = (nbytes + 3) // 4 + 64
        ws = symm.empty(nflt, dtype=torch.float32, device=self.device)
        ws.zero_()
        hdl = symm.rendezvous(ws, dist.group.WORLD.group_name)
        flag_off = (nbytes + 3) // 4  # flags live behind the kernel's own workspace, inside the mapped buffer
        z_ptrs = torch.tensor([int(pp) for pp in hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
        flag_ptrs = torch.tensor([int(pp) + 4 * flag_off for pp in hdl.buffer_ptrs], dtype=torch.int64,
                                 device=self.device)
        torch.cuda.synchronize()
        dist.barrier()
        self._sym = dict(b=b, ws=ws, hdl=hdl, z_ptrs=z_ptrs, flag_ptrs=flag_ptrs)
        return self._sym

    # ------------------------------------------------------------------ backward
    def backward_head(self, g_p, head_ctx):
        enc, h1, a1, saved = head_ctx
        lin1, bn, _, lin2 = self.head
        w2 = self._p(lin2, "weight").view(lin2.out_features, lin2.in_features)
        w1 = self._p(lin1, "weight").view(lin1.out_features, lin1.in_features)
        hw = self._head_ws
        ops.linear_wgrad(g_p, a1, self._g(lin2, "weight").view_as(w2), ws=hw)
        da1 = ops.linear_dgrad(g_p, w2, ws=hw)
        dh1 = ops.bn1d_relu_bwd(da1, a1, h1, saved, self._p(bn, "weight"), self._g(bn, "weight"), self._g(bn, "bias"))
        ops.linear_wgrad(dh1, enc, self._g(lin1, "weight").view_as(w1), ws=hw)
        ops.colsum_acc(dh1, self._g(lin1, "bias"))
        return ops.linear_dgrad(dh1, w1, ws=hw)

    def backward_trunk(self, d_enc, ctx, after_stage=None):
        """after_stage(i): called when the gradients of ResNet stage i (and everything after it) are final --
        the hook data-parallel training uses to overlap the gradient all-reduce with the rest of backward."""
        dev = self.device
        n = ctx["n"]
        da = ops.avgpool_bwd(d_enc, ctx["final_shape"])
        scratch = ops.new_scratch(2048, dev)
        # Weight gradients are off the critical path (they only feed the optimiser): they go to a side stream so
        # the tensor-core-bound wgrad kernels overlap with the HBM-bound BatchNorm / dgrad chain.  Their inputs
        # are kept alive until the streams join (no allocator reuse while the side stream may still read them).
        main = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        side, keep = self._side, []
        overlap = self.overlap_wgrad
        # size the weight-gradient workspace for the largest launch of this pass up front (on the main stream, so the
        # side stream never sees it reallocated under a running kernel)
        plan = self._wgrad_plan(ctx) if self.batch_wgrad_reduce else None
        if plan is None:
            self._wgrad_ws.get(self._max_wgrad_bytes(ctx), dev)

        def conv_wgrad(conv, x, dy):
            """dW of one convolution (enqueued through wgrad(): side stream)."""
            if plan is None:
                ops.conv2d_wgrad(x, dy, conv.k, conv.stride, dw=self._g(conv, "weight"), ws=self._wgrad_ws)
            else:
                ops.conv2d_wgrad_partials(x, dy, conv.k, conv.stride, self._g(conv, "weight"), plan["region"][id(conv)])

        def wgrad(fn, *alive):
            if not overlap:
                return fn()
            keep.extend(alive)
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                fn()
        stage_of = {}
        f = self.encoder.features
        for si, layer in enumerate((f[4], f[5], f[6], f[7])):
            for blk in layer:
                stage_of[id(blk)] = si
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk, rec = self.blocks[bi], ctx["blocks"][bi]
            a_in = rec["a_in"]
            ih, iw = rec["in_hw"]
            convs = rec["convs"]
            # last conv of the block: BN backward with the block-output ReLU mask; g feeds the shortcut
            conv, bn, x, y, out, sv = convs[-1]
            if rec.get("prefinished"):
                # da already is g = (complete gradient) * relu'(out) and scratch holds this BatchNorm's sums: both
                # came out of the next block's finishing dgrad.  dy goes to its own buffer (g stays the shortcut's).
                g = da
                dy = ops.bn_backward(da, None, y, sv, self._p(bn, "weight"), self._g(bn, "weight"),
                                     self._g(bn, "bias"), scratch=scratch, reduce_done=True)
                rec["prefinished"] = False
            else:
                dy, g = ops.bn_backward(da, rec["mask"], y, sv, self._p(bn, "weight"), self._g(bn, "weight"),
                                        self._g(bn, "bias"), want_g=True, scratch=scratch, dy=da)
            da = None
            wgrad(lambda x=x, dy=dy, conv=conv: conv_wgrad(conv, x, dy), x, dy)
            cur_hw = (x.shape[1], x.shape[2])
            for i in range(len(convs) - 1, 0, -1):
                conv, bn, x, y, act, sv = convs[i]
                pconv, pbn, px, py, pact, psv = convs[i - 1]
                # inner BN + ReLU: the ReLU mask is recomputed from y (no read of the stored activation) and the
                # reduction pass of the BN backward runs inside the dgrad epilogue
                fused = self.fuse_bn_reduce and not (conv.k == 1 and conv.stride == 2)
                if fused:
                    dx = ops.conv2d_dgrad_bnreduce(dy, self._wt(conv), tuple(x.shape), conv.k, conv.stride, py, psv,
                                                   self._p(pbn, "weight"), self._p(pbn, "bias"), scratch)
                else:
                    dx = ops.conv2d_dgrad(dy, self._wt(conv), tuple(x.shape), conv.k, conv.stride)
                dy = ops.bn_backward(dx, None, py, psv, self._p(pbn, "weight"), self._g(pbn, "weight"),
                                     self._g(pbn, "bias"), scratch=scratch, dy=dx, beta=self._p(pbn, "bias"),
                                     reduce_done=fused)
                wgrad(lambda px=px, dy=dy, pconv=pconv: conv_wgrad(pconv, px, dy), px, dy)
            conv1 = convs[0][0]
            # the dgrad of conv1 completes the gradient of this block's input = the previous block's output
            prev = ctx["blocks"][bi - 1] if bi > 0 else None
            finish = (self.fuse_block_bn and prev is not None and conv1.k == 1 and conv1.stride == 1
                      and conv1.cin % 128 == 0 and prev["mask"] is not None)
            if finish:
                y_prev = prev["convs"][-1][3]
            if "down" in rec:
                dconv, dbn, yd, dsv = rec["down"]
                dyd = ops.bn_backward(g, None, yd, dsv, self._p(dbn, "weight"), self._g(dbn, "weight"),
                                      self._g(dbn, "bias"), scratch=scratch, dy=g)
                wgrad(lambda a_in=a_in, dyd=dyd, dconv=dconv: conv_wgrad(dconv, a_in, dyd), a_in, dyd)
                if finish:  # down-sampling branch first, conv1 finishes.  Its stride-2 form only touches the even /
                    # even pixel lattice: scattered into a buffer that is never zeroed, the finishing dgrad reads the
                    # other pixels as 0 (no 100-400 MB memset, no re-read of it)
                    lattice = self.finish_lattice and dconv.stride == 2 and dconv.k == 1
                    da = ops.conv2d_dgrad(dyd, self._wt(dconv), tuple(a_in.shape), 1, dconv.stride,
                                          scatter_only=lattice)
                    ops.conv2d_dgrad_finish(dy, self._wt(conv1), da, y_prev, prev["mask"], scratch,
                                            acc_stride=2 if lattice else 1)
                else:
                    da = ops.conv2d_dgrad(dy, self._wt(conv1), tuple(a_in.shape), conv1.k, conv1.stride)
                    ops.conv2d_dgrad(dyd, self._wt(dconv), tuple(a_in.shape), 1, dconv.stride, out=da, accumulate=True)
            else:
                da = g  # identity shortcut: start from g and add the main branch
                if finish:
                    ops.conv2d_dgrad_finish(dy, self._wt(conv1), da, y_prev, prev["mask"], scratch)
                else:  # (TMA reduce-add)
                    ops.conv2d_dgrad(dy, self._wt(conv1), tuple(a_in.shape), conv1.k, conv1.stride, out=da,
                                     accumulate=True)
            if finish:
                prev["prefinished"] = True
            if bi == 0 or stage_of[id(self.blocks[bi - 1])] != stage_of[id(blk)]:
                if plan is not None:  # this stage's weight gradients: one ordered reduction of all their slabs
                    wgrad(lambda t=plan["tables"][stage_of[id(blk)]]: ops.wgrad_reduce_batched(t))
                if after_stage is not None:
                    after_stage(stage_of[id(blk)])
        # stem: max-pool + ReLU + BN backward, then the 7x7 weight gradient
        xpad, y0, sv0, pool_idx = ctx["stem"]
        h, w = ctx["hw"]
        bn0 = self.stem_bn
        dy0 = ops.stem_pool_bn_backward(da, pool_idx, y0, sv0, self._p(bn0, "weight"), self._p(bn0, "bias"),
                                        self._g(bn0, "weight"), self._g(bn0, "bias"), scratch=scratch)
        dwp = ops.stem_wgrad(xpad, dy0, h, w, ws=self._stem_ws)
        _lib.call("peclr_stem_unpack_grad", dwp, self._g(self.stem_conv, "weight"), ops._s())
        if overlap:
            main.wait_stream(side)  # join: all weight gradients are complete when backward returns
        keep.clear()
        if after_stage is not None:
            after_stage(-1)

    def _wgrad_plan(self, ctx):
        """Workspace regions (one per convolution with more than one pixel split) and the per-stage reduction tables
        for this batch geometry.  Built once: the flat gradient buffer and the workspace do not move."""
        key = (ctx["n"], ctx["hw"], self.grads.data_ptr())
        if key in self._wgrad_plans:
            return self._wgrad_plans[key]
        f = self.encoder.features
        stage_of = {id(blk): si for si, layer in enumerate((f[4], f[5], f[6], f[7])) for blk in layer}
        items, total = [], 0
        for blk, rec in zip(self.blocks, ctx["blocks"]):
            shapes = [(conv, tuple(x.shape)) for conv, _, x, _, _, _ in rec["convs"]]
            if "down" in rec:
                shapes.append((rec["down"][0], tuple(rec["a_in"].shape)))
            for conv, (n, h, w, cin) in shapes:
                need = _lib.call("peclr_conv2d_wgrad_workspace_bytes", n, h, w, cin, conv.cout, conv.k, conv.stride)
                ks = _lib.call("peclr_conv2d_wgrad_splits", n, h, w, cin, conv.cout, conv.k, conv.stride)
                if need < 0 or ks < 1:
                    raise _lib.PeclrKernelError("weight gradient: unsupported geometry")
                items.append((stage_of[id(blk)], conv, need, ks, total))
                total += (need + 255) // 256 * 256
        buf = torch.empty((max(total, 256),), dtype=torch.uint8, device=self.device)
        region, per_stage = {}, {0: [], 1: [], 2: [], 3: []}
        for stage, conv, need, ks, off in items:
            region[id(conv)] = buf[off:off + need] if need else None
            if need:
                s = self._seg(conv)
                per_stage[stage].append((buf.data_ptr() + off, self.grads.data_ptr() + 4 * s.begin, s.size, ks))
        tables = {st: ops.build_reduce_table(ent, self.device) for st, ent in per_stage.items()}
        plan = {"buf": buf, "region": region, "tables": tables, "bytes": total}
        self._wgrad_plans = {key: plan}  # (one geometry at a time: the workspace is GBs for the deep trunks)
        return plan

    def _max_wgrad_bytes(self, ctx):
        key = (ctx["n"], ctx["hw"])
        cache = self.__dict__.setdefault("_wgrad_bytes_cache", {})
        if key not in cache:
            need = 0
            for rec in ctx["blocks"]:
                shapes = [(tuple(x.shape), conv.cout, conv.k, conv.stride) for conv, _, x, _, _, _ in rec["convs"]]
                if "down" in rec:
                    shapes.append((tuple(rec["a_in"].shape), rec["down"][0].cout, 1, rec["down"][0].stride))
                for (n, h, w, cin), cout, k, stride in shapes:
                    need = max(need, _lib.call("peclr_conv2d_wgrad_workspace_bytes", n, h, w, cin, cout, k, stride))
            cache[key] = need
        return cache[key]

    # ------------------------------------------------------------------ optimiser plumbing
    def optimizer_tables(self, weight_decay):
        if self._opt_tables is None or self._opt_tables["wd"] != weight_decay:
            t = ops.build_opt_tables([s.size for s in self.segs], self.weight_decay_table(weight_decay), self.device)
            t["wd"] = weight_decay
            self._opt_tables = t
        return self._opt_tables

    def optimizer_step(self, lr, step, weight_decay, lars=True, betas=(0.9, 0.999), eps=1e-8):
        self._require_cuda()
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        if self.w_bf16 is None:
            self.sync_weights()
        ops.lars_adam_step(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.optimizer_tables(weight_decay),
                           lr, step, p_bf16=self.w_bf16, lars=lars, betas=betas, adam_eps=eps)
        self.refresh_derived_weights()
        self.weights_dirty = False

    def zero_grad(self):
        self.grads.zero_()

    def attach_grads(self):
        """Re-bind .grad views (e.g. after someone called zero_grad(set_to_none=True) on a parameter)."""
        for s in self.segs:
            p = s.module._parameters[s.pname]
            if p.grad is None:
                p.grad = self._view(self.grads, s)
