"""Fused training steps over the engines (no autograd): the supervised step of third_party/unimatch/supervised.py:273-289
(BASELINE config 2) and the SemiVL weak-to-strong consistency step of semivl.py:224-346, with the flat-buffer AdamW of
semivl.py:326-328 / experiments.py:246-255 and the data-parallel gradient all-reduce of semivl.py:139-140.

All trainable tensors (backbone attn.* + pos_embed, every decode-head tensor) live in ONE flat fp32
buffer; gradients, Adam moments likewise, so the optimizer is two kernel launches (one per learning-rate class) and the
data-parallel exchange is a handful of NCCL all-reduces of contiguous slices, issued as the slices become final (GradExchange)."""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as L
from . import ops


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat):
    """Data-parallel gradient exchange of a (slice of the) flat gradient buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if _world() > 1:
        dist.all_reduce(flat)
    return flat


class GradExchange:
    """Bucketed data-parallel gradient exchange over the flat gradient buffer, overlapped with the rest of the backward pass.

    The reference wraps the model in DDP (semivl.py:139-140), whose reducer all-reduces gradient buckets while autograd is
    still running.  Here the backward pass is a hand-scheduled kernel sequence, so the schedule is explicit: `reduce(lo, hi)`
    is called as soon as the slice [lo, hi) of the flat buffer is final (the head after `head.backward`, groups of encoder
    layers from the last one down); it makes a side stream wait for the kernels issued so far and launches the NCCL
    all-reduce there, so the transfer over NVLink runs under the remaining backward kernels.  `finish()` reduces whatever was
    not covered and makes the compute stream wait for the side stream before AdamW.  Sums only: the 1/world mean is folded
    into the AdamW kernel's `gscale`."""

    def __init__(self, flat):
        self.flat, self.done = flat, []
        self.comm = torch.cuda.Stream(device=flat.device) if flat.is_cuda and _world() > 1 else None

    def begin(self):
        self.done = []

    def reduce(self, lo, hi):
        if _world() == 1 or hi <= lo:
            return
        assert all(hi <= a or lo >= b for a, b in self.done), "gradient slice exchanged twice"
        self.done.append((lo, hi))
        if self.comm is None:
            allreduce_sum_(self.flat[lo:hi])
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            allreduce_sum_(self.flat[lo:hi])

    def finish(self):
        pos = 0
        for a, b in sorted(self.done) + [(self.flat.numel(), self.flat.numel())]:
            self.reduce(pos, a)
            pos = b
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.done = []


CONF_MODES = {"pixelwise": 0, "pixelratio": 1, "pixelavg": 2}


class OptimCfg:
    """experiments.py:246-255: AdamW, one learning-rate class per key of `paramwise_cfg.custom_keys` that reaches a trainable tensor
    (backbone, head, conv_encoder), uniform weight decay; poly decay over `scheduler_max_iters` (default: total_iters) after an optional
    linear warm-up (semivl.py:338-345); `total_iters` also drives the MaskCLIP-lambda ramp (semivl.py:312-316)."""

    def __init__(self, lr=1e-4, weight_decay=0.01, backbone_lr_mult=0.01, head_lr_mult=10.0, betas=(0.9, 0.999), eps=1e-8,
                 total_iters=1000, power=0.9, conv_encoder_lr_mult=0.1, warmup_iters=0, warmup_ratio=1e-6, scheduler_max_iters=None):
        self.lr, self.wd, self.bb_mult, self.head_mult, self.ce_mult = lr, weight_decay, backbone_lr_mult, head_lr_mult, conv_encoder_lr_mult
        self.betas, self.eps, self.total_iters, self.power = betas, eps, total_iters, power
        self.warmup_iters, self.warmup_ratio = warmup_iters, warmup_ratio
        self.scheduler_max_iters = scheduler_max_iters if scheduler_max_iters is not None else total_iters


class Trainer:
    def __init__(self, model, optim=None, hp=None, head_chunk_bytes=24e9):
        self.model = model
        # activation budget of ONE gradient-tracked head pass of the SemiVL step (see _head_chunk); None = never split
        self.head_chunk_bytes = head_chunk_bytes
        self.opt = optim or OptimCfg()
        self.hp = dict(conf_thresh=0.95, conf_mode="pixelwise", mcc_conf_thresh=0.9, mcc_loss_reduce="mean_all", mcc_lambda=0.1)
        if hp:
            self.hp.update(hp)
        assert self.hp["conf_mode"] in CONF_MODES, self.hp["conf_mode"]                       # utils/train_utils.py:36-48
        assert self.hp["mcc_loss_reduce"] in ("mean", "mean_valid", "mean_all"), self.hp["mcc_loss_reduce"]   # semivl.py:113
        self.iters = 0
        bb = [(n, p) for n, p in model.backbone.named_parameters() if p.requires_grad]
        hd = [(n, p) for n, p in model.decode_head.named_parameters() if p.requires_grad]
        conv = getattr(model, "conv_encoder", None)
        ce = [(n, p) for n, p in conv.named_parameters() if p.requires_grad] if conv is not None else []
        self.n_bb = sum(p.numel() for _, p in bb)
        self.n_hd = sum(p.numel() for _, p in hd)
        self.n_ce = sum(p.numel() for _, p in ce)
        dev = next(model.parameters()).device
        n = self.n_bb + self.n_hd + self.n_ce
        self.p_flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.g_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.g_bb, self.g_hd, self.g_ce = {}, {}, {}
        off = 0
        for group, gd in ((bb, self.g_bb), (hd, self.g_hd), (ce, self.g_ce)):
            for name, p in group:
                k = p.numel()
                self.p_flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.p_flat[off:off + k].view(p.shape)            # parameters become views of the flat buffer
                gd[name] = self.g_flat[off:off + k].view(p.shape)
                off += k
        if _world() > 1:
            # DDP broadcasts rank 0's module state at construction (semivl.py:139-140): replicas must start from the same parameters
            # whatever each rank's RNG produced for the randomly initialised head; the Adam moments are zeros on every rank already
            dist.broadcast(self.p_flat, src=0)
        self.vit, self.head = model.backbone.engine, model.decode_head.engine
        self.ce = conv.engine if conv is not None else None
        self.vit.cache.volatile = set(self.g_bb)
        self.head.cache.volatile = set(self.g_hd)
        if self.ce is not None:
            self.ce.cache.volatile = set(self.g_ce)
        for e in (self.vit, self.head):
            e.cache.batched = not e.precise       # throughput mode: operand copies / staged weight gradients move in one launch per step
        self.world = _world()
        self._capturing, self._graph, self._graph_key, self._hyper = False, None, None, None
        self.exchange = GradExchange(self.g_flat)
        # exchange schedule of the encoder gradients: layers are contiguous and ascending in the flat buffer, the backward pass
        # runs from the last layer down -> one bucket per `bucket_layers` layers, cut at the first parameter of a layer
        self.layer_lo, self.bucket_layers = {}, 3
        off, self._layers_end, order, tail = 0, 0, [], False
        for name, p in bb:
            if name.startswith("layers."):
                assert not tail, f"{name}: the encoder layers must be contiguous in the flat buffer"
                li = int(name.split(".")[1])
                self.layer_lo.setdefault(li, off)
                order.append(li)
                self._layers_end = off + p.numel()
            else:
                tail = bool(order)
            off += p.numel()
        assert order == sorted(order), "encoder layers must be laid out in ascending order in the flat buffer"
        self._bucket_hi = self._layers_end
        self._f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)

    # ------------------------------------------------------------------ helpers
    def _pb(self):
        return {n: p.data for n, p in self.model.backbone.named_parameters()}

    def _ph(self):
        return {n: p.data for n, p in self.model.decode_head.named_parameters()}

    def _pc(self):
        """conv-encoder parameters and BatchNorm buffers (running statistics are updated in place by training-mode passes)"""
        c = self.model.conv_encoder
        d = {n: p.data for n, p in c.named_parameters()}
        d.update({n: b for n, b in c.named_buffers() if "num_batches_tracked" not in n})
        return d

    def lr_at(self, it):
        """Base learning-rate FACTOR times lr used BY optimizer step number `it` (1-based).  semivl.py:338-345 rewrites the rate after each
        optimizer.step() from the 0-based index i of the iteration just run -- linear warm-up `1 - (1 - i / warmup_iters) * (1 - warmup_ratio)`
        while i < warmup_iters, else poly `(1 - i / scheduler_max_iters) ** 0.9` -- so step 1 runs at the initial rate and step `it` >= 2 at
        the rate written after iteration i = it - 2."""
        o = self.opt
        if it <= 1:
            return o.lr
        i = it - 2
        if i < o.warmup_iters:
            return o.lr * (1.0 - (1.0 - i / o.warmup_iters) * (1.0 - o.warmup_ratio))
        return o.lr * (1.0 - i / o.scheduler_max_iters) ** o.power

    def _head_grads_final(self):
        self.exchange.reduce(self.n_bb, self.n_bb + self.n_hd)

    def _conv_grads_final(self):
        self.exchange.reduce(self.n_bb + self.n_hd, self.n_bb + self.n_hd + self.n_ce)

    def _layer_grads_final(self, i):
        """called by the encoder backward after layer i's weight gradients are complete (layers run 11 -> 0)"""
        if i in self.layer_lo and i > 0 and i % self.bucket_layers == 0:
            self.exchange.reduce(self.layer_lo[i], self._bucket_hi)
            self._bucket_hi = self.layer_lo[i]

    def _step_scalars(self, it):
        """(lr of the backbone class, lr of the head class, 1 - beta1^t, sqrt(1 - beta2^t), lr of the conv-encoder class) of optimizer step
        number `it` (1-based) -- the device vector svl_adamw_dev reads (lr_index 0, 1 and 4)."""
        o = self.opt
        base = self.lr_at(it)
        return base * o.bb_mult, base * o.head_mult, 1.0 - o.betas[0] ** it, (1.0 - o.betas[1] ** it) ** 0.5, base * o.ce_mult

    def optimizer_step(self):
        self.exchange.finish()                                        # sum over ranks; the mean is folded into gscale below
        self._bucket_hi = self._layers_end
        o, gs, nb = self.opt, 1.0 / self.world, self.n_bb
        classes = [(0, nb, 0, o.bb_mult), (nb, self.n_hd, 1, o.head_mult)]          # (offset, length, lr_index of svl_adamw_dev, lr multiplier)
        if self.n_ce:
            classes.append((nb + self.n_hd, self.n_ce, 4, o.ce_mult))
        if self._capturing:
            # inside a CUDA-graph capture: per-step scalars come from device memory (self._hyper), host counters move at replay time
            for lo, k, idx, _ in classes:
                L.call("svl_adamw_dev", self.p_flat[lo:], self.g_flat[lo:], self.m_flat[lo:], self.v_flat[lo:], k, self._hyper, idx, o.betas[0],
                       o.betas[1], o.eps, o.wd, gs)
            self._refresh_operands()
            return
        self.iters += 1
        lr = self.lr_at(self.iters)
        for lo, k, _, mult in classes:
            L.call("svl_adamw", self.p_flat[lo:], self.g_flat[lo:], self.m_flat[lo:], self.v_flat[lo:], k, lr * mult, o.betas[0], o.betas[1],
                   o.eps, o.wd, self.iters, gs)
        self._bump_caches()
        self._refresh_operands()

    def _refresh_operands(self):
        for e in (self.vit, self.head):
            e.cache.refresh()

    def _bump_caches(self):
        for e in (self.vit, self.head, self.ce):
            if e is not None:
                e.cache.bump()

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Optimizer-side state for checkpoint / resume: the flat parameter, first- and second-moment buffers and the step counter (the
        reference saves `optimizer.state_dict()` next to the model, semivl.py:423-433)."""
        return dict(p_flat=self.p_flat.clone(), m_flat=self.m_flat.clone(), v_flat=self.v_flat.clone(), iters=self.iters,
                    layout=dict(n_bb=self.n_bb, n_hd=self.n_hd, n_ce=self.n_ce))

    def load_state_dict(self, sd):
        lay = sd["layout"]
        assert (lay["n_bb"], lay["n_hd"], lay["n_ce"]) == (self.n_bb, self.n_hd, self.n_ce), "checkpoint belongs to a different trainable-parameter layout"
        self.p_flat.copy_(sd["p_flat"])
        self.m_flat.copy_(sd["m_flat"])
        self.v_flat.copy_(sd["v_flat"])
        self.iters = int(sd["iters"])
        self.invalidate_weights()

    def invalidate_weights(self):
        """Call after ANY in-place edit of the model's weights that did not go through this trainer (model.load_state_dict, manual surgery):
        the parameters are `.data` views of the flat buffer, whose version counter never moves, so the engines' cached bf16 operand copies
        -- of frozen tensors too -- must be dropped by hand; a captured CUDA graph holds casts of the old values and is dropped as well."""
        for e in (self.vit, self.head, self.ce):
            if e is not None:
                e.cache.clear()
        self._graph, self._graph_key = None, None

    # ------------------------------------------------------------------ CUDA-graph replay of the supervised step
    def graphed_supervised_step(self, img, mask):
        """`supervised_step(img, mask)` as ONE CUDA-graph launch (~690 kernel launches and their host-side descriptor set-up
        collapse into a replay).  Captured on first use for the given input shapes; inputs are copied into the graph's static
        buffers, the per-step AdamW scalars (LR schedule, bias corrections) into a device vector.  With more than one rank the
        bucketed NCCL all-reduces of GradExchange are captured too: the side stream forks from and re-joins the capturing stream,
        so the replayed graph keeps the exchange overlapped with the backward kernels."""
        key = (tuple(img.shape), tuple(mask.shape))
        if self._graph is None or self._graph_key != key:
            self._g_img, self._g_mask = img.clone(), mask.clone()
            self._hyper = torch.zeros(5, device=img.device)
            self._hyper_host = [torch.zeros(5, dtype=torch.float32).pin_memory() for _ in range(8)]
            self._hyper_done = [torch.cuda.Event() for _ in range(8)]
            self._hyper_i = 0
            self.supervised_step(self._g_img, self._g_mask, update=False)        # eager pass: frozen-weight operand cache, lazy attributes
            self._bump_caches()                                                   # trainable-weight casts must be part of the graph
            self._refresh_operands()                                              # batched mode: job tables built (host -> device copies) before capture
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            self._capturing, l0 = True, L.launches
            try:
                with torch.cuda.graph(self._graph):
                    self._g_loss = self.supervised_step(self._g_img, self._g_mask, update=True)
            finally:
                self._capturing = False
            self._graph_key, self._graph_launches = key, L.launches - l0      # svl kernels recorded into the graph
            L.launches = l0
        self._g_img.copy_(img, non_blocking=True)
        self._g_mask.copy_(mask, non_blocking=True)
        # the step's scalars travel through a ring of PINNED host slots: a copy from pageable memory would block the host until everything
        # queued before it (the previous step) has run, i.e. one graph launch latency of GPU idle time per step
        slot = self._hyper_i % len(self._hyper_host)
        self._hyper_i += 1
        self._hyper_done[slot].synchronize()                       # the copy that last read this slot has run (long ago)
        self._hyper_host[slot].copy_(torch.tensor(self._step_scalars(self.iters + 1), dtype=torch.float32))
        self._hyper.copy_(self._hyper_host[slot], non_blocking=True)
        self._hyper_done[slot].record()
        self._graph.replay()
        L.launches += self._graph_launches
        self.iters += 1
        self._bump_caches()                         # the cached operand copies now belong to the graph: eager calls must re-cast
        return self._g_loss

    @staticmethod
    def _ptr_array(items):
        return (C.c_void_p * 3)(*[(t.data_ptr() if t is not None else None) for t in (list(items) + [None] * 3)[:3]])

    def _ce(self, low_rows, dlow_rows, R, N, hl, wl, H, W, targets, loss_out):
        """targets: list of (labels int64 [R,H,W], weights f32 or None, coef device scalar)."""
        labels = self._ptr_array([t[0] for t in targets])
        weights = self._ptr_array([t[1] for t in targets])
        coefs = self._ptr_array([t[2] for t in targets])
        L.call("svl_upsample_ce", low_rows, dlow_rows, R, N, hl, wl, H, W, len(targets), labels, weights, coefs, loss_out, 1.0, 255)

    # ------------------------------------------------------------------ supervised step (BASELINE config 2)
    def supervised_step(self, img, mask, update=True):
        """model(img) -> CE(ignore 255) -> backward -> AdamW.  Returns the loss as a device scalar (no host sync)."""
        m = self.model
        H, W = img.shape[-2:]
        text = m._text(img.device)
        self.g_flat.zero_()
        self.exchange.begin()
        pb, ph = self._pb(), self._ph()
        feats, _, vctx = self.vit.forward(m.renormalize_img_for_clip(img), pb, need_grad=True, want_global=False)
        conv_feats = cctx = None
        if self.ce is not None:                    # skr04: the conv encoder sees the ImageNet-normalised image (model/vlm.py:113,120-121)
            pc, csync = self._pc(), m.conv_encoder.sync
            cf, cctx = self.ce.forward(img, pc, training=True, need_grad=True, sync=csync)
            conv_feats = [cf]
        low, hctx = self.head.forward(feats, text, ph, need_grad=True, conv_feats=conv_feats)
        R, N, hl, wl = low.shape
        cnt, coef, loss = self._f(1), self._f(1), self._f(3)
        L.call("svl_count_valid", mask, mask.numel(), 255, cnt)
        L.call("svl_reciprocal", cnt, coef, 1.0, 1.0)
        d_low = torch.zeros_like(low)
        self._ce(low, d_low, R, N, hl, wl, H, W, [(mask, None, coef)], loss)
        dfe = self.head.backward(hctx, d_low, ph, self.g_hd)
        del hctx
        if update:
            self._head_grads_final()
        if cctx is not None:                       # [taps..., embedding, conv feature]: the conv encoder's backward does not depend on the ViT's
            self.ce.backward(cctx, dfe[len(feats)], pc, self.g_ce, sync=csync)
            del cctx
            if update:
                self._conv_grads_final()
            dfe = dfe[:len(feats)]
        self.vit.backward(vctx, dfe, pb, self.g_bb, on_layer_done=self._layer_grads_final if update else None)
        del vctx
        if update:
            self.optimizer_step()
        return loss[0]

    # ------------------------------------------------------------------ SemiVL step (semivl.py:224-346)
    def _head_chunk(self, N, hl, wl):
        """Images per gradient-tracked head pass.  The head's saved activations are per (image, class) map -- ~16 MB per map at 128 x 128
        output pixels in bf16 (twice that in the split-operand mode), i.e. 2.4 GB per image at N = 150 -- and no op of
        the head mixes images (GroupNorm per map, class attention per image position), so the head batch can be cut into groups of images
        that run forward -> loss -> backward one after the other with exactly the same result: the peak is one group's activations
        instead of the whole 4b-image head batch (BASELINE configs 4 / 5: 124 GB at ADE b=8, 235 GB predicted at COCO b=16 unsplit)."""
        if not self.head_chunk_bytes:
            return 1 << 30
        per_img = N * 16e6 * (hl * wl) / (128.0 * 128.0) * (2.0 if self.head.precise else 1.0)
        return max(1, int(self.head_chunk_bytes // per_img))

    def semivl_step(self, batch, drop_masks=None, update=True):
        """batch keys follow semivl.py:203-221.  One encoder pass over (img_x | img_w | img_s1 | img_s2); a no-grad head pass over the
        weak views (their prediction is detached, semivl.py:251: it only yields pseudo-labels) and gradient-tracked head passes over
        (x | w_fp | s1 | s2), cut into image groups by _head_chunk -- the perturbed copy of the labelled images is never computed (the
        reference discards it, semivl.py:247) --; a no-grad teacher pass on img_w_other and the frozen MaskCLIP pass on (img_w | img_w_other)."""
        m, hp = self.model, self.hp
        b = batch["img_x"].shape[0]
        H, W = batch["img_x"].shape[-2:]
        dev = batch["img_x"].device
        text = m._text(dev)
        pb, ph = self._pb(), self._ph()
        lam = hp["mcc_lambda"]
        if isinstance(lam, (list, tuple)):                               # linear schedule of the MaskCLIP lambda (semivl.py:312-316)
            prog = self.iters / self.opt.total_iters
            lam = lam[0] * (1 - prog) + lam[1] * prog
        self.g_flat.zero_()
        self.exchange.begin()
        img_s1 = torch.empty_like(batch["img_s1"])
        img_s2 = torch.empty_like(batch["img_s2"])
        L.call("svl_cutmix_img", batch["img_s1"], batch["img_s1_other"], batch["mix1"], img_s1, b, 3, H * W)
        L.call("svl_cutmix_img", batch["img_s2"], batch["img_s2_other"], batch["mix2"], img_s2, b, 3, H * W)
        # ---- teacher passes (no grad)
        fo, _, _ = self.vit.forward(m.renormalize_img_for_clip(batch["img_w_other"]), pb, need_grad=False, want_global=False)
        cf_o = None
        if self.ce is not None:                    # model.eval() for the teacher pass (semivl.py:228-232): running BatchNorm statistics
            pc, csync = self._pc(), m.conv_encoder.sync
            cf_o, _ = self.ce.forward(batch["img_w_other"], pc, training=False, need_grad=False)
        N, hl, wl = text.shape[0], 4 * fo[-1].shape[1], 4 * fo[-1].shape[2]
        chunk = self._head_chunk(N, hl, wl)
        conf_o = torch.empty(b, H, W, device=dev)
        lab_o = torch.empty(b, H, W, device=dev, dtype=torch.int64)
        for i0 in range(0, b, 4 * chunk):                           # nothing is saved in a no-grad pass: 4x larger groups
            i1 = min(b, i0 + 4 * chunk)
            low_o, _ = self.head.forward([f[i0:i1] for f in fo], text, ph, need_grad=False, conv_feats=None if cf_o is None else [cf_o[i0:i1]])
            L.call("svl_softmax_max", low_o, conf_o[i0:i1], lab_o[i0:i1], i1 - i0, N, hl, wl, H, W, 1.0, 0.0)
            del low_o
        del fo, cf_o
        mclip = mclip_o = None
        if lam != 0:
            mc = m.forward_maskclip(torch.cat((batch["img_w"], batch["img_w_other"])), hp["mcc_conf_thresh"])
            mclip, mclip_o = mc[:b], mc[b:]
            mclip = torch.where(batch["ignore_mask"] == 255, 255, mclip)
            mclip_o = torch.where(batch["ignore_mask_other"] == 255, 255, mclip_o)
        # ---- student passes
        imgs = torch.cat((batch["img_x"], batch["img_w"], img_s1, img_s2))
        feats, _, vctx = self.vit.forward(m.renormalize_img_for_clip(imgs), pb, need_grad=True, want_global=False)
        cf = cctx = None
        if self.ce is not None:
            # two training-mode calls like the reference's two student forwards (semivl.py:245-249): the BatchNorm statistics of (x | w) and
            # of (s1 | s2) are separate batches (and the running statistics move twice per step)
            cfa, ctx_a = self.ce.forward(imgs[:2 * b], pc, training=True, need_grad=True, sync=csync)
            cfb, ctx_b = self.ce.forward(imgs[2 * b:], pc, training=True, need_grad=True, sync=csync)
            cf, cctx = torch.cat((cfa, cfb)), (ctx_a, ctx_b)
            del cfa, cfb
        nfe = len(feats)
        pert = list(feats) + ([cf] if cf is not None else [])          # every feature the perturbed pass drops channels of (builder.py:78-85)
        if drop_masks is None:
            drop_masks = [torch.bernoulli(torch.full((b, f.shape[-1]), 1.0 - m.fp_rate, device=dev)) for f in pert]
        scale = 1.0 / (1.0 - m.fp_rate)
        dm = [(dmk.reshape(b, 1, 1, -1).to(dev) * scale) for dmk in drop_masks]
        # pseudo-labels of the weak views (pred_w.detach().softmax.max, semivl.py:251-252): no gradient ever reaches these rows
        conf_w = torch.empty(b, H, W, device=dev)
        lab_w = torch.empty(b, H, W, device=dev, dtype=torch.int64)
        for i0 in range(0, b, 4 * chunk):                           # nothing is saved in a no-grad pass: 4x larger groups
            i1 = min(b, i0 + 4 * chunk)
            low_w, _ = self.head.forward([f[b + i0:b + i1] for f in feats], text, ph, need_grad=False,
                                         conv_feats=None if cf is None else [cf[b + i0:b + i1]])
            L.call("svl_softmax_max", low_w, conf_w[i0:i1], lab_w[i0:i1], i1 - i0, N, hl, wl, H, W, 1.0, 0.0)
            del low_w
        # ---- targets (cutmix of pseudo-labels, confidences, ignore masks; confidence weights per cfg['conf_mode'])
        npx = float(b * H * W)
        mode = CONF_MODES[hp["conf_mode"]]
        def coef(count, numer):
            out = self._f(1)
            L.call("svl_reciprocal", count, out, numer, 1.0)
            return out

        def conf_target(lab_b, conf_b, ign_b, box, numer, want_lab):
            """(labels, per-pixel weights | None, device coefficient, #valid of the mixed ignore mask) of one consistency loss:
            confidence_weighted_loss(CE(pred, cutmix(mask_w, lab_b)), cutmix(conf_w, conf_b), cutmix(ignore, ign_b)) * numer."""
            lab = torch.empty_like(lab_w) if want_lab else None
            wgt = torch.empty_like(conf_w) if mode == 0 else None
            cnt = self._f(1)
            L.call("svl_cutmix_weights", lab_w, lab_b, conf_w, conf_b, batch["ignore_mask"], ign_b, box, lab, wgt, None, cnt, lab_w.numel(),
                   hp["conf_thresh"])
            if mode == 0:
                return lab, wgt, coef(cnt, numer), cnt
            stats, cf = self._f(b, 3), self._f(1)
            row_w = self._f(b) if mode == 1 else None
            L.call("svl_conf_stats", conf_w, conf_b, batch["ignore_mask"], ign_b, box, stats, b, H * W, hp["conf_thresh"])
            L.call("svl_conf_coef", stats, b, mode, numer, cf, row_w)
            if mode == 1:
                wgt = torch.empty_like(conf_w)
                L.call("svl_fill_rows", wgt, row_w, b, H * W)
            return lab, wgt, cf, cnt

        def mc_coef(mcl, cnt_ign, k):
            """device coefficient of one MaskCLIP consistency term: lambda * k / denominator(mcc_loss_reduce) (semivl.py:52-58)."""
            red = hp["mcc_loss_reduce"]
            if red == "mean_all":
                return one * (lam * k / npx)
            if red == "mean_valid":
                return coef(cnt_ign, lam * k)
            c = self._f(1)                                                  # 'mean': CrossEntropyLoss(ignore_index=255) mean over labels != 255
            L.call("svl_count_valid", mcl, mcl.numel(), 255, c)
            return coef(c, lam * k)

        one = torch.ones(1, device=dev)
        tgt = {}
        for key, box in (("s1", batch["mix1"]), ("s2", batch["mix2"])):
            lab, wgt, cf, cnt = conf_target(lab_o, conf_o, batch["ignore_mask_other"], box, 0.125, True)
            mcl = None
            if lam != 0:
                mcl = torch.empty_like(lab_w)
                L.call("svl_cutmix_weights", mclip, mclip_o, None, None, None, None, box, mcl, None, None, None, mcl.numel(), 0.0)
            tgt[key] = (lab, wgt, cf, cnt, mcl)
        _, w_fp, cf_fp, cnt_fp = conf_target(lab_w, conf_w, batch["ignore_mask"], None, 0.25, False)
        cnt_x = self._f(1)
        L.call("svl_count_valid", batch["mask_x"], batch["mask_x"].numel(), 255, cnt_x)
        # loss groups: (encoder rows, feature scale, targets, loss accumulator); every coefficient is a full-batch quantity on the device
        lx, lfp = self._f(3), self._f(3)
        t_fp = [(lab_w, w_fp, cf_fp)]
        if lam != 0:
            t_fp.append((mclip, None, mc_coef(mclip, cnt_fp, 0.5)))
        groups = [(0, None, [(batch["mask_x"], None, coef(cnt_x, 0.5))], lx), (b, dm, t_fp, lfp)]
        ls = {}
        for r0, key in ((2 * b, "s1"), (3 * b, "s2")):
            lab, wgt, cf, cnt, mcl = tgt[key]
            t = [(lab, wgt, cf)]
            if lam != 0:
                t.append((mcl, None, mc_coef(mcl, cnt, 0.25)))
            ls[key] = self._f(3)
            groups.append((r0, None, t, ls[key]))
        # ---- gradient-tracked head passes, one image group at a time: forward -> fused upsample + CE (loss and d_low) -> backward
        dfe = [torch.zeros(4 * b, *f.shape[1:], device=dev, dtype=torch.float32) for f in pert]
        sl = lambda t, i0, i1: None if t is None else t[i0:i1]
        for r0, scale_k, targets, acc in groups:
            for i0 in range(0, b, chunk):
                i1 = min(b, i0 + chunk)
                hf = [f[r0 + i0:r0 + i1] for f in pert]
                if scale_k is not None:                     # feature perturbation of the weak views (builder.py:78-85)
                    hf = [f * k[i0:i1] for f, k in zip(hf, scale_k)]
                low, hctx = self.head.forward(hf[:nfe], text, ph, need_grad=True, conv_feats=hf[nfe:] or None)
                del hf
                d_low = torch.zeros_like(low)
                self._ce(low, d_low, i1 - i0, N, hl, wl, H, W, [(sl(lb, i0, i1), sl(wg, i0, i1), cf) for lb, wg, cf in targets], acc)
                dh = self.head.backward(hctx, d_low, ph, self.g_hd)
                del hctx, low, d_low
                for j, (g, d) in enumerate(zip(dfe, dh)):   # the clean weak views get no gradient: their rows carry the perturbed copy's
                    g[r0 + i0:r0 + i1] = d if scale_k is None else d * scale_k[j][i0:i1]
                del dh
        total = lx[0] + lfp[0] + lfp[1] + ls["s1"][0] + ls["s1"][1] + ls["s2"][0] + ls["s2"][1]
        terms = dict(loss_x=lx[0] * 2, loss_fp=lfp[0] * 4, loss_s1=ls["s1"][0] * 8, loss_s2=ls["s2"][0] * 8)
        if lam != 0:
            terms.update(loss_mc_fp=lfp[1] / (lam * 0.5), loss_mc_s1=ls["s1"][1] / (lam * 0.25), loss_mc_s2=ls["s2"][1] / (lam * 0.25))
        # ---- encoder backward
        if update:
            self._head_grads_final()
        if cctx is not None:
            d_cf = dfe.pop()
            self.ce.backward(cctx[0], d_cf[:2 * b], pc, self.g_ce, sync=csync)
            self.ce.backward(cctx[1], d_cf[2 * b:], pc, self.g_ce, sync=csync)
            del cctx, d_cf
            if update:
                self._conv_grads_final()
        self.vit.backward(vctx, dfe, pb, self.g_bb, on_layer_done=self._layer_grads_final if update else None)
        del vctx
        if update:
            self.optimizer_step()
        return total, terms
