"""Fused training steps over the engines (no autograd): the supervised step of third_party/unimatch/supervised.py:273-289
(BASELINE config 2) and the SemiVL weak-to-strong consistency step of semivl.py:224-346, with the flat-buffer AdamW of
semivl.py:326-328 / experiments.py:246-255 and the data-parallel gradient all-reduce of semivl.py:139-140.

All trainable tensors (backbone attn.* + pos_embed, every decode-head tensor) live in ONE flat fp32
buffer; gradients, Adam moments likewise, so the optimizer is two kernel launches (one per learning-rate class) and the
data-parallel exchange is one NCCL all-reduce."""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as L
from . import ops


def allreduce_sum_(flat):
    """Data-parallel gradient exchange: ONE all-reduce of the flat gradient buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
    return flat


class OptimCfg:
    def __init__(self, lr=1e-4, weight_decay=0.01, backbone_lr_mult=0.01, head_lr_mult=10.0, betas=(0.9, 0.999), eps=1e-8,
                 total_iters=1000, power=0.9):
        self.lr, self.wd, self.bb_mult, self.head_mult = lr, weight_decay, backbone_lr_mult, head_lr_mult
        self.betas, self.eps, self.total_iters, self.power = betas, eps, total_iters, power


class Trainer:
    def __init__(self, model, optim=None, hp=None):
        self.model = model
        self.opt = optim or OptimCfg()
        self.hp = dict(conf_thresh=0.95, conf_mode="pixelwise", mcc_conf_thresh=0.9, mcc_loss_reduce="mean_all", mcc_lambda=0.1)
        if hp:
            self.hp.update(hp)
        assert self.hp["conf_mode"] == "pixelwise" and self.hp["mcc_loss_reduce"] == "mean_all", \
            "the fused loss path implements conf_mode='pixelwise' and mcc_loss_reduce='mean_all' (VOC/COCO/ADE experiments)"
        self.iters = 0
        bb = [(n, p) for n, p in model.backbone.named_parameters() if p.requires_grad]
        hd = [(n, p) for n, p in model.decode_head.named_parameters() if p.requires_grad]
        self.n_bb = sum(p.numel() for _, p in bb)
        self.n_hd = sum(p.numel() for _, p in hd)
        dev = next(model.parameters()).device
        n = self.n_bb + self.n_hd
        self.p_flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.g_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.g_bb, self.g_hd = {}, {}
        off = 0
        for group, gd in ((bb, self.g_bb), (hd, self.g_hd)):
            for name, p in group:
                k = p.numel()
                self.p_flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.p_flat[off:off + k].view(p.shape)            # parameters become views of the flat buffer
                gd[name] = self.g_flat[off:off + k].view(p.shape)
                off += k
        self.vit, self.head = model.backbone.engine, model.decode_head.engine
        self.vit.cache.volatile = set(self.g_bb)
        self.head.cache.volatile = set(self.g_hd)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)

    # ------------------------------------------------------------------ helpers
    def _pb(self):
        return {n: p.data for n, p in self.model.backbone.named_parameters()}

    def _ph(self):
        return {n: p.data for n, p in self.model.decode_head.named_parameters()}

    def lr_now(self):
        """poly schedule applied after each step (semivl.py:338-345)"""
        return self.opt.lr * (1.0 - self.iters / self.opt.total_iters) ** self.opt.power if self.iters > 0 else self.opt.lr

    def optimizer_step(self):
        allreduce_sum_(self.g_flat)                                   # NCCL sum over ranks; the mean is folded into gscale below
        self.iters += 1
        lr, o = self.lr_now() if self.iters > 1 else self.opt.lr, self.opt
        gs = 1.0 / self.world
        nb = self.n_bb
        L.call("svl_adamw", self.p_flat, self.g_flat, self.m_flat, self.v_flat, nb, lr * o.bb_mult, o.betas[0], o.betas[1], o.eps, o.wd,
               self.iters, gs)
        L.call("svl_adamw", self.p_flat[nb:], self.g_flat[nb:], self.m_flat[nb:], self.v_flat[nb:], self.n_hd, lr * o.head_mult, o.betas[0],
               o.betas[1], o.eps, o.wd, self.iters, gs)
        self.vit.cache.bump()
        self.head.cache.bump()

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
        pb, ph = self._pb(), self._ph()
        feats, _, vctx = self.vit.forward(m.renormalize_img_for_clip(img), pb, need_grad=True, want_global=False)
        low, hctx = self.head.forward(feats, text, ph, need_grad=True)
        R, N, hl, wl = low.shape
        cnt, coef, loss = self._f(1), self._f(1), self._f(3)
        L.call("svl_count_valid", mask, mask.numel(), 255, cnt)
        L.call("svl_reciprocal", cnt, coef, 1.0, 1.0)
        d_low = torch.zeros_like(low)
        self._ce(low, d_low, R, N, hl, wl, H, W, [(mask, None, coef)], loss)
        dfe = self.head.backward(hctx, d_low, ph, self.g_hd)
        del hctx
        self.vit.backward(vctx, dfe, pb, self.g_bb)
        del vctx
        if update:
            self.optimizer_step()
        return loss[0]

    # ------------------------------------------------------------------ SemiVL step (semivl.py:224-346)
    def semivl_step(self, batch, drop_masks=None, update=True):
        """batch keys follow semivl.py:203-221.  One encoder pass over (img_x | img_w | img_s1 | img_s2), one head pass over
        (x | w | w_fp | s1 | s2) -- the perturbed copy of the labelled images is never computed (the reference discards it,
        semivl.py:247) -- a no-grad teacher pass on img_w_other and the frozen MaskCLIP pass on (img_w | img_w_other)."""
        m, hp = self.model, self.hp
        b = batch["img_x"].shape[0]
        H, W = batch["img_x"].shape[-2:]
        dev = batch["img_x"].device
        text = m._text(dev)
        pb, ph = self._pb(), self._ph()
        lam = hp["mcc_lambda"]
        self.g_flat.zero_()
        img_s1 = torch.empty_like(batch["img_s1"])
        img_s2 = torch.empty_like(batch["img_s2"])
        L.call("svl_cutmix_img", batch["img_s1"], batch["img_s1_other"], batch["mix1"], img_s1, b, 3, H * W)
        L.call("svl_cutmix_img", batch["img_s2"], batch["img_s2_other"], batch["mix2"], img_s2, b, 3, H * W)
        # ---- teacher passes (no grad)
        fo, _, _ = self.vit.forward(m.renormalize_img_for_clip(batch["img_w_other"]), pb, need_grad=False, want_global=False)
        low_o, _ = self.head.forward(fo, text, ph, need_grad=False)
        del fo
        N, hl, wl = low_o.shape[1:]
        conf_o = torch.empty(b, H, W, device=dev)
        lab_o = torch.empty(b, H, W, device=dev, dtype=torch.int64)
        L.call("svl_softmax_max", low_o, conf_o, lab_o, b, N, hl, wl, H, W, 1.0, 0.0)
        mclip = mclip_o = None
        if lam != 0:
            mc = m.forward_maskclip(torch.cat((batch["img_w"], batch["img_w_other"])), hp["mcc_conf_thresh"])
            mclip, mclip_o = mc[:b], mc[b:]
            mclip = torch.where(batch["ignore_mask"] == 255, 255, mclip)
            mclip_o = torch.where(batch["ignore_mask_other"] == 255, 255, mclip_o)
        # ---- student passes
        imgs = torch.cat((batch["img_x"], batch["img_w"], img_s1, img_s2))
        feats, _, vctx = self.vit.forward(m.renormalize_img_for_clip(imgs), pb, need_grad=True, want_global=False)
        if drop_masks is None:
            drop_masks = [torch.bernoulli(torch.full((b, f.shape[-1]), 1.0 - m.fp_rate, device=dev)) for f in feats]
        scale = 1.0 / (1.0 - m.fp_rate)
        dm = [(dmk.reshape(b, 1, 1, -1).to(dev) * scale) for dmk in drop_masks]
        # head batch: [x | w | w_fp | s1 | s2]
        hf = [torch.cat((f[:2 * b], f[b:2 * b] * k, f[2 * b:])) for f, k in zip(feats, dm)]
        low, hctx = self.head.forward(hf, text, ph, need_grad=True)
        del hf
        conf_w = torch.empty(b, H, W, device=dev)
        lab_w = torch.empty(b, H, W, device=dev, dtype=torch.int64)
        L.call("svl_softmax_max", low[b:2 * b], conf_w, lab_w, b, N, hl, wl, H, W, 1.0, 0.0)
        # ---- targets (cutmix of pseudo-labels, confidences, ignore masks; confidence weights)
        npx = float(b * H * W)
        tgt = {}
        for key, box in (("s1", batch["mix1"]), ("s2", batch["mix2"])):
            lab = torch.empty_like(lab_w)
            wgt = torch.empty_like(conf_w)
            cnt = self._f(1)
            L.call("svl_cutmix_weights", lab_w, lab_o, conf_w, conf_o, batch["ignore_mask"], batch["ignore_mask_other"], box, lab, wgt, None,
                   cnt, lab.numel(), hp["conf_thresh"])
            mcl = None
            if lam != 0:
                mcl = torch.empty_like(lab_w)
                L.call("svl_cutmix_weights", mclip, mclip_o, None, None, None, None, box, mcl, None, None, None, mcl.numel(), 0.0)
            tgt[key] = (lab, wgt, cnt, mcl)
        w_fp = torch.empty_like(conf_w)
        cnt_fp = self._f(1)
        L.call("svl_cutmix_weights", lab_w, lab_w, conf_w, conf_w, batch["ignore_mask"], batch["ignore_mask"], None, None, w_fp, None, cnt_fp,
               lab_w.numel(), hp["conf_thresh"])
        cnt_x = self._f(1)
        L.call("svl_count_valid", batch["mask_x"], batch["mask_x"].numel(), 255, cnt_x)

        def coef(count, numer):
            out = self._f(1)
            L.call("svl_reciprocal", count, out, numer, 1.0)
            return out
        losses = self._f(7)               # x, s1, s2, fp, mc_s1, mc_s2, mc_fp  (each already multiplied by its weight in the total)
        d_low = torch.zeros_like(low)
        one = torch.ones(1, device=dev)
        rows = lambda t, i: t[i * b:(i + 1) * b]
        lx = self._f(3)
        self._ce(rows(low, 0), rows(d_low, 0), b, N, hl, wl, H, W, [(batch["mask_x"], None, coef(cnt_x, 0.5))], lx)
        lfp = self._f(3)
        t_fp = [(lab_w, w_fp, coef(cnt_fp, 0.25))]
        if lam != 0:
            t_fp.append((mclip, None, one * (lam * 0.5 / npx)))
        self._ce(rows(low, 2), rows(d_low, 2), b, N, hl, wl, H, W, t_fp, lfp)
        ls = {}
        for i, key in ((3, "s1"), (4, "s2")):
            lab, wgt, cnt, mcl = tgt[key]
            t = [(lab, wgt, coef(cnt, 0.125))]
            if lam != 0:
                t.append((mcl, None, one * (lam * 0.25 / npx)))
            ls[key] = self._f(3)
            self._ce(rows(low, i), rows(d_low, i), b, N, hl, wl, H, W, t, ls[key])
        total = lx[0] + lfp[0] + lfp[1] + ls["s1"][0] + ls["s1"][1] + ls["s2"][0] + ls["s2"][1]
        terms = dict(loss_x=lx[0] * 2, loss_fp=lfp[0] * 4, loss_s1=ls["s1"][0] * 8, loss_s2=ls["s2"][0] * 8)
        if lam != 0:
            terms.update(loss_mc_fp=lfp[1] / (lam * 0.5), loss_mc_s1=ls["s1"][1] / (lam * 0.25), loss_mc_s2=ls["s2"][1] / (lam * 0.25))
        # ---- backward
        dhf = self.head.backward(hctx, d_low, ph, self.g_hd)
        del hctx
        dfe = []
        for d, k in zip(dhf, dm):
            g = torch.empty(4 * b, *d.shape[1:], device=dev, dtype=torch.float32)
            g[:b] = d[:b]
            g[b:2 * b] = d[b:2 * b] + d[2 * b:3 * b] * k
            g[2 * b:] = d[3 * b:]
            dfe.append(g)
        del dhf
        self.vit.backward(vctx, dfe, pb, self.g_bb)
        del vctx
        if update:
            self.optimizer_step()
        return total, terms
