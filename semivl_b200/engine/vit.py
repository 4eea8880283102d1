"""CLIP ViT-B/16 image encoder with MaskCLIP v-path taps: forward and hand-scheduled backward over the C-ABI kernels.

Mirrors MaskClipVisionTransformer.forward / TransformerEncoderLayer.forward(+forward_qkv)
(third_party/maskclip/models/backbones/maskclip_vit.py:110-144,492-596) for the in-scope configuration
(pre_norm, final_norm, return_clip_embed, return_qkv, with_cls_token).

Data layout: tokens are rows of [B*L, E]; the residual stream is fp32, every GEMM operand is bf16 (fast) or a split
bf16 pair (precise).  Only `attn.*` weights/biases and `pos_embed` receive gradients (model/vlm.py:80-88): the backward
runs dgrad everywhere and wgrad only for in_proj / out_proj.
"""
import torch

from .. import lib as L
from .. import ops


class WeightCache:
    """bf16 (or split) GEMM-operand copies of fp32 parameters, refreshed when the parameter's version changes.

    Batched mode (`batched = True`, set by the Trainer for the bf16 throughput mode): the operand copies of the VOLATILE parameters (the ones
    the fused AdamW kernel rewrites every step) are persistent tensors, each described once by an index map (any operand layout -- reshape,
    transpose, tap-major conv layouts, zero padding -- is a gather of the parameter), and `refresh()` rewrites all of them in ONE kernel
    launch after the optimizer step instead of two or three ATen launches per tensor (160 of the 214 ATen launches of a step).  Weight
    gradients that the kernels produce in an operand layout are staged the same way and `scatter_grads()` adds them into the parameter-layout
    gradient buffer in one launch."""

    def __init__(self):
        self._c = {}
        self.gen = 0                 # bumped by whoever updates parameters behind torch's back (the fused AdamW kernel)
        self.volatile = None         # optional set of parameter names that change every step
        self.batched = False
        self._jobs, self._gjobs = {}, {}          # key -> (ptr, idx, tensor, flags)
        self._table, self._gtable = None, None

    def get(self, key, param, fn):
        vol = self.volatile is not None and key[0] in self.volatile
        ver = (param.data_ptr(), param._version, self.gen if vol else 0)
        hit = self._c.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                hit = (ver, fn(param.detach()))
            self._c[key] = hit
        return hit[1]

    @staticmethod
    def _index_map(t, layout_fn):
        """(int32 map operand position -> parameter element, -1 = padding; operand shape) of the gather `layout_fn`"""
        assert t.numel() < (1 << 24), "index maps are derived in fp32 (exact below 2^24 elements)"
        ar = torch.arange(1, t.numel() + 1, device=t.device, dtype=torch.float32).reshape(t.shape)
        lay = layout_fn(ar)
        return (lay.contiguous().reshape(-1).to(torch.int32) - 1).contiguous(), tuple(lay.shape)

    def get_layout(self, key, param, layout_fn, precise, raw_f32=False):
        """Operand copy of `param` in the layout `layout_fn(fp32 tensor) -> fp32 tensor` (a pure gather: reshape / permute / slice / zero
        padding): bf16, split bf16 pair (precise) or fp32 (raw_f32)."""
        if self.batched and not precise and self.volatile is not None and key[0] in self.volatile:
            job = self._jobs.get(key)
            ver = (param.data_ptr(), param._version)          # the Trainer's `.data` views stay at version 0; anyone else's in-place update shows
            if job is None or job[4] != ver:
                with torch.no_grad():
                    lay = layout_fn(param.detach().float()).contiguous()
                    dst = lay if raw_f32 else lay.to(torch.bfloat16)
                    idx = self._index_map(param, layout_fn)[0]
                    flags = 1 if raw_f32 else 0
                    if lay.dim() == 2:                       # the plain transpose of the parameter as [R, C]?  -> tiled path of the kernel
                        C_, R_ = lay.shape
                        k = torch.arange(idx.numel(), device=idx.device)
                        if R_ > 1 and C_ > 1 and bool((idx.long() == (k % R_) * C_ + k // R_).all()):
                            flags |= 2 | (R_ << 8)
                    job = (param.data_ptr(), idx, dst, flags, ver)
                self._jobs[key] = job
                self._table = None
            return job[2]

        def fn(t):
            lay = layout_fn(t.float()).contiguous()
            return lay if raw_f32 else ops.prep_weight(lay, precise)
        return self.get(key, param, fn)

    def grad_staging(self, key, grad, layout_fn):
        """Persistent zeroed fp32 buffer in the operand layout `layout_fn` of the parameter whose gradient (a view of the flat gradient
        buffer) is `grad`; the kernels accumulate into it and scatter_grads() moves it.  None when not in batched mode."""
        if not self.batched:
            return None
        job = self._gjobs.get(key)
        if job is None or job[0] != grad.data_ptr():
            with torch.no_grad():
                idx, shape = self._index_map(grad, layout_fn)
                job = (grad.data_ptr(), idx, torch.zeros(shape, device=grad.device, dtype=torch.float32), 0, None)
            self._gjobs[key] = job
            self._gtable = None
        return job[2]

    @staticmethod
    def _build_table(jobs, scatter):
        rows, starts, nb = [], [], 0
        for ptr, idx, t, flags, _ in jobs.values():
            src, dst = (t.data_ptr(), ptr) if scatter else (ptr, t.data_ptr())
            rows.append([src, dst, idx.data_ptr(), idx.numel(), flags])
            starts.append(nb)
            if not scatter and flags & 2:                   # one block per 32 x 32 tile
                R_ = flags >> 8
                nb += ((R_ + 31) // 32) * ((idx.numel() // R_ + 31) // 32)
            else:
                nb += (idx.numel() + 1023) // 1024
        dev = next(iter(jobs.values()))[1].device
        return (torch.tensor(rows, dtype=torch.int64).to(dev), torch.tensor(starts + [nb], dtype=torch.int32).to(dev), len(rows), nb)

    def refresh(self):
        """rewrite every registered operand copy from the current parameter values (one launch)"""
        if self._jobs:
            if self._table is None:
                self._table = self._build_table(self._jobs, False)
            L.call("svl_param_jobs", *self._table, 0)

    def scatter_grads(self):
        """add every staged weight gradient into the parameter-layout gradient buffer and re-zero the staging buffers (one launch)"""
        if self._gjobs:
            if self._gtable is None:
                self._gtable = self._build_table(self._gjobs, True)
            L.call("svl_param_jobs", *self._gtable, 1)

    def bump(self):
        self.gen += 1

    def clear(self):
        self._c.clear()
        self._jobs.clear()
        self._gjobs.clear()
        self._table = self._gtable = None


class VitCfg:
    def __init__(self, embed=768, heads=12, layers=12, patch=16, out_indices=(0, 4, 12), eps=1e-6, proj_dim=512, img_size=512):
        self.embed, self.heads, self.layers, self.patch = embed, heads, layers, patch
        self.out_indices, self.eps, self.proj_dim, self.img_size = tuple(out_indices), eps, proj_dim, img_size


class VitEngine:
    """Functional encoder over a parameter dict `p` (reference names without prefix, e.g. 'layers.0.attn.attn.in_proj_weight')."""

    def __init__(self, cfg, precise=False):
        self.cfg = cfg
        self.precise = precise
        self.cache = WeightCache()

    # ------------------------------------------------------------------ weights
    def _w(self, p, name, transpose=False, rows=None):
        """GEMM operand of parameter `name` viewed as 2-D [out, in] (transpose=True -> [in, out] for data gradients)."""
        prm = p[name]
        return self.cache.get_layout((name, transpose, self.precise), prm,
                                     (lambda t: t.reshape(t.shape[0], -1).t()) if transpose else (lambda t: t.reshape(t.shape[0], -1)), self.precise)

    def _tap_layers(self):
        c = self.cfg
        taps = {i for i in c.out_indices if i < c.layers}
        taps.add(c.layers - 1)
        return taps

    # ------------------------------------------------------------------ pos embed (maskclip_vit.py:431-490)
    def _pos(self, p, hp, wp, need_grad):
        pos = p["pos_embed"]                                  # [1, 1 + g*g, E]
        if pos.shape[1] == hp * wp + 1:
            return pos[0], None
        g = int(round((pos.shape[1] - 1) ** 0.5))
        assert g * g + 1 == pos.shape[1], "square position grids only (maskclip_vit.py:278-285)"
        out = torch.empty(hp * wp + 1, pos.shape[2], device=pos.device, dtype=torch.float32)
        L.call("svl_pos_resize_fwd", pos.detach().contiguous(), out, g, g, hp, wp, pos.shape[2])
        return out, ((g, hp, wp) if need_grad else None)

    # ------------------------------------------------------------------ forward
    def forward(self, img, p, need_grad=True, want_global=True):
        """img f32 [B,3,H,W] -> (feats: list of f32 [B,h,w,C] contiguous NHWC (taps..., clip embedding), global [B,512] or None, ctx)."""
        c, pr = self.cfg, self.precise
        E, H = c.embed, c.heads
        B = img.shape[0]
        img = img.contiguous().float()
        a, hp, wp = ops.patchify(img, pr, c.patch)
        hw = hp * wp
        Lq = hw + 1
        M = B * Lq
        dev = img.device
        f32 = dict(device=dev, dtype=torch.float32)
        ctx = dict(B=B, hp=hp, wp=wp, layers=[]) if need_grad else None

        patches = torch.empty(B * hw, E, **f32)
        ops.gemm(a, self._w(p, "patch_embed.projection.weight"), patches, n=E, k=3 * c.patch * c.patch, precise=pr,
                 bias=p.get("patch_embed.projection.bias"))
        pos, pos_ctx = self._pos(p, hp, wp, need_grad)
        x0 = ops.assemble_tokens(patches, p["cls_token"].reshape(-1), pos, B, hw).view(M, E)
        x, mu0, rs0 = ops.layernorm_fwd(x0, p["ln0.weight"], p["ln0.bias"], c.eps, out_dtype=L.F32, save_stats=need_grad)
        if need_grad:
            ctx.update(x0=x0, mu0=mu0, rs0=rs0, pos_ctx=pos_ctx)

        taps = self._tap_layers()
        feats_v = {}
        pre_dt = torch.float32 if pr else torch.bfloat16
        # throughput mode saves gelu'(z) instead of z: the FFN data-gradient epilogue becomes a multiply
        gelu_act = L.ACT_GELU if pr else L.ACT_GELU_DSAVE
        for i in range(c.layers):
            pre = f"layers.{i}."
            want_v = i in taps
            last = i == c.layers - 1
            S = {}
            y, mu1, rs1 = ops.layernorm_fwd(x, p[pre + "ln1.weight"], p[pre + "ln1.bias"], c.eps, precise=pr, save_stats=need_grad)
            qkv = ops.new_act(M, 3 * E, pr, dev)
            ops.gemm(y, self._w(p, pre + "attn.attn.in_proj_weight"), qkv, n=3 * E, k=E, precise=pr, bias=p[pre + "attn.attn.in_proj_bias"],
                     out_dtype=ops.act_dtype(pr))
            wout = self._w(p, pre + "attn.attn.out_proj.weight")
            bout = p[pre + "attn.attn.out_proj.bias"]
            w1, b1 = self._w(p, pre + "ffn.layers.0.0.weight"), p[pre + "ffn.layers.0.0.bias"]
            w2, b2 = self._w(p, pre + "ffn.layers.1.weight"), p[pre + "ffn.layers.1.bias"]
            g2, be2 = p[pre + "ln2.weight"], p[pre + "ln2.bias"]
            x_in = x
            need_x = (not last) or want_global           # the x path after the last layer only feeds the global embedding
            if need_x:
                att, lse = ops.attention_fwd(qkv, B, Lq, H, pr, want_lse=need_grad)
                x_mid = torch.empty(M, E, **f32)
                ops.gemm(att, wout, x_mid, n=E, k=E, precise=pr, bias=bout, residual=x_in)
                y2, mu2, rs2 = ops.layernorm_fwd(x_mid, g2, be2, c.eps, precise=pr, save_stats=need_grad)
                hpre = torch.empty(M, 4 * E, device=dev, dtype=pre_dt) if (need_grad and not last) else None
                hact = ops.new_act(M, 4 * E, pr, dev)
                ops.gemm(y2, w1, hact, n=4 * E, k=E, precise=pr, bias=b1, act=gelu_act if hpre is not None else L.ACT_GELU, preact_out=hpre, out_dtype=ops.act_dtype(pr))
                x = torch.empty(M, E, **f32)
                ops.gemm(hact, w2, x, n=E, k=4 * E, precise=pr, bias=b2, residual=x_mid)
                del hact, y2
                if need_grad and not last:
                    S.update(att=att, lse=lse, x_mid=x_mid, mu2=mu2, rs2=rs2, hpre=hpre)
            if want_v:
                # MaskCLIP v-path (forward_qkv, maskclip_vit.py:110-118,131-132): out_proj applied to V, + x, then the FFN block
                v1 = torch.empty(M, E, **f32)
                ops.gemm(qkv, wout, v1, n=E, k=E, precise=pr, a_koff=2 * E, bias=bout, residual=x_in)
                yv, muv, rsv = ops.layernorm_fwd(v1, g2, be2, c.eps, precise=pr, save_stats=need_grad)
                vpre = torch.empty(M, 4 * E, device=dev, dtype=pre_dt) if need_grad else None
                hv = ops.new_act(M, 4 * E, pr, dev)
                ops.gemm(yv, w1, hv, n=4 * E, k=E, precise=pr, bias=b1, act=gelu_act if vpre is not None else L.ACT_GELU, preact_out=vpre, out_dtype=ops.act_dtype(pr))
                v2 = torch.empty(M, E, **f32)
                ops.gemm(hv, w2, v2, n=E, k=4 * E, precise=pr, bias=b2, residual=v1)
                del hv, yv
                feats_v[i] = v2
                if need_grad:
                    S.update(v1=v1, muv=muv, rsv=rsv, vpre=vpre)
            if need_grad:
                S.update(x_in=x_in, y=y, qkv=qkv, mu1=mu1, rs1=rs1, has_x=need_x and not last, has_v=want_v)
                ctx["layers"].append(S)

        # final norm + CLIP projection (maskclip_vit.py:538-555,588-589)
        gf, bf = p["ln1.weight"], p["ln1.bias"]
        wproj = self._w(p, "proj.weight")
        v_last = feats_v[c.layers - 1]
        vn, muf, rsf = ops.layernorm_fwd(v_last, gf, bf, c.eps, precise=pr, save_stats=need_grad)
        emb_raw = torch.empty(M, c.proj_dim, **f32)
        ops.gemm(vn, wproj, emb_raw, n=c.proj_dim, k=E, precise=pr)
        emb_n, _, inv = ops.l2norm_fwd(emb_raw, eps=0.0)          # x / x.norm() (no eps clamp in the reference)
        glob = None
        if want_global:
            xn, _, _ = ops.layernorm_fwd(x.view(B, Lq * E)[:, :E], gf, bf, c.eps, precise=pr, save_stats=False, ldx=Lq * E)
            graw = torch.empty(B, c.proj_dim, **f32)
            ops.gemm(xn, wproj, graw, n=c.proj_dim, k=E, precise=pr)
            glob, _, _ = ops.l2norm_fwd(graw, eps=0.0)
        if need_grad:
            ctx.update(v_last=v_last, muf=muf, rsf=rsf, emb_n=emb_n, inv=inv)

        def drop_cls(t, C):
            out = torch.empty(B, hp, wp, C, **f32)
            ops.cast(t, L.F32, out, L.F32, hw, C, ld_src=C, ld_dst=C, batch=B, src_batch_stride=Lq * C, dst_batch_stride=hw * C, src_offset=C)
            return out
        feats = [drop_cls(feats_v[i], E) for i in c.out_indices if i < c.layers]
        if c.layers in c.out_indices:
            feats.append(drop_cls(emb_n, c.proj_dim))
        return feats, glob, ctx

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, dfeats, p, grads, on_layer_done=None):
        """dfeats: list (same order as forward's feats) of f32 [B,h,w,C] contiguous or None.
        `grads` maps parameter name -> f32 gradient tensor (same shape as the parameter), accumulated in place."""
        c, pr = self.cfg, self.precise
        E, H = c.embed, c.heads
        B, hp, wp = ctx["B"], ctx["hp"], ctx["wp"]
        hw = hp * wp
        Lq = hw + 1
        M = B * Lq
        dev = ctx["x0"].device
        f32 = dict(device=dev, dtype=torch.float32)
        gdt = L.F32 if pr else L.BF16                       # storage of intermediate (non-operand) gradients
        gtorch = torch.float32 if pr else torch.bfloat16
        gelu_dact = L.ACT_GELU if pr else L.ACT_SAVED

        def add_cls(g, C):
            out = torch.zeros(B, Lq, C, **f32)
            ops.cast(g.contiguous(), L.F32, out, L.F32, hw, C, ld_src=C, ld_dst=C, batch=B, src_batch_stride=hw * C, dst_batch_stride=Lq * C,
                     dst_offset=C)
            return out.view(M, C)

        tap_ids = [i for i in c.out_indices if i < c.layers]
        dv = {}
        for i, g in zip(tap_ids, dfeats[:len(tap_ids)]):
            if g is not None:
                dv[i] = add_cls(g, E)
        demb = dfeats[len(tap_ids)] if c.layers in c.out_indices and len(dfeats) > len(tap_ids) else None
        if demb is not None:
            d_embn = add_cls(demb, c.proj_dim)
            d_raw = torch.empty(M, c.proj_dim, **f32)
            ops.l2norm_bwd(d_embn, L.F32, ctx["emb_n"], ctx["inv"], d_raw)
            d_vn = torch.empty(M, E, device=dev, dtype=gtorch)
            ops.gemm(ops.to_act(d_raw, pr), self._w(p, "proj.weight", transpose=True), d_vn, n=E, k=c.proj_dim, precise=pr)
            last = c.layers - 1
            dlast, _ = ops.layernorm_bwd(d_vn, gdt, ctx["v_last"], p["ln1.weight"], ctx["muf"], ctx["rsf"], dres1=dv.get(last))
            dv[last] = dlast

        def wg(name):
            return grads[name]

        dx, dx_act = None, None          # gradient of the residual stream entering the layer above
        for i in reversed(range(c.layers)):
            S = ctx["layers"][i]
            pre = f"layers.{i}."
            wout_t = self._w(p, pre + "attn.attn.out_proj.weight", transpose=True)
            win_t = self._w(p, pre + "attn.attn.in_proj_weight", transpose=True)
            w1_t = self._w(p, pre + "ffn.layers.0.0.weight", transpose=True)
            w2_t = self._w(p, pre + "ffn.layers.1.weight", transpose=True)
            g_wout, g_bout = wg(pre + "attn.attn.out_proj.weight"), wg(pre + "attn.attn.out_proj.bias")
            g_win, g_bin = wg(pre + "attn.attn.in_proj_weight"), wg(pre + "attn.attn.in_proj_bias")
            g2 = p[pre + "ln2.weight"]
            dv2 = dv.get(i) if S["has_v"] else None
            dv1 = dvv = None
            have_x = dx is not None and S["has_x"]
            if dv2 is not None:
                dhv = ops.new_act(M, 4 * E, pr, dev)
                ops.gemm(ops.to_act(dv2, pr), w2_t, dhv, n=4 * E, k=E, precise=pr, dact_src=S["vpre"], dact_kind=gelu_dact,
                         out_dtype=ops.act_dtype(pr))
                dyv = torch.empty(M, E, device=dev, dtype=gtorch)
                ops.gemm(dhv, w1_t, dyv, n=E, k=4 * E, precise=pr)
                del dhv
                dv1, dv1_act = ops.layernorm_bwd(dyv, gdt, S["v1"], g2, S["muv"], S["rsv"], dres1=dv2, act_precise=pr)
                ops.wgrad(dv1_act, S["qkv"], g_wout, m=E, n=E, precise=pr, x_koff=2 * E)
                ops.colsum(dv1, L.F32, M, E, g_bout)
                if have_x:
                    dvv = torch.empty(M, E, device=dev, dtype=gtorch)
                    ops.gemm(dv1_act, wout_t, dvv, n=E, k=E, precise=pr)
                else:                                        # consumed as a GEMM operand below
                    dvv = ops.new_act(M, E, pr, dev)
                    ops.gemm(dv1_act, wout_t, dvv, n=E, k=E, precise=pr, out_dtype=ops.act_dtype(pr))
            if have_x:
                if dx_act is None:
                    dx_act = ops.to_act(dx, pr)
                dh = ops.new_act(M, 4 * E, pr, dev)
                ops.gemm(dx_act, w2_t, dh, n=4 * E, k=E, precise=pr, dact_src=S["hpre"], dact_kind=gelu_dact, out_dtype=ops.act_dtype(pr))
                dy2 = torch.empty(M, E, device=dev, dtype=gtorch)
                ops.gemm(dh, w1_t, dy2, n=E, k=4 * E, precise=pr)
                del dh
                dx_mid, dx_mid_act = ops.layernorm_bwd(dy2, gdt, S["x_mid"], g2, S["mu2"], S["rs2"], dres1=dx, act_precise=pr)
                ops.wgrad(dx_mid_act, S["att"], g_wout, m=E, n=E, precise=pr)
                ops.colsum(dx_mid, L.F32, M, E, g_bout)
                datt = ops.new_act(M, E, pr, dev)
                ops.gemm(dx_mid_act, wout_t, datt, n=E, k=E, precise=pr, out_dtype=ops.act_dtype(pr))
                dqkv = ops.attention_bwd(S["qkv"], S["att"], datt, S["lse"], B, Lq, H, pr, dv_add=dvv, dv_add_dtype=gdt)
                ops.wgrad(dqkv, S["y"], g_win, m=3 * E, n=E, precise=pr)
                ops.colsum(dqkv, ops.act_dtype(pr), M, 3 * E, g_bin)
                dy1 = torch.empty(M, E, device=dev, dtype=gtorch)
                ops.gemm(dqkv, win_t, dy1, n=E, k=3 * E, precise=pr)
                dx, dx_act = ops.layernorm_bwd(dy1, gdt, S["x_in"], p[pre + "ln1.weight"], S["mu1"], S["rs1"], dres1=dx_mid, dres2=dv1,
                                               act_precise=pr)
            elif dvv is not None:
                # only the V third of in_proj sees a gradient (no gradient reaches this layer's attention output)
                dvv_act = dvv
                ops.wgrad(dvv_act, S["y"], g_win[2 * E:], m=E, n=E, precise=pr)
                ops.colsum(dvv, ops.act_dtype(pr), M, E, g_bin[2 * E:])
                dy1 = torch.empty(M, E, device=dev, dtype=gtorch)
                wv_t = self._w(p, pre + "attn.attn.in_proj_weight", transpose=True)            # [E, 3E]: columns 2E.. are the V rows
                ops.gemm(dvv_act, wv_t, dy1, n=E, k=E, precise=pr, b_col0=2 * E)
                dx, dx_act = ops.layernorm_bwd(dy1, gdt, S["x_in"], p[pre + "ln1.weight"], S["mu1"], S["rs1"], dres1=dx, dres2=dv1,
                                               act_precise=pr)
            # else: nothing reaches this layer (cannot happen with the reference's tap configuration)
            if on_layer_done is not None:
                on_layer_done(i)                         # layer i's weight gradients are final: the trainer may start exchanging them

        if dx is not None:
            dx0, _ = ops.layernorm_bwd(dx, L.F32, ctx["x0"], p["ln0.weight"], ctx["mu0"], ctx["rs0"])
            dpos = torch.empty(Lq, E, **f32)
            ops.batch_sum(dx0.view(B, Lq * E), dpos.view(-1))
            if ctx["pos_ctx"] is None:
                ops.axpy(grads["pos_embed"].view(-1), dpos.view(-1))
            else:
                g, oh, ow = ctx["pos_ctx"]                        # transpose of the bicubic resize, straight into the parameter gradient
                L.call("svl_pos_resize_bwd", dpos, grads["pos_embed"], g, g, oh, ow, E)
        return grads
