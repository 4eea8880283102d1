"""Vision-language-guided decode head: forward and hand-scheduled backward over the C-ABI kernels.

Mirrors VLGHead.forward, SemanticTransformer, ASPPModule / ASPPPooling and Up
(model/decode_heads/vlg_head.py:27-67,70-113,116-137,192-251).

Layout: every map tensor is NHWC with one "map" per (image, class) pair, maps ordered (image, class): [B*N, h, w, C].
Convolutions are implicit GEMMs on the tcgen05 engine (one tap per filter position, TMA zero-fill padding);
GEMM operands are bf16 (fast) or split bf16 pairs (precise); conv outputs that feed a GroupNorm are kept raw
(bf16 / fp32 in precise mode) next to the normalised operand because both are needed by the backward.
"""
import torch

from .. import lib as L
from .. import ops
from .vit import WeightCache


def _f3(d):
    return [((i - 1) * d, (j - 1) * d) for i in range(3) for j in range(3)]


class HeadCfg:
    def __init__(self, channels=128, text_channels=128, up_channels=(64, 32), skip_channels=(32, 16), num_layers=2, num_heads=4,
                 pool=4, conv1_ksize=7, in_dim=512, skip_dim=768, dilations=(6, 12, 18), align_corners=False):
        self.C, self.Ct, self.up, self.skip = channels, text_channels, tuple(up_channels), tuple(skip_channels)
        self.layers, self.heads, self.pool, self.ks = num_layers, num_heads, pool, conv1_ksize
        self.in_dim, self.skip_dim, self.dil, self.align_corners = in_dim, skip_dim, tuple(dilations), align_corners


class HeadEngine:
    def __init__(self, cfg, precise=False):
        self.cfg, self.precise = cfg, precise
        self.cache = WeightCache()

    # ------------------------------------------------------------------ weight operands
    @staticmethod
    def _layout(kind):
        """operand layout of a parameter as a pure gather of its elements (reshape / permute / slice / zero padding)"""
        def pad_cols(t):                   # conv1 [C,1,ks,ks] -> [C, 64]
            w = t.new_zeros(t.shape[0], 64)
            w[:, : t.shape[2] * t.shape[3]] = t.reshape(t.shape[0], -1)
            return w

        def pad_rows(t):                   # -> [64, C]
            w = t.new_zeros(64, t.shape[0])
            w[: t.shape[2] * t.shape[3]] = t.reshape(t.shape[0], -1).t()
            return w
        return {
            "lin": lambda t: t.reshape(t.shape[0], -1),
            "lin_t": lambda t: t.reshape(t.shape[0], -1).t(),
            "conv": lambda t: t.permute(2, 3, 0, 1).reshape(-1, t.shape[1]),            # [Co,Ci,kh,kw] -> [(tap, co), ci]
            "conv_t": lambda t: t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]),          # -> [(tap, ci), co]
            "convT": lambda t: t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]),           # ConvTranspose [Ci,Co,2,2] -> [(q, co), ci]
            "convT_t": lambda t: t.permute(2, 3, 0, 1).reshape(-1, t.shape[1]),         # -> [(q, ci), co]
            "conv1": pad_cols,
            "conv1_t": pad_rows,
            "projA": lambda t: t.reshape(t.shape[0], -1)[:, : 4 * t.shape[0]],          # aspp.project [C, 5C,1,1] -> first 4C input channels
            "projA_t": lambda t: t.reshape(t.shape[0], -1)[:, : 4 * t.shape[0]].t(),
            "projB": lambda t: t.reshape(t.shape[0], -1)[:, 4 * t.shape[0]:],
            "projB_t": lambda t: t.reshape(t.shape[0], -1)[:, 4 * t.shape[0]:].t(),
            "out1": lambda t: t[0].permute(1, 2, 0).reshape(-1),                        # head [1,C,3,3] -> f32 [9*C]
            "rep4": lambda t: t.repeat(4),                                              # ConvTranspose bias, once per 2 x 2 output quadrant
        }[kind]

    def _prep(self, p, name, kind):
        return self.cache.get_layout((name, kind, self.precise), p[name], self._layout(kind), self.precise, raw_f32=kind in ("out1", "rep4"))

    def _text_f32(self, text):
        """(fp32 contiguous copy of the frozen text-embedding table, memo dict for operands derived from it): made once per table instead
        of once per pass; a handful of tables can be live (pl_text / mcc_text variants of the SemiVL step)"""
        key = (text.data_ptr(), text._version, tuple(text.shape), text.dtype)
        cache = self.__dict__.setdefault("_text_cache", {})
        hit = cache.get(key)
        if hit is None:
            if len(cache) >= 4:
                cache.clear()
            hit = (text.float().contiguous(), {})
            cache[key] = hit
        return hit

    def _text_t_operand(self, memo, txt_n, Kp):
        """B operand [in_dim, Kp] of d(image embedding) = d_sim @ text_n (classes zero-padded to the MMA K step): a function of the frozen
        table only, kept in the table's memo"""
        k = ("txt_t", Kp, self.precise)
        if k not in memo:
            txt_t = torch.zeros(txt_n.shape[1], Kp, device=txt_n.device, dtype=torch.float32)
            txt_t[:, : txt_n.shape[0]] = txt_n.t()
            memo[k] = ops.prep_weight(txt_t, self.precise)
        return memo[k]

    def _wgrad_buf(self, grads, name, kind, shape):
        """(buffer the weight-gradient kernel accumulates into in operand layout `kind`, staged?): the persistent staging buffer of the
        batched mode (moved by cache.scatter_grads() at the end of backward) or a fresh zero tensor the caller adds into grads[name]"""
        st = self.cache.grad_staging((name, kind), grads[name], self._layout(kind))
        if st is not None:
            return st.view(shape), True
        return torch.zeros(shape, device=grads[name].device, dtype=torch.float32), False

    # ------------------------------------------------------------------ helpers
    def _conv_gn(self, x_act, nb, h, w, cin, wname, gname, G, p, dil, out_act, out_col0, need_grad, res=None):
        """raw = conv3x3(x) (no bias) -> relu(GN(raw)) (+res) written into out_act[:, out_col0:]. Returns saved dict."""
        pr = self.precise
        wt = p[wname]
        cout, ks = wt.shape[0], wt.shape[2]
        raw = torch.empty(nb * h * w, cout, device=x_act.device, dtype=torch.float32 if pr else torch.bfloat16)
        filt = _f3(dil) if ks == 3 else [(0, 0)]
        gs = dict(maps=nb, G=G)                 # the GroupNorm statistics come out of the convolution's epilogue where the library can (conv_roll)
        ops.gemm(x_act, self._prep(p, wname, "conv"), raw, n=cout, k=cin, precise=pr, conv=(nb, h, w), filt=filt, b_row_stride=cout, gn_stats=gs)
        mean, rstd = ops.gn_relu_fwd(raw, L.dtype_of(raw), p[gname + ".weight"], p[gname + ".bias"], out_act, ops.act_dtype(pr), nb, h * w, cout, G,
                                     out_col0=out_col0, res=res, res_dtype=ops.act_dtype(pr), save_stats=need_grad, stats=gs)
        return dict(raw=raw, mean=mean, rstd=rstd) if need_grad else None

    def _conv_gn_bwd(self, S, dy, dy_dtype, dy_col0, x_act, nb, h, w, cin, wname, gname, G, p, grads, dil, dx_out, dx_dtype, accumulate=False):
        """Backward of _conv_gn: GN/ReLU backward, conv wgrad and dgrad (dx_out written or accumulated)."""
        pr = self.precise
        wt = p[wname]
        cout, ks = wt.shape[0], wt.shape[2]
        filt = _f3(dil) if ks == 3 else [(0, 0)]
        d_raw = ops.new_act(nb * h * w, cout, pr, x_act.device)
        ops.gn_relu_bwd(dy, dy_dtype, S["raw"], L.dtype_of(S["raw"]), p[gname + ".weight"], p[gname + ".bias"], S["mean"], S["rstd"], d_raw,
                        ops.act_dtype(pr), grads[gname + ".weight"], grads[gname + ".bias"], nb, h * w, cout, G, dy_col0=dy_col0)
        dw, staged = self._wgrad_buf(grads, wname, "conv", (len(filt), cout, cin))
        ops.wgrad(d_raw, x_act, dw, m=cout, n=cin, precise=pr, conv=(nb, h, w), filt=filt)
        if not staged:
            grads[wname].add_(dw.view(ks, ks, cout, cin).permute(2, 3, 0, 1))
        if dx_out is not None:
            mirrored = [(-a, -b) for a, b in filt]
            if (not pr and cin == 128 and cout in (32, 64) and ks == 3 and dil == 1 and not accumulate and dx_dtype == L.BF16 and
                    dx_out.shape[-1] == cin and (33 <= w <= 64 or w >= 96)):
                # 128 data-gradient channels = 3 x 128 accumulator columns per strip, more than one MMA holds: two 64-channel halves on the rolling
                # convolution kernel (conv_roll.cu) instead of one launch of the generic engine (281 -> 2 x 97 us at config-2 size)
                wt_t = self._prep(p, wname, "conv_t")                  # [(tap, ci), co]
                for half in range(2):
                    taps = [(fy, fx, 0, t * cin + half * 64, 0) for t, (fy, fx) in enumerate(mirrored)]
                    ops.gemm(d_raw, wt_t, dx_out[:, half * 64:], n=64, k=cout, conv=(nb, h, w), taps=taps, out_dtype=dx_dtype, ldc=cin)
            else:
                ops.gemm(d_raw, self._prep(p, wname, "conv_t"), dx_out, n=cin, k=cout, precise=pr, conv=(nb, h, w), filt=mirrored,
                         b_row_stride=cin, out_dtype=dx_dtype, accumulate=accumulate)
        return d_raw

    # ------------------------------------------------------------------ class-attention encoder layer (SemanticTransformer.transformer)
    def _tlayer_fwd(self, x, p, pre, nseq, seq, need_grad):
        pr, c = self.precise, self.cfg
        E = c.C + c.Ct
        F_ = p[pre + "ffn.layers.0.0.weight"].shape[0]
        M = x.shape[0]
        dev = x.device
        y, mu1, rs1 = ops.layernorm_fwd(x, p[pre + "ln1.weight"], p[pre + "ln1.bias"], 1e-5, precise=pr, save_stats=need_grad)
        qkv = ops.new_act(M, 3 * E, pr, dev)
        ops.gemm(y, self._prep(p, pre + "attn.attn.in_proj_weight", "lin"), qkv, n=3 * E, k=E, precise=pr, bias=p[pre + "attn.attn.in_proj_bias"],
                 out_dtype=ops.act_dtype(pr))
        att, lse = ops.attention_fwd(qkv, nseq, seq, c.heads, pr, want_lse=need_grad)
        x_mid = torch.empty(M, E, device=dev, dtype=torch.float32)
        ops.gemm(att, self._prep(p, pre + "attn.attn.out_proj.weight", "lin"), x_mid, n=E, k=E, precise=pr, bias=p[pre + "attn.attn.out_proj.bias"],
                 residual=x)
        y2, mu2, rs2 = ops.layernorm_fwd(x_mid, p[pre + "ln2.weight"], p[pre + "ln2.bias"], 1e-5, precise=pr, save_stats=need_grad)
        hpre = torch.empty(M, F_, device=dev, dtype=torch.float32 if pr else torch.bfloat16) if need_grad else None
        hact = ops.new_act(M, F_, pr, dev)
        ops.gemm(y2, self._prep(p, pre + "ffn.layers.0.0.weight", "lin"), hact, n=F_, k=E, precise=pr, bias=p[pre + "ffn.layers.0.0.bias"],
                 act=L.ACT_GELU, preact_out=hpre, out_dtype=ops.act_dtype(pr))
        out = torch.empty(M, E, device=dev, dtype=torch.float32)
        ops.gemm(hact, self._prep(p, pre + "ffn.layers.1.weight", "lin"), out, n=E, k=F_, precise=pr, bias=p[pre + "ffn.layers.1.bias"], residual=x_mid)
        S = dict(x=x, y=y, mu1=mu1, rs1=rs1, qkv=qkv, att=att, lse=lse, x_mid=x_mid, y2=y2, mu2=mu2, rs2=rs2, hpre=hpre, hact=hact) if need_grad else None
        return out, S

    def _tlayer_bwd(self, S, dout, p, pre, nseq, seq, grads):
        pr, c = self.precise, self.cfg
        E = c.C + c.Ct
        F_ = p[pre + "ffn.layers.0.0.weight"].shape[0]
        M = dout.shape[0]
        dev = dout.device
        gdt = L.F32 if pr else L.BF16
        gtorch = torch.float32 if pr else torch.bfloat16
        adt = ops.act_dtype(pr)
        dout_act = ops.to_act(dout, pr)
        # FFN
        ops.wgrad(dout_act, S["hact"], grads[pre + "ffn.layers.1.weight"], m=E, n=F_, precise=pr)
        ops.colsum(dout, L.F32, M, E, grads[pre + "ffn.layers.1.bias"])
        dh = ops.new_act(M, F_, pr, dev)
        ops.gemm(dout_act, self._prep(p, pre + "ffn.layers.1.weight", "lin_t"), dh, n=F_, k=E, precise=pr, dact_src=S["hpre"], dact_kind=L.ACT_GELU,
                 out_dtype=adt)
        ops.wgrad(dh, S["y2"], grads[pre + "ffn.layers.0.0.weight"], m=F_, n=E, precise=pr)
        ops.colsum(dh, adt, M, F_, grads[pre + "ffn.layers.0.0.bias"])
        dy2 = torch.empty(M, E, device=dev, dtype=gtorch)
        ops.gemm(dh, self._prep(p, pre + "ffn.layers.0.0.weight", "lin_t"), dy2, n=E, k=F_, precise=pr)
        dx_mid, dx_mid_act = ops.layernorm_bwd(dy2, gdt, S["x_mid"], p[pre + "ln2.weight"], S["mu2"], S["rs2"], dres1=dout, act_precise=pr,
                                               dgamma=grads[pre + "ln2.weight"], dbeta=grads[pre + "ln2.bias"])
        # attention
        ops.wgrad(dx_mid_act, S["att"], grads[pre + "attn.attn.out_proj.weight"], m=E, n=E, precise=pr)
        ops.colsum(dx_mid, L.F32, M, E, grads[pre + "attn.attn.out_proj.bias"])
        datt = ops.new_act(M, E, pr, dev)
        ops.gemm(dx_mid_act, self._prep(p, pre + "attn.attn.out_proj.weight", "lin_t"), datt, n=E, k=E, precise=pr, out_dtype=adt)
        dqkv = ops.attention_bwd(S["qkv"], S["att"], datt, S["lse"], nseq, seq, c.heads, pr)
        ops.wgrad(dqkv, S["y"], grads[pre + "attn.attn.in_proj_weight"], m=3 * E, n=E, precise=pr)
        ops.colsum(dqkv, adt, M, 3 * E, grads[pre + "attn.attn.in_proj_bias"])
        dy1 = torch.empty(M, E, device=dev, dtype=gtorch)
        ops.gemm(dqkv, self._prep(p, pre + "attn.attn.in_proj_weight", "lin_t"), dy1, n=E, k=3 * E, precise=pr)
        dx, _ = ops.layernorm_bwd(dy1, gdt, S["x"], p[pre + "ln1.weight"], S["mu1"], S["rs1"], dres1=dx_mid,
                                  dgamma=grads[pre + "ln1.weight"], dbeta=grads[pre + "ln1.bias"])
        return dx

    # ------------------------------------------------------------------ Up block (vlg_head.py:116-137)
    def _up_fwd(self, x_act, skip, skip_hw, nb, B, N, h, w, cin, name, G, p, need_grad):
        pr, dev = self.precise, x_act.device
        adt = ops.act_dtype(pr)
        wt = p[name + "up.weight"]                       # [cin, cup, 2, 2]
        cup, cs = wt.shape[1], skip.shape[-1]
        ccat = cup + cs
        cout = p[name + "conv.0.weight"].shape[0]
        H2, W2 = 2 * h, 2 * w
        cat = ops.new_act(nb * H2 * W2, ccat, pr, dev)
        ops.gemm(x_act, self._prep(p, name + "up.weight", "convT"), cat, n=4 * cup, k=cin, precise=pr, bias=self._prep(p, name + "up.bias", "rep4"),
                 out_mode=L.OUT_CONVT2X2, out_hw=(h, w), out_dtype=adt, m=nb * h * w)
        L.call("svl_skip_fill", skip, L.F32, cs, cat, adt, cat.shape[-1], cup, B, N, skip_hw[0], skip_hw[1], cs, H2, W2)
        a0 = ops.new_act(nb * H2 * W2, cout, pr, dev)
        S0 = self._conv_gn(cat, nb, H2, W2, ccat, name + "conv.0.weight", name + "conv.1", G, p, 1, a0, 0, need_grad)
        a1 = ops.new_act(nb * H2 * W2, cout, pr, dev)
        S1 = self._conv_gn(a0, nb, H2, W2, cout, name + "conv.3.weight", name + "conv.4", G, p, 1, a1, 0, need_grad)
        S = dict(x=x_act, cat=cat, a0=a0, S0=S0, S1=S1, skip=skip, skip_hw=skip_hw) if need_grad else None
        return a1, S

    def _up_bwd(self, S, dy, dy_dtype, nb, B, N, h, w, cin, name, G, p, grads):
        """dy: gradient of the block output [nb*4hw, cout].  Returns (dx [nb*hw, cin] (intermediate-gradient dtype), d_skip_pre operand)."""
        pr, dev = self.precise, dy.device
        adt = ops.act_dtype(pr)
        gtorch = torch.float32 if pr else torch.bfloat16
        gdt = L.F32 if pr else L.BF16
        wt = p[name + "up.weight"]
        cup, cs = wt.shape[1], S["skip"].shape[-1]
        ccat = cup + cs
        cout = p[name + "conv.0.weight"].shape[0]
        H2, W2 = 2 * h, 2 * w
        d_a0 = torch.empty(nb * H2 * W2, cout, device=dev, dtype=gtorch)
        self._conv_gn_bwd(S["S1"], dy, dy_dtype, 0, S["a0"], nb, H2, W2, cout, name + "conv.3.weight", name + "conv.4", G, p, grads, 1, d_a0, gdt)
        d_cat = ops.new_act(nb * H2 * W2, ccat, pr, dev)
        self._conv_gn_bwd(S["S0"], d_a0, gdt, 0, S["cat"], nb, H2, W2, ccat, name + "conv.0.weight", name + "conv.1", G, p, grads, 1, d_cat, adt)
        del d_a0
        ldp = d_cat.shape[-1]                               # physical row length of d_cat
        # skip branch: gradient w.r.t. the pre-ReLU skip projection
        sh, sw = S["skip_hw"]
        d_skip = ops.new_act(B * sh * sw, cs, pr, dev)
        d_sum = torch.empty(B, H2 * W2, cs, device=dev, dtype=torch.float32)            # sum over the N per-class copies first ...
        L.call("svl_class_sum", d_cat, adt, ldp, cup, d_sum, B, N, H2 * W2, cs)
        L.call("svl_skip_grad", d_sum, L.F32, cs, 0, S["skip"], L.F32, cs, d_skip, adt, d_skip.shape[-1], B, 1, sh, sw, cs, H2, W2)   # ... then resize^T
        # transposed conv: bias, weight and data gradients; d_cat is read as [nb, h, 2w', 2*ldp] (pixel (2y+qy, 2x+qx) -> x' = qy*w + x, column block qx)
        ops.colsum(d_cat, adt, nb * H2 * W2, cup, grads[name + "up.bias"], ld=ldp)
        taps_w = [(0, qy * w, 0, qx * ldp, qy * 2 + qx) for qy in range(2) for qx in range(2)]
        dwq, staged = self._wgrad_buf(grads, name + "up.weight", "convT_t", (4, cin, cup))
        ops.wgrad(S["x"], d_cat, dwq, m=cin, n=cup, precise=pr, conv=(nb, h, w), taps=taps_w, ld_x=2 * ldp, x_map_w=2 * w, x_lo=ccat,
                  slot_stride=cin * cup, ld_dw=cup)
        if not staged:
            grads[name + "up.weight"].add_(dwq.view(2, 2, cin, cup).permute(2, 3, 0, 1))
        dx = torch.empty(nb * h * w, cin, device=dev, dtype=gtorch)
        taps_d = [(0, qy * w, qx * ldp, (qy * 2 + qx) * cin, 0) for qy in range(2) for qx in range(2)]
        ops.gemm(d_cat, self._prep(p, name + "up.weight", "convT_t"), dx, n=cin, k=cup, precise=pr, conv=(nb, h, w), taps=taps_d, lda=2 * ldp,
                 a_map_w=2 * w, a_lo=ccat)
        return dx, d_skip

    # ------------------------------------------------------------------ forward
    def forward(self, feats, text, p, need_grad=True, conv_feats=None):
        """feats: [skip taps (shallow..deep)..., clip embedding], each f32 NHWC [B,h,w,C]; text f32/f16 [N, in_dim];
        conv_feats: optional list of f32 NHWC [B,sh,sw,Cc] features of a conv encoder (`skip_from_conv_feat`, vlg_head.py:196-205): they
        follow the ViT taps in the skip list, deepest first, at their own resolution.
        Returns (low-resolution logits f32 [B, N, 4h, 4w], ctx)."""
        c, pr = self.cfg, self.precise
        adt = ops.act_dtype(pr)
        gtorch = torch.float32 if pr else torch.bfloat16
        emb = feats[-1]
        skips_in = list(feats[:-1])[::-1] + list(conv_feats or [])[::-1]      # deepest first (vlg_head.py:196-208)
        n_taps = len(feats) - 1
        B, h, w, D = emb.shape
        dev = emb.device
        hw = h * w
        text, txt_memo = self._text_f32(text)
        N = text.shape[0]
        nb = B * N
        C = c.C
        f32 = dict(device=dev, dtype=torch.float32)
        ctx = dict(B=B, N=N, h=h, w=w) if need_grad else None

        # similarity (vlg_head.py:215-217)
        img_n, img_act, inv_img = ops.l2norm_fwd(emb.reshape(B * hw, D), want_f32=need_grad, act_precise=pr, eps=1e-12)
        txt_n, txt_act, _ = ops.l2norm_fwd(text, want_f32=True, act_precise=pr, eps=1e-12)
        Np = (N + 7) // 8 * 8
        sim = torch.empty(B * hw, Np, **f32)
        ops.gemm(img_act, txt_act, sim, n=N, k=D, precise=pr)
        # conv1 as im2col GEMM (vlg_head.py:220-221)
        col = ops.new_act(nb * hw, 64, pr, dev)
        L.call("svl_sim_im2col", sim, Np, col, adt, col.shape[-1], B, N, h, w, c.ks, 64)
        x1 = ops.new_act(nb * hw, C, pr, dev)
        ops.gemm(col, self._prep(p, "conv1.weight", "conv1"), x1, n=C, k=64, precise=pr, bias=p["conv1.bias"], out_dtype=adt)
        # ASPP (vlg_head.py:84-113)
        G8 = C // 16
        cat = ops.new_act(nb * hw, 4 * C, pr, dev)
        Sb = []
        for j, d in enumerate((1,) + c.dil):
            Sb.append(self._conv_gn(x1, nb, h, w, C, f"aspp.aspp_convs.{j}.0.weight", f"aspp.aspp_convs.{j}.1", G8, p, d, cat, j * C, need_grad))
        gap = torch.empty(nb, C, **f32)
        L.call("svl_map_sum", x1, adt, x1.shape[-1], gap, nb, hw, C, 1.0 / hw)
        gap_act = ops.to_act(gap, pr)
        graw = torch.empty(nb, C, **f32)
        ops.gemm(gap_act, self._prep(p, "aspp.aspp_convs.4.gap.1.weight", "lin"), graw, n=C, k=C, precise=pr)
        pool_act = ops.new_act(nb, C, pr, dev)
        gmean, grstd = ops.gn_relu_fwd(graw, L.F32, p["aspp.aspp_convs.4.gap.2.weight"], p["aspp.aspp_convs.4.gap.2.bias"], pool_act, adt, nb, 1, C, G8,
                                       save_stats=need_grad)
        rowb = torch.empty(nb, C, **f32)
        ops.gemm(pool_act, self._prep(p, "aspp.project.0.weight", "projB"), rowb, n=C, k=C, precise=pr)
        praw = torch.empty(nb * hw, C, device=dev, dtype=gtorch)
        ops.gemm(cat, self._prep(p, "aspp.project.0.weight", "projA"), praw, n=C, k=4 * C, precise=pr, row_bias=rowb, row_bias_div=hw)
        x2 = ops.new_act(nb * hw, C, pr, dev)
        pmean, prstd = ops.gn_relu_fwd(praw, L.dtype_of(praw), p["aspp.project.1.weight"], p["aspp.project.1.bias"], x2, adt, nb, hw, C, G8, res=x1,
                                       res_dtype=adt, save_stats=need_grad)
        # text projection (vlg_head.py:226-227)
        t = torch.empty(N, c.Ct, **f32)
        ops.gemm(txt_act, self._prep(p, "text_proj.0.weight", "lin"), t, n=c.Ct, k=D, precise=pr, bias=p["text_proj.0.bias"], act=L.ACT_RELU)
        # SemanticTransformer layers (vlg_head.py:39-67,229-231)
        hp, wp = h // c.pool, w // c.pool
        xcur = x2
        Sl = []
        for l in range(c.layers):
            tok0 = torch.empty(B * hp * wp * N, C + c.Ct, **f32)
            L.call("svl_pool_tokens", xcur, adt, xcur.shape[-1], t, tok0, B, N, h, w, C, c.Ct, c.pool)
            tok1, St = self._tlayer_fwd(tok0, p, f"layers.{l}.transformer.", B * hp * wp, N, need_grad)
            xnext = ops.new_act(nb * hw, C, pr, dev)
            L.call("svl_unpool_add", xcur, adt, xcur.shape[-1], tok1, C + c.Ct, xnext, adt, xnext.shape[-1], B, N, h, w, C, hp, wp)
            Sl.append(St)
            xcur = xnext
        # skip projections (vlg_head.py:233-234)
        sk, fa = [], []
        for j, f in enumerate(skips_in):
            cs = c.skip[j]
            sh, sw, cin = f.shape[1], f.shape[2], f.shape[3]
            a = ops.to_act(f.reshape(B * sh * sw, cin), pr)
            s = torch.empty(B * sh * sw, cs, **f32)
            ops.gemm(a, self._prep(p, f"skip_proj.{j}.0.weight", "conv"), s, n=cs, k=cin, precise=pr, conv=(B, sh, sw), filt=_f3(1),
                     b_row_stride=cs, bias=p[f"skip_proj.{j}.0.bias"], act=L.ACT_RELU)
            sk.append(s)
            fa.append(a)
        geo = [(f.shape[1], f.shape[2], f.shape[3]) for f in skips_in]
        # decoder (vlg_head.py:236-240)
        u1, Su1 = self._up_fwd(xcur, sk[0], geo[0][:2], nb, B, N, h, w, C, "up1.", c.up[0] // 16, p, need_grad)
        u2, Su2 = self._up_fwd(u1, sk[1], geo[1][:2], nb, B, N, 2 * h, 2 * w, c.up[0], "up2.", c.up[1] // 16, p, need_grad)
        low = torch.empty(B, N, 4 * h, 4 * w, **f32)
        L.call("svl_conv_out1_fwd", u2, adt, u2.shape[-1], self._prep(p, "head.weight", "out1"), p["head.bias"], low, nb, 4 * h, 4 * w, c.up[1])
        if need_grad:
            ctx.update(img_n=img_n, inv_img=inv_img, txt_n=txt_n, txt_memo=txt_memo, txt_act=txt_act, col=col, x1=x1, cat=cat, Sb=Sb, gap_act=gap_act, graw=graw,
                       gmean=gmean, grstd=grstd, pool_act=pool_act, praw=praw, pmean=pmean, prstd=prstd, t=t, Sl=Sl, x_final=xcur, sk=sk, fa=fa,
                       Su1=Su1, Su2=Su2, u2=u2, hp=hp, wp=wp, geo=geo, n_taps=n_taps)
        return low, ctx

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, d_low, p, grads, need_feat_grads=True):
        """d_low: f32 [B,N,4h,4w].  Accumulates parameter gradients into `grads` (name -> param-shaped f32 tensor) and returns
        the gradients of the input features [d_skip_shallow..., d_emb] as f32 NHWC tensors."""
        c, pr = self.cfg, self.precise
        adt = ops.act_dtype(pr)
        gtorch = torch.float32 if pr else torch.bfloat16
        gdt = L.F32 if pr else L.BF16
        B, N, h, w = ctx["B"], ctx["N"], ctx["h"], ctx["w"]
        nb, hw, C = B * N, h * w, c.C
        dev = d_low.device
        f32 = dict(device=dev, dtype=torch.float32)
        G8 = C // 16
        d_low = d_low.contiguous()

        # output conv
        u2 = ctx["u2"]
        d_u2 = torch.empty(nb * 16 * hw, c.up[1], device=dev, dtype=gtorch)
        dw9, staged = self._wgrad_buf(grads, "head.weight", "out1", (9 * c.up[1],))
        L.call("svl_conv_out1_bwd", d_low, u2, adt, u2.shape[-1], self._prep(p, "head.weight", "out1"), d_u2, gdt, d_u2.shape[-1], dw9,
               grads["head.bias"], nb, 4 * h, 4 * w, c.up[1], n_launch=2)
        if not staged:
            grads["head.weight"].add_(dw9.view(3, 3, c.up[1]).permute(2, 0, 1)[None])
        # decoder
        d_u1, d_sk1 = self._up_bwd(ctx["Su2"], d_u2, gdt, nb, B, N, 2 * h, 2 * w, c.up[0], "up2.", c.up[1] // 16, p, grads)
        del d_u2
        d_x, d_sk0 = self._up_bwd(ctx["Su1"], d_u1, gdt, nb, B, N, h, w, C, "up1.", c.up[0] // 16, p, grads)
        del d_u1
        # skip projections
        d_skips = []
        for j, d_sk in ((0, d_sk0), (1, d_sk1)):
            cs = c.skip[j]
            sh, sw, cin = ctx["geo"][j]
            name = f"skip_proj.{j}.0."
            dw, staged = self._wgrad_buf(grads, name + "weight", "conv", (9, cs, cin))
            ops.wgrad(d_sk, ctx["fa"][j], dw, m=cs, n=cin, precise=pr, conv=(B, sh, sw), filt=_f3(1))
            if not staged:
                grads[name + "weight"].add_(dw.view(3, 3, cs, cin).permute(2, 3, 0, 1))
            ops.colsum(d_sk, adt, B * sh * sw, cs, grads[name + "bias"])
            if need_feat_grads:
                d_f = torch.empty(B, sh, sw, cin, **f32)
                ops.gemm(d_sk, self._prep(p, name + "weight", "conv_t"), d_f, n=cin, k=cs, precise=pr, conv=(B, sh, sw),
                         filt=[(-a, -b) for a, b in _f3(1)], b_row_stride=cin)
                d_skips.append(d_f)
            else:
                d_skips.append(None)
        # back to the caller's order: ViT taps shallow..deep, [embedding], conv features shallow..deep
        nt = ctx["n_taps"]
        d_taps, d_conv = d_skips[:nt][::-1], d_skips[nt:][::-1]
        # SemanticTransformer layers
        hp, wp = ctx["hp"], ctx["wp"]
        Ed = C + c.Ct
        tsum = torch.zeros(N * Ed, **f32)                   # column sums of the layers' token gradients: svl_colsum accumulates
        d_cur, d_cur_dtype = d_x, gdt
        for l in reversed(range(c.layers)):
            d_tok1 = torch.empty(B * hp * wp * N, Ed, **f32)
            L.call("svl_unpool_bwd", d_cur, d_cur_dtype, d_cur.shape[-1], d_tok1, Ed, B, N, h, w, C, hp, wp)
            d_tok0 = self._tlayer_bwd(ctx["Sl"][l], d_tok1, p, f"layers.{l}.transformer.", B * hp * wp, N, grads)
            d_prev = torch.empty(nb * hw, C, **f32)
            if d_cur_dtype in (L.F32, L.BF16):          # one pass: f32 copy of the incoming gradient + pooling gradient
                L.call("svl_pool_tokens_bwd_from", d_cur, d_cur_dtype, d_cur.shape[-1], d_tok0, Ed, d_prev, B, N, h, w, C, c.pool)
            else:
                ops.cast(d_cur, d_cur_dtype, d_prev, L.F32, nb * hw, C)
                L.call("svl_pool_tokens_bwd", d_tok0, Ed, d_prev, B, N, h, w, C, c.pool)
            ops.colsum(d_tok0, L.F32, B * hp * wp, N * Ed, tsum, ld=N * Ed)
            d_cur, d_cur_dtype = d_prev, L.F32
        d_t = tsum.view(N, Ed)[:, C:]
        # text projection (weights only; the text embeddings are frozen inputs)
        d_t_pre = d_t * (ctx["t"] > 0)
        ops.wgrad(ops.to_act(d_t_pre.contiguous(), pr), ctx["txt_act"], grads["text_proj.0.weight"], m=c.Ct, n=c.in_dim, precise=pr)
        grads["text_proj.0.bias"].add_(d_t_pre.sum(0))
        # ASPP
        d_x1 = d_cur                                          # f32 accumulator: the residual x + aspp(x) passes the gradient through
        d_praw = ops.new_act(nb * hw, C, pr, dev)
        ops.gn_relu_bwd(d_cur, L.F32, ctx["praw"], L.dtype_of(ctx["praw"]), p["aspp.project.1.weight"], p["aspp.project.1.bias"], ctx["pmean"],
                        ctx["prstd"], d_praw, adt, grads["aspp.project.1.weight"], grads["aspp.project.1.bias"], nb, hw, C, G8)
        gproj = grads["aspp.project.0.weight"].view(C, 5 * C)
        ops.wgrad(d_praw, ctx["cat"], gproj, m=C, n=4 * C, precise=pr, ld_dw=5 * C)
        d_rowb = torch.empty(nb, C, **f32)
        L.call("svl_map_sum", d_praw, adt, d_praw.shape[-1], d_rowb, nb, hw, C, 1.0)
        d_rowb_act = ops.to_act(d_rowb, pr)
        ops.wgrad(d_rowb_act, ctx["pool_act"], gproj[:, 4 * C:], m=C, n=C, precise=pr, ld_dw=5 * C)
        d_pool = torch.empty(nb, C, **f32)
        ops.gemm(d_rowb_act, self._prep(p, "aspp.project.0.weight", "projB_t"), d_pool, n=C, k=C, precise=pr)
        d_graw = ops.new_act(nb, C, pr, dev)
        ops.gn_relu_bwd(d_pool, L.F32, ctx["graw"], L.F32, p["aspp.aspp_convs.4.gap.2.weight"], p["aspp.aspp_convs.4.gap.2.bias"], ctx["gmean"],
                        ctx["grstd"], d_graw, adt, grads["aspp.aspp_convs.4.gap.2.weight"], grads["aspp.aspp_convs.4.gap.2.bias"], nb, 1, C, G8)
        ops.wgrad(d_graw, ctx["gap_act"], grads["aspp.aspp_convs.4.gap.1.weight"].view(C, C), m=C, n=C, precise=pr)
        d_gap = torch.empty(nb, C, **f32)
        ops.gemm(d_graw, self._prep(p, "aspp.aspp_convs.4.gap.1.weight", "lin_t"), d_gap, n=C, k=C, precise=pr)
        d_cat = ops.new_act(nb * hw, 4 * C, pr, dev)
        ops.gemm(d_praw, self._prep(p, "aspp.project.0.weight", "projA_t"), d_cat, n=4 * C, k=C, precise=pr, out_dtype=adt)
        for j, d in enumerate((1,) + c.dil):
            self._conv_gn_bwd(ctx["Sb"][j], d_cat, adt, j * C, ctx["x1"], nb, h, w, C, f"aspp.aspp_convs.{j}.0.weight", f"aspp.aspp_convs.{j}.1", G8, p,
                              grads, d, d_x1, L.F32, accumulate=True)
        L.call("svl_map_bcast_add", d_x1, d_gap, L.F32, C, nb, hw, C, 1.0 / hw)
        # conv1
        d_x1_act = ops.to_act(d_x1, pr)
        dw1, staged = self._wgrad_buf(grads, "conv1.weight", "conv1", (C, 64))
        ops.wgrad(d_x1_act, ctx["col"], dw1, m=C, n=64, precise=pr)
        if not staged:
            grads["conv1.weight"].add_(dw1[:, : c.ks * c.ks].reshape(C, 1, c.ks, c.ks))
        ops.colsum(d_x1, L.F32, nb * hw, C, grads["conv1.bias"])
        self.cache.scatter_grads()                                 # every staged weight gradient of this pass -> grads (one launch)
        if not need_feat_grads:
            return d_taps + [None] + d_conv
        d_col = torch.empty(nb * hw, 64, device=dev, dtype=gtorch)
        ops.gemm(d_x1_act, self._prep(p, "conv1.weight", "conv1_t"), d_col, n=64, k=C, precise=pr)
        Kp = (N + 15) // 16 * 16
        d_sim = ops.new_act(B * hw, Kp, pr, dev)
        L.call("svl_sim_col2im", d_col, gdt, 64, d_sim, adt, d_sim.shape[-1], Kp, B, N, h, w, c.ks)
        # similarity: d(normalised image embedding) = d_sim @ text_n
        d_img_n = torch.empty(B * hw, c.in_dim, device=dev, dtype=gtorch)
        ops.gemm(d_sim, self._text_t_operand(ctx["txt_memo"], ctx["txt_n"], Kp), d_img_n, n=c.in_dim, k=Kp, precise=pr)
        d_emb = torch.empty(B, h, w, c.in_dim, **f32)
        ops.l2norm_bwd(d_img_n, gdt, ctx["img_n"], ctx["inv_img"], d_emb.view(B * hw, c.in_dim))
        return d_taps + [d_emb] + d_conv
