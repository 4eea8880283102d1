"""CPU-only checks: the C-ABI library loads and exports every symbol include/semivl_b200.h declares, the ctypes descriptors
match the C structs, the mmseg-style registry / Config / build_model mirror builds the reference's parameter inventory, and
the data-parallel gradient exchange is exact under gloo with world_size 2.  (No compute kernels are called without a GPU.)"""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    return ge.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "semivl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(svl_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 40
    lib = ctypes.CDLL(built)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    from semivl_b200 import lib as L
    assert L.version() == 100
    bound = set(L._PROTOS) | {"svl_gemm", "svl_wgrad", "svl_last_error", "svl_version", "svl_check_device", "svl_attention_bwd_workspace", "svl_gn_workspace",
                                 "svl_bn_workspace", "svl_conv_gn_splits"}
    assert names == bound, (names - bound, bound - names)


def test_descriptor_layout_matches_c(built, tmp_path):
    """sizeof / offsetof of the two descriptor structs as seen by a C compiler == the ctypes mirror."""
    from semivl_b200 import lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "semivl_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(svl_gemm_desc), offsetof(svl_gemm_desc, out), offsetof(svl_gemm_desc, block_n),'
                   'sizeof(svl_wgrad_desc), offsetof(svl_wgrad_desc, dw), offsetof(svl_wgrad_desc, splits));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(L.GemmDesc), L.GemmDesc.out.offset, L.GemmDesc.block_n.offset,
            ctypes.sizeof(L.WgradDesc), L.WgradDesc.dw.offset, L.WgradDesc.splits.offset]
    assert got == want


def test_no_gpu_calls_fail_loudly(built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semivl_b200 import lib as L
    with pytest.raises(L.SvlError):
        L.check_device()


def test_registry_and_build_model_mirror_reference_inventory(built):
    from oracle import semivl_oracle as O
    from semivl_b200.model import build_model
    from semivl_b200.registry import BACKBONES, HEADS, SEGMENTORS, Config
    assert "MaskClipVisionTransformer" in BACKBONES and "VLGHead" in HEADS and "VLM" in SEGMENTORS
    cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=21, crop_size=224, dataset='pascal', text_embedding_variant='single',
               mcc_text='concept4_single', pl_text='single', clip_encoder='mcvit16', disable_dropout=True, fp_rate=0.5,
               model_args=dict(pretrained=None), clip_encoder_args=dict(pretrained=None))
    m = build_model(cfg)
    sd = m.state_dict()
    shapes = O.param_shapes(O.ModelCfg(img_size=224))
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    # SURVEY.md §8c probe numbers of the reference: total / backbone / trainable backbone / head parameters at 224^2
    n = lambda it: sum(p.numel() for p in it)
    assert n(m.parameters()) == 175238369
    assert n(m.backbone.parameters()) == 86192640 + 0 or abs(n(m.backbone.parameters()) - 86.193e6) < 2e3
    assert abs(n(p for p in m.backbone.parameters() if p.requires_grad) - 28.500e6) < 1e3
    assert n(m.decode_head.parameters()) == 2217185
    trainable = [k for k, p in m.named_parameters() if p.requires_grad and not k.startswith("clip_encoder")]
    assert len(trainable) == 117                    # tensors that receive gradients in the reference
    assert m.loaded_mcc_text_feat.shape == (98, 512)
    from semivl_b200.text_embeddings import get_class_to_concept_idxs
    g = get_class_to_concept_idxs(m.load_mcc_text_embedding)
    assert len(g) == 21 and g[0] == list(range(45)) and g[20][-1] == 97
    c = Config.fromfile(os.path.join(ROOT, "semivl_b200", "configs", "_base_", "models", "mcvit16.py"))
    assert c.backbone.out_indices is None and c["backbone"]["embed_dims"] == 768


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semivl_b200.train import allreduce_sum_
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)
    allreduce_sum_(flat)
    q.put((rank, flat))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_gloo_world2(built):
    """N-rank summed flat gradient == sum of the per-rank gradients (semivl.py:139-140 DDP semantics; the mean is folded into AdamW)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    want = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    assert torch.allclose(res[0], want) and torch.allclose(res[1], want)


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semivl_b200.train import GradExchange
    flat = torch.randn(1000, generator=torch.Generator().manual_seed(100 + rank))
    ex = GradExchange(flat)
    ex.begin()
    ex.reduce(900, 1000)          # "head" slice first, then encoder buckets from the last layers down, like Trainer does
    ex.reduce(600, 900)
    ex.reduce(250, 600)
    ex.finish()                   # the uncovered prefix [0, 250)
    from semivl_b200.evaluate import reduce_counts          # mIoU histograms: one all-reduce of the [3, K] int64 matrix (supervised.py:158-160)
    counts = reduce_counts(torch.arange(3 * 4, dtype=torch.int64).view(3, 4) * (rank + 1))
    assert torch.equal(counts, torch.arange(3 * 4, dtype=torch.int64).view(3, 4) * 3)
    q.put((rank, flat))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_gloo_world2(built):
    """GradExchange: slices exchanged as they become final + finish() of the remainder == one all-reduce of the whole buffer."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    want = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    assert torch.allclose(res[0], want) and torch.allclose(res[1], want)


def test_poly_lr_schedule_matches_reference_order(built):
    """semivl.py:338-345 rewrites the rate AFTER optimizer.step() from the 0-based index of the iteration just run, so optimizer
    step `it` (1-based) runs at poly_lr(it - 2) (and steps 1, 2 at the initial rate)."""
    from oracle import semivl_oracle as O
    from semivl_b200.train import OptimCfg, Trainer
    tr = Trainer.__new__(Trainer)
    tr.opt = OptimCfg(lr=1e-4, total_iters=50)
    lr, seen = 1e-4, []
    for i in range(10):                     # the reference loop: step with the current rate, then rewrite it from index i
        seen.append(lr)
        lr = O.poly_lr(1e-4, i, 50)
    assert np.allclose([tr.lr_at(it) for it in range(1, 11)], seen, rtol=1e-12)
    a, b_, bc1, bc2s, c_ = tr._step_scalars(3)
    assert np.isclose(c_, seen[2] * 0.1)
    assert np.isclose(a, seen[2] * 0.01) and np.isclose(b_, seen[2] * 10.0)
    assert np.isclose(bc1, 1 - 0.9 ** 3) and np.isclose(bc2s, (1 - 0.999 ** 3) ** 0.5)


def test_clip_checkpoint_conversion_and_pretrained_load(built, tmp_path):
    """third_party/maskclip/convert_clip_weights.py:27-64 + maskclip_vit.py:378-410: an OpenAI-CLIP-named visual state dict is renamed,
    saved, and loaded by `pretrained=` into a model with a DIFFERENT token grid (position table resized on load, cls entry kept)."""
    from oracle import semivl_oracle as O
    from semivl_b200.convert_clip_weights import clip_visual_to_mmseg, rename_visual_key
    from semivl_b200.model.maskclip_vit import MaskClipVisionTransformer
    assert rename_visual_key("transformer.resblocks.7.attn.in_proj_weight") == "layers.7.attn.attn.in_proj_weight"
    assert rename_visual_key("transformer.resblocks.0.mlp.c_fc.bias") == "layers.0.ffn.layers.0.0.bias"
    assert rename_visual_key("transformer.resblocks.11.mlp.c_proj.weight") == "layers.11.ffn.layers.1.weight"
    assert rename_visual_key("transformer.resblocks.3.ln_2.weight") == "layers.3.ln2.weight"
    assert rename_visual_key("ln_pre.bias") == "ln0.bias" and rename_visual_key("ln_post.weight") == "ln1.weight"
    E, layers, g0 = 64, 2, 14                                # a small tower with CLIP's structure: 224 / 16 = 14 x 14 positions
    gen = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=gen).half()     # CLIP checkpoints are fp16
    clip = {"visual.class_embedding": r(E), "visual.positional_embedding": r(g0 * g0 + 1, E), "visual.conv1.weight": r(E, 3, 16, 16),
            "visual.ln_pre.weight": r(E), "visual.ln_pre.bias": r(E), "visual.ln_post.weight": r(E), "visual.ln_post.bias": r(E),
            "visual.proj": r(E, 512), "logit_scale": r(1), "token_embedding.weight": r(10, 8)}
    for i in range(layers):
        pre = f"visual.transformer.resblocks.{i}."
        clip.update({pre + "attn.in_proj_weight": r(3 * E, E), pre + "attn.in_proj_bias": r(3 * E), pre + "attn.out_proj.weight": r(E, E),
                     pre + "attn.out_proj.bias": r(E), pre + "ln_1.weight": r(E), pre + "ln_1.bias": r(E), pre + "ln_2.weight": r(E),
                     pre + "ln_2.bias": r(E), pre + "mlp.c_fc.weight": r(4 * E, E), pre + "mlp.c_fc.bias": r(4 * E),
                     pre + "mlp.c_proj.weight": r(E, 4 * E), pre + "mlp.c_proj.bias": r(E)})
    conv = clip_visual_to_mmseg(clip, backbone_prefix=True)
    assert all(v.dtype == torch.float32 for v in conv["state_dict"].values())
    path = str(tmp_path / "clip2mmseg.pth")
    torch.save(conv, path)
    kw = dict(patch_size=16, patch_bias=False, in_channels=3, embed_dims=E, num_layers=layers, num_heads=1, mlp_ratio=4, out_indices=(0, layers),
              qkv_bias=True, with_cls_token=True, output_cls_token=False, norm_cfg=dict(type='LN', eps=1e-6), act_cfg=dict(type='GELU'),
              patch_norm=False, pre_norm=True, final_norm=True, return_clip_embed=True, return_qkv=True, interpolate_mode='bicubic', num_fcs=2)
    try:
        m = MaskClipVisionTransformer(img_size=(96, 96), pretrained=path, **kw)
    except (AssertionError, TypeError, ValueError) as e:         # the engine is specialised for CLIP ViT-B/16 widths
        pytest.skip(f"small tower not constructible: {e}")
    m.init_weights()
    rep = m.load_report
    assert not rep.unexpected_keys, rep.unexpected_keys
    assert not rep.missing_keys, rep.missing_keys
    sd = m.state_dict()
    assert torch.equal(sd["layers.1.ffn.layers.0.0.weight"], clip["visual.transformer.resblocks.1.mlp.c_fc.weight"].float())
    assert torch.equal(sd["cls_token"], clip["visual.class_embedding"].float()[None, None])
    assert sd["proj.weight"].shape == (512, E, 1, 1) and torch.equal(sd["proj.weight"][:, :, 0, 0], clip["visual.proj"].float().t())
    want = O.resize_pos_embed(clip["visual.positional_embedding"].float()[None], (6, 6), (g0, g0))
    assert sd["pos_embed"].shape == (1, 37, E) and torch.allclose(sd["pos_embed"], want, atol=1e-6)


def test_input_stage_draws_and_restatement_match_reference(built, golden_dir):
    """tests/golden/input_stage.npz holds outputs of the reference's crop / hflip / normalize / obtain_cutmix_box under seeded RNGs:
    the host-side samplers of semivl_b200.input_pipeline must make the same draws in the same order, and the oracle restatement of the
    per-pixel work must reproduce the tensors bit for bit."""
    import random
    from oracle import semivl_oracle as O
    from semivl_b200 import input_pipeline as P
    g = dict(np.load(os.path.join(golden_dir, "input_stage.npz"), allow_pickle=False))
    for i, case in enumerate(g["cases"]):
        h, w, size, pad = (int(v) for v in str(case).split("|"))
        random.seed(500 + i)
        x0, y0 = P.sample_crop(w, h, size)
        flip = P.sample_hflip(0.5)
        t, lab, ign = O.crop_flip_normalize_reference(g[f"img{i}"], g[f"mask{i}"], size, x0, y0, flip, pad)
        assert np.array_equal(t.numpy(), g[f"out_img{i}"]), case
        assert np.array_equal(lab.numpy(), g[f"out_mask{i}"].astype(np.int64)) and np.array_equal(ign.numpy(), g[f"out_ign{i}"].astype(np.int64))
    for j, want in enumerate(g["boxes"]):
        random.seed(900 + j)
        np.random.seed(900 + j)
        params = P.sample_cutmix_box(48, p=0.5)
        box = np.zeros((48, 48), np.uint8)
        if params is not None:
            x, y, bw, bh = params
            box[y:y + bh, x:x + bw] = 1
        assert np.array_equal(box, want), j


def test_bench_reference_arm_contract(built):
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the contract's keys, on a small crop."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-crop", "64"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_conv_encoder_oracle_layer1_matches_torchvision(built):
    """oracle/resnetv1c_oracle.py (groundwork for the skr04 conv encoder): the residual stage is pinned against torchvision's ResNet-101
    `layer1` with identical weights, in training (batch statistics) and eval (running statistics) mode; the deep stem's shapes are checked."""
    torchvision = pytest.importorskip("torchvision")
    from oracle import resnetv1c_oracle as R
    gen = torch.Generator().manual_seed(0)
    tv = torchvision.models.resnet101(weights=None).layer1
    sd = {}
    for k, v in tv.state_dict().items():
        if k.endswith("num_batches_tracked"):
            continue
        t = torch.randn(v.shape, generator=gen) * (0.1 if v.dim() > 1 else 0.5)
        sd[k] = t.abs() + 0.5 if k.endswith("running_var") or (k.endswith("weight") and v.dim() == 1) else t
    tv.load_state_dict(sd, strict=False)
    p = {"conv_encoder.layer1." + k: v for k, v in sd.items()}
    assert set(p) == {k for k in R.param_shapes() if ".layer1." in k}
    x = torch.randn(2, 64, 12, 10, generator=gen)
    for training in (False, True):           # eval first: a training-mode forward updates torchvision's running statistics in place
        tv.train(training)
        with torch.no_grad():
            want = tv(x)
            got = x
            for i in range(3):
                got = R.bottleneck_forward(got, p, f"conv_encoder.layer1.{i}.", training)
        assert torch.allclose(got, want, atol=1e-5, rtol=1e-5), training
    shapes = R.param_shapes()
    full = {k: torch.randn(s, generator=gen).abs() + 0.1 if k.endswith("running_var") else torch.randn(s, generator=gen) * 0.1 for k, s in shapes.items()}
    out = R.conv_encoder_forward(torch.randn(1, 3, 64, 64, generator=gen), full)
    assert len(out) == 1 and tuple(out[0].shape) == (1, 256, 16, 16)


def test_trainer_exchange_schedule_covers_the_flat_buffer(built):
    """The order in which Trainer hands slices of the flat gradient buffer to GradExchange during a backward pass (head first, then encoder
    layers 9-11, 6-8, 3-5, then the rest at finish): disjoint, inside the buffer, each encoder bucket cut at a layer boundary, and together
    with finish() covering every trainable element exactly once."""
    from semivl_b200.model import build_model
    from semivl_b200.train import OptimCfg, Trainer
    cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=21, crop_size=64, dataset='pascal', text_embedding_variant='single',
               mcc_text='single', pl_text='single', clip_encoder=None, disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None), precise=False)
    tr = Trainer(build_model(cfg), OptimCfg())
    n = tr.n_bb + tr.n_hd
    assert tr.g_flat.numel() == n and tr.p_flat.numel() == n

    class Rec:
        def __init__(self):
            self.ranges = []

        def reduce(self, lo, hi):
            self.ranges.append((lo, hi))

    rec = Rec()
    tr.exchange = rec
    tr._head_grads_final()                                  # after head.backward
    for i in reversed(range(12)):                            # the encoder backward walks the layers from the last one down
        tr._layer_grads_final(i)
    assert rec.ranges[0] == (tr.n_bb, n)
    enc = rec.ranges[1:]
    assert [hi for _, hi in enc] == [tr._layers_end] + [lo for lo, _ in enc[:-1]]            # contiguous, descending
    assert [lo for lo, _ in enc] == [tr.layer_lo[9], tr.layer_lo[6], tr.layer_lo[3]]
    covered = sum(hi - lo for lo, hi in rec.ranges)
    rest = tr.layer_lo[3]                                    # position table + layers 0-2: exchanged by finish()
    assert covered + rest == n
    # every trainable backbone tensor is an attention or position-embedding tensor (vlm.py:80-88 freeze filter)
    assert all(("attn" in k) or ("pos_embed" in k) for k in tr.g_bb)


def _bcast_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semivl_b200.model import build_model
    from semivl_b200.train import OptimCfg, Trainer
    torch.manual_seed(1000 + rank)                           # ranks deliberately initialise the random head differently
    cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=21, crop_size=64, dataset='pascal', text_embedding_variant='single',
               mcc_text='single', pl_text='single', clip_encoder=None, disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None), precise=False)
    model = build_model(cfg)
    before = torch.cat([p.detach().reshape(-1) for p in model.decode_head.parameters()]).clone()
    tr = Trainer(model, OptimCfg())
    idx = torch.randint(0, tr.p_flat.numel(), (4096,), generator=torch.Generator().manual_seed(5))
    digest = tr.p_flat[idx].tolist() + [tr.p_flat.double().sum().item(), tr.p_flat[tr.n_bb:].double().abs().sum().item()]
    q.put((rank, before[:64].tolist(), digest, model.decode_head.conv1.weight.detach().reshape(-1)[:16].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_broadcasts_initial_parameters_gloo_world2(built):
    """DDP broadcasts rank 0's parameters at construction (semivl.py:139-140): two ranks seeded differently must leave Trainer.__init__ with
    identical flat parameter buffers (= rank 0's), and the module parameters must be views of that buffer."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: rest for r, *rest in (q.get(timeout=240) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
    assert res[0][0] != res[1][0], "the two ranks were meant to start from different random heads"
    assert res[0][1] == res[1][1]                 # sampled entries + sums of the flat parameter buffer
    assert res[0][2] == res[1][2]                 # the module parameters are views of the (broadcast) flat buffer


def _syncbn_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semivl_b200.model.resnet import sync_sums
    sums = torch.arange(2 * 8, dtype=torch.float32).view(2, 8) * (rank + 1)
    count = sync_sums(sums, 100 * (rank + 1))
    q.put((rank, sums.tolist(), count))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_statistics_exchange_gloo_world2(built):
    """SyncBN hook of the conv encoder (semivl.py:136 convert_sync_batchnorm): the per-channel sums of every BatchNorm and the
    row count are SUM all-reduced over the ranks before mean / rstd are formed."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: rest for r, *rest in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
    want = (torch.arange(2 * 8, dtype=torch.float32).view(2, 8) * 3).tolist()
    assert res[0][0] == want and res[1][0] == want
    assert res[0][1] == 300.0 and res[1][1] == 300.0


def test_lr_warmup_and_scheduler_max_iters(built):
    """semivl.py:186,338-345: linear warm-up `lr * (1 - (1 - i / warmup_iters) * (1 - warmup_ratio))` while i < warmup_iters, then poly over
    scheduler_max_iters (which may exceed total_iters), rewritten AFTER the optimizer step of iteration i."""
    from semivl_b200.train import OptimCfg, Trainer
    tr = Trainer.__new__(Trainer)
    tr.opt = OptimCfg(lr=2e-4, total_iters=40, warmup_iters=5, warmup_ratio=1e-3, scheduler_max_iters=80)
    lr, seen = 2e-4, []
    for i in range(12):                     # the reference loop: step with the current rate, then rewrite it from index i
        seen.append(lr)
        if i < 5:
            lr = 2e-4 * (1 - (1 - i / 5) * (1 - 1e-3))
        else:
            lr = 2e-4 * (1 - i / 80) ** 0.9
    assert np.allclose([tr.lr_at(it) for it in range(1, 13)], seen, rtol=1e-12)


def test_operand_index_maps_and_gradient_scatter_semantics(built):
    """Host side of the one-launch parameter plumbing (engine/vit.py WeightCache, svl_param_jobs): every operand layout of the head / encoder
    is a pure gather of the parameter, the index map derived by pushing an arange through the layout function reproduces it (-1 = zero
    padding), and scattering a staged gradient through the same map equals the reference-layout update the engines used to do with
    permute + add_ (vlg_head.py parameters keep the reference's [Co, Ci, kh, kw] / [Ci, Co, 2, 2] layouts)."""
    import torch
    from semivl_b200.engine.head import HeadEngine
    from semivl_b200.engine.vit import WeightCache
    g = torch.Generator().manual_seed(3)
    shapes = {"lin": (24, 40), "lin_t": (24, 40), "conv": (32, 16, 3, 3), "conv_t": (32, 16, 3, 3), "convT": (16, 8, 2, 2), "convT_t": (16, 8, 2, 2),
              "conv1": (8, 1, 5, 5), "conv1_t": (8, 1, 5, 5), "projA": (8, 40, 1, 1), "projA_t": (8, 40, 1, 1), "projB": (8, 40, 1, 1),
              "projB_t": (8, 40, 1, 1), "out1": (1, 32, 3, 3)}
    for kind, shape in shapes.items():
        fn = HeadEngine._layout(kind)
        prm = torch.randn(*shape, generator=g)
        idx, lshape = WeightCache._index_map(prm, fn)
        lay = fn(prm).contiguous()
        assert tuple(lay.shape) == lshape
        flat = prm.reshape(-1)
        got = torch.where(idx >= 0, flat[idx.clamp_min(0).long()], torch.zeros(()))
        assert torch.equal(got, lay.reshape(-1)), kind
        real = idx[idx >= 0]
        assert real.unique().numel() == real.numel(), kind          # injective: the scatter needs no atomics
    # gradient scatter of a conv weight gradient staged as [(tap, co), ci] == grads.add_(dw.view(ks, ks, co, ci).permute(2, 3, 0, 1))
    co, ci = 32, 16
    dw = torch.randn(9, co, ci, generator=g)
    grad = torch.randn(co, ci, 3, 3, generator=g)
    want = grad + dw.view(3, 3, co, ci).permute(2, 3, 0, 1)
    idx, _ = WeightCache._index_map(grad, HeadEngine._layout("conv"))
    got = grad.clone().reshape(-1)
    got.index_add_(0, idx.long(), dw.reshape(-1))
    assert torch.allclose(got.view_as(grad), want)
    # transposed conv: dwq [4, ci, cu] staged in the "convT_t" layout of a [ci, cu, 2, 2] parameter
    ci, cu = 16, 8
    dwq = torch.randn(4, ci, cu, generator=g)
    grad = torch.zeros(ci, cu, 2, 2)
    idx, _ = WeightCache._index_map(grad, HeadEngine._layout("convT_t"))
    got = grad.clone().reshape(-1)
    got.index_add_(0, idx.long(), dwq.reshape(-1))
    assert torch.allclose(got.view_as(grad), dwq.view(2, 2, ci, cu).permute(2, 3, 0, 1))
    # the padded conv1 layout drops its padding columns on the way back
    C, ks = 8, 5
    dw1 = torch.randn(C, 64, generator=g)
    grad = torch.zeros(C, 1, ks, ks)
    idx, _ = WeightCache._index_map(grad, HeadEngine._layout("conv1"))
    keep = idx >= 0
    got = grad.clone().reshape(-1)
    got.index_add_(0, idx[keep].long(), dw1.reshape(-1)[keep])
    assert torch.allclose(got.view_as(grad), dw1[:, : ks * ks].reshape(C, 1, ks, ks))


def test_conv_gn_splits_plan_through_the_c_abi(built):
    """svl_conv_gn_splits() is pure host arithmetic on the descriptor (no CUDA call): the number of GroupNorm partial sums per map the rolling
    convolution kernel writes depends on the map geometry only -- 16-row units x 128-pixel columns x 16 epilogue warps, or x 8 warps per map in
    the dual-map form (maps up to 64 pixels wide) -- and is 0 for every problem the kernel does not take."""
    from semivl_b200 import lib as L

    def desc(nb, h, w, cin, cout, taps=9, dil=1, out_dtype=L.BF16, ldc=None):
        d = L.GemmDesc()
        d.a_conv, d.nb, d.h, d.w, d.m = 1, nb, h, w, nb * h * w
        d.lda, d.ldb, d.b_rows = cin, cin, taps * cout
        d.n, d.k_per_tap, d.num_taps = cout, cin, taps
        k = 0
        for i in range(3):
            for j in range(3):
                if k < taps:
                    d.tap_dy[k], d.tap_dx[k], d.tap_b_row[k] = (i - 1) * dil, (j - 1) * dil, k * cout
                    k += 1
        d.out, d.out_dtype, d.ldc = 4096, out_dtype, ldc if ldc is not None else cout      # a fake, aligned pointer: nothing is dereferenced
        return d

    f = L.lib().svl_conv_gn_splits
    assert f(ctypes.byref(desc(336, 128, 128, 32, 32))) == 8 * 1 * 16          # 8 units of 16 rows, one 128-pixel column
    assert f(ctypes.byref(desc(5, 40, 204, 64, 64))) == 3 * 2 * 16             # 40 rows -> 3 units, 204 pixels -> 2 columns
    assert f(ctypes.byref(desc(336, 64, 64, 128, 64))) == 4 * 8                # dual-map form: 4 units x 8 warps per map
    assert f(ctypes.byref(desc(2, 7, 33, 32, 32))) == 1 * 8
    assert f(ctypes.byref(desc(2, 64, 80, 32, 32))) == 0                       # 65..95-pixel rows stay on the generic engine
    assert f(ctypes.byref(desc(2, 64, 32, 32, 32))) == 0                       # too narrow
    assert f(ctypes.byref(desc(2, 64, 128, 32, 128))) == 0                     # 3 x 128 accumulator columns do not fit one MMA
    assert f(ctypes.byref(desc(2, 64, 128, 32, 32, dil=2))) == 0               # dilated taps
    assert f(ctypes.byref(desc(2, 64, 128, 32, 32, taps=4))) == 0
    assert f(ctypes.byref(desc(2, 64, 128, 32, 32, out_dtype=L.F32))) == 0     # bf16 outputs only
    # batch independence: the plan does not depend on the number of maps
    assert f(ctypes.byref(desc(1, 128, 128, 64, 32))) == f(ctypes.byref(desc(300, 128, 128, 64, 32)))
