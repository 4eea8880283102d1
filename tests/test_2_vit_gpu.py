"""GPU parity of the ViT encoder engine (forward + backward) against the CPU oracle (oracle/semivl_oracle.py, itself
pinned to the unmodified reference by tests/golden).  precise mode: rel 2e-4 on features, 2e-3 on gradients;
fast (bf16) mode: rel 3e-2 / 6e-2 (stated, measured values are printed)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("crop,b", [(64, 2), (72, 1), (224, 1)])
@pytest.mark.parametrize("precise", [True, False])
def test_vit_forward_backward(crop, b, precise):
    from oracle import semivl_oracle as O
    from semivl_b200 import lib
    from semivl_b200.engine.vit import VitCfg, VitEngine
    lib.check_device()
    mc = O.ModelCfg(img_size=crop, num_classes=21)
    sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
    pcpu = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("backbone.")}
    g = torch.Generator().manual_seed(crop)
    img = torch.randn(b, 3, crop, crop, generator=g)
    feats_ref, glob_ref = O.vit_forward(img, pcpu, mc)
    ws = [torch.randn(f.shape, generator=g) / f.shape[1] ** 0.5 for f in feats_ref]
    loss = sum((f * w).sum() for f, w in zip(feats_ref, ws))
    loss.backward()

    eng = VitEngine(VitCfg(img_size=crop), precise=precise)
    p = {k[len("backbone."):]: v.detach().cuda() for k, v in pcpu.items()}
    feats, glob, ctx = eng.forward(img.cuda(), p, need_grad=True)
    tol_f, tol_g = (2e-4, 2e-3) if precise else (3e-2, 6e-2)
    for f, fr in zip(feats, feats_ref):
        r = _rel(f.permute(0, 3, 1, 2), fr.detach())
        print(f"crop {crop} precise {precise}: feat rel {r:.2e}")
        assert r < tol_f
    assert _rel(glob, glob_ref.detach()) < tol_f
    grads = {k: torch.zeros_like(v) for k, v in p.items() if ("attn" in k or "pos_embed" in k)}
    eng.backward(ctx, [w.permute(0, 2, 3, 1).contiguous().cuda() for w in ws], p, grads)
    worst = 0.0
    for k, gv in grads.items():
        gr = pcpu["backbone." + k].grad
        assert gr is not None, k
        r = _rel(gv, gr)
        worst = max(worst, r)
        assert r < tol_g, (k, r)
    print(f"crop {crop} precise {precise}: worst grad rel {worst:.2e}")
