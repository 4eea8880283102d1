"""Arithmetic-mode sweep: logit parity against the CPU oracle at 512x512 / N=21 and the eager step time at batch 16 for
bf16 everywhere, split-bf16 x3 in the head only, in the encoder only, and everywhere.  Usage: python scratch/mode_sweep.py [crop] [batch]"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
from semivl_b200.model import build_model
from semivl_b200.train import OptimCfg, Trainer
crop = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bt = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = 21
mc = O.ModelCfg(img_size=crop, num_classes=N)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
g = torch.Generator().manual_seed(17)
img1 = torch.randn(1, 3, crop, crop, generator=g)
torch.set_num_threads(os.cpu_count())
t0 = time.time()
with torch.no_grad():
    ref, ref_low = O.model_forward(img1, sd, text, mc, return_lowres=True)
print(f"oracle forward {time.time() - t0:.1f} s; logit range {ref.abs().max().item():.4f}")
top2 = ref.topk(2, dim=1).values
margin = top2[:, 0] - top2[:, 1]
imgs = torch.randn(bt, 3, crop, crop, generator=g).cuda()
mask = torch.randint(0, N, (bt, crop, crop), generator=g).cuda()
for mode in (False, 'head', 'encoder', True):
    cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=N, crop_size=crop, dataset='pascal', text_embedding_variant='single',
               mcc_text='single', pl_text='single', clip_encoder=None, disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None), precise=mode)
    m = build_model(cfg); m.load_state_dict(sd); m = m.cuda()
    with torch.no_grad():
        y = m(img1.cuda()).cpu()
    err = (y - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    same = (y.argmax(1) == ref.argmax(1))
    dec = margin > 2 * err
    tr = Trainer(m, OptimCfg())
    for _ in range(3):
        tr.supervised_step(imgs, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.supervised_step(imgs, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"mode={str(mode):8s} logits rel {rel:.3e}  argmax agree {same.float().mean().item():.5f} (decidable {dec.float().mean().item():.3f}, "
          f"exact on decidable {bool(same[dec].all())})  eager step {ms:.1f} ms = {bt / ms * 1e3:.0f} img/s")
    del tr, m
    torch.cuda.empty_cache()
