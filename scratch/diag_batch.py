"""Is an image's output independent of its batch?  Full config-2 batch (16 x 512^2) vs sub-batches, stage by stage, both modes,
and image 5 / 13 of the full batch against the CPU oracle."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import semivl_oracle as O
from semivl_b200.model import build_model
torch.set_num_threads(os.cpu_count())
def cfg(precise):
    return dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=21, crop_size=512, dataset='pascal', text_embedding_variant='single',
                mcc_text='single', pl_text='single', clip_encoder=None, disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None), precise=precise)
mc = O.ModelCfg(img_size=512, num_classes=21)
sd = O.fixture_state_dict(O.param_shapes(mc, with_clip_encoder=False), seed=0)
text = torch.from_numpy(np.load("semivl_b200/configs/_base_/datasets/text_embedding/voc12_wbg_single.npy"))
img = torch.randn(16, 3, 512, 512, generator=torch.Generator().manual_seed(17))
with torch.no_grad():
    ref = {i: O.model_forward(img[i:i + 1], sd, text, mc, return_lowres=True)[1] for i in (5, 13)}
for precise in (False, True):
    m = build_model(cfg(precise)); m.load_state_dict(sd); m = m.cuda()
    x = img.cuda()
    with torch.no_grad():
        f_full = m.extract_feat(x)
        f_part = m.extract_feat(x[5:7].contiguous())
        for a, b in zip(f_full[0][0], f_part[0][0]):
            print(f"precise={precise} backbone feat {tuple(a.shape)} max|full[5:7]-part| = {(a[5:7].float() - b.float()).abs().max().item():.3e} (range {a.abs().max().item():.3e})")
        low_full = m.forward_lowres(x).float()
        low_part = m.forward_lowres(x[5:7].contiguous()).float()
        low_part1 = m.forward_lowres(x[5:6].contiguous()).float()
        low_again = m.forward_lowres(x).float()
        r = low_full.abs().max().item()
        print(f"precise={precise} lowres logits range {r:.3f}: full vs b2 {(low_full[5:7] - low_part).abs().max().item() / r:.3e}  b2 vs b1 {(low_part[:1] - low_part1).abs().max().item() / r:.3e}"
              f"  full run-to-run {(low_full - low_again).abs().max().item() / r:.3e}")
        for i, rf in ref.items():
            rr = rf.abs().max().item()
            print(f"precise={precise} image {i}: full-batch vs oracle {(low_full[i:i+1].cpu() - rf).abs().max().item() / rr:.3e}   "
                  + (f"b2 vs oracle {(low_part[i-5:i-4].cpu() - rf).abs().max().item() / rr:.3e}" if i == 5 else ""))
    del m
    torch.cuda.empty_cache()
