"""Peak memory and step time of the north-star configurations 3/4/5 at growing per-GPU batch.
Usage: python scratch/mem_probe.py <config> [batches...]   config 3: supervised 801^2 N=19; 4: SemiVL 512^2 N=150; 5: SemiVL 641^2 N=81."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from semivl_b200.model import build_model
from semivl_b200.train import OptimCfg, Trainer
cfgno = int(sys.argv[1])
crop, N, ds, semivl, default_b = {3: (801, 19, 'cityscapes', False, [1, 2]), 4: (512, 150, 'ade', True, [1, 2, 4, 8]),
                                  5: (641, 81, 'coco', True, [1, 2, 4, 8, 16]), 2: (512, 21, 'pascal', False, [16])}[cfgno]
batches = [int(a) for a in sys.argv[2:]] or default_b
cfg = dict(model='mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb', nclass=N, crop_size=crop, dataset=ds, text_embedding_variant='single',
           mcc_text='single', pl_text='single', clip_encoder='mcvit16', disable_dropout=True, fp_rate=0.5, model_args=dict(pretrained=None),
           clip_encoder_args=dict(pretrained=None), precise=False)
torch.manual_seed(0)
model = build_model(cfg).cuda()
tr = Trainer(model, OptimCfg(lr=1e-4, total_iters=100000))
prev = None
for b in batches:
    if prev is not None and prev[1] / prev[0] * b > 165e9:
        print(f"config {cfgno} b={b}: skipped, predicted peak {prev[1] / prev[0] * b / 1e9:.0f} GB"); continue
    batch = {k: v.cuda() for k, v in bench.synth_batch(torch, b, crop, N, 1234, "cpu", semivl).items()}
    torch.cuda.reset_peak_memory_stats()
    try:
        for _ in range(2):
            (tr.semivl_step(batch) if semivl else tr.supervised_step(batch["img_x"], batch["mask_x"]))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = (tr.semivl_step(batch)[0] if semivl else tr.supervised_step(batch["img_x"], batch["mask_x"]))
        e1.record(); torch.cuda.synchronize()
        peak = torch.cuda.max_memory_allocated()
        prev = (b, peak)
        print(f"config {cfgno} crop {crop} N {N} b={b}: {e0.elapsed_time(e1) / 3:.1f} ms/step, {b / (e0.elapsed_time(e1) / 3) * 1e3:.2f} img/s, "
              f"peak {peak / 1e9:.1f} GB, loss {float(out):.4f}", flush=True)
    except torch.OutOfMemoryError as e:
        print(f"config {cfgno} b={b}: OOM ({str(e)[:80]})", flush=True)
        break
    del batch
    torch.cuda.empty_cache()
