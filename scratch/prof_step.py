import sys, re, collections, torch
sys.path.insert(0, ".")
import bench
from semivl_b200 import lib as L
from semivl_b200.model import build_model
from semivl_b200.train import OptimCfg, Trainer
from torch.profiler import profile, ProfilerActivity
L.check_device()
wl = sys.argv[1] if len(sys.argv) > 1 else "supervised"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
model = build_model(bench.model_cfg(512, 21, False)).cuda()
tr = Trainer(model, OptimCfg())
batch = {k: v.cuda() for k, v in bench.synth_batch(torch, b, 512, 21, 1234, "cuda", wl == "semivl").items()}
step = (lambda: tr.semivl_step(batch)) if wl == "semivl" else (lambda: tr.supervised_step(batch["img_x"], batch["mask_x"]))
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        m = re.search(r'(\w+_kernel)', e.name)
        k = m.group(1) if m else e.name[:40]
        agg[k][0] += 1; agg[k][1] += e.device_time
tot = sum(v for _, v in agg.values())
print(f"total kernel time {tot/1e3:.2f} ms")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:32]:
    print(f"{k[:44]:44s} {n:4d} {v:9.0f} us {100*v/tot:5.1f}%")
