"""torch.profiler (CUPTI) kernel-time table of one training step.  Usage: python scratch/prof_step.py <config 2..5> [batch] [contraction-dump]"""
import sys, re, collections, types, torch
sys.path.insert(0, ".")
import bench
from semivl_b200 import lib as L, ops
from semivl_b200.model import build_model
from semivl_b200.train import OptimCfg, Trainer
from torch.profiler import profile, ProfilerActivity
L.check_device()
c = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
b = int(sys.argv[2]) if len(sys.argv) > 2 else c["batch"]
args = types.SimpleNamespace(nclass=c["nclass"], crop=c["crop"], dataset=c["dataset"])
semivl = c["workload"] == "semivl"
torch.manual_seed(0)
model = build_model(bench.model_cfg(args, False)).cuda()
tr = Trainer(model, OptimCfg())
batch = {k: v.cuda() for k, v in bench.synth_batch(torch, b, c["crop"], c["nclass"], 1234, "cpu", semivl).items()}
step = (lambda: tr.semivl_step(batch)) if semivl else (lambda: tr.supervised_step(batch["img_x"], batch["mask_x"]))
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
print(f"config {sys.argv[1] if len(sys.argv) > 1 else 2} batch {b}: total kernel time {tot/1e3:.2f} ms")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:36]:
    print(f"{k[:44]:44s} {n:4d} {v:9.0f} us {100*v/tot:5.1f}%")
if len(sys.argv) > 3:
    ops.PROFILE = []
    step(); torch.cuda.synchronize()
    by = collections.OrderedDict()
    for a, bv, f, lab in ops.PROFILE:
        e = by.setdefault(lab, [0, 0.0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(bv); e[2] += f
    ops.PROFILE = None
    print("contractions by shape (count, total us, TF/s):")
    for lab, (n, t, f) in sorted(by.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{n:4d} {t*1e3:10.1f} us {f/t/1e9:8.1f} TF/s  {lab}")
