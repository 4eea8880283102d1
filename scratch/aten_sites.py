"""Which lines of semivl_b200/ still launch ATen kernels inside one supervised step?  (TorchDispatchMode: every non-view aten op on a
CUDA tensor, attributed to the innermost semivl_b200 frame.)  Usage: python scratch/aten_sites.py [config]"""
import sys, collections, traceback, types, torch
sys.path.insert(0, ".")
import bench
from semivl_b200 import lib as L
from semivl_b200.model import build_model
from semivl_b200.train import OptimCfg, Trainer
from torch.utils._python_dispatch import TorchDispatchMode
L.check_device()
c = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
args = types.SimpleNamespace(nclass=c["nclass"], crop=c["crop"], dataset=c["dataset"])
torch.manual_seed(0)
model = build_model(bench.model_cfg(args, False)).cuda()
tr = Trainer(model, OptimCfg())
batch = {k: v.cuda() for k, v in bench.synth_batch(torch, 2, c["crop"], c["nclass"], 1234, "cpu", False).items()}
step = lambda: tr.supervised_step(batch["img_x"], batch["mask_x"])
for _ in range(2): step()
VIEWS = ("view", "reshape", "permute", "transpose", "t.", "slice", "select", "expand", "detach", "as_strided", "unsqueeze", "squeeze", "alias",
         "_unsafe_view", "empty", "unbind", "split", "narrow", "size", "stride", "is_", "_local_scalar", "lift_fresh", "unfold", "chunk", "numel")
sites = collections.Counter()
class Mode(TorchDispatchMode):
    def __torch_dispatch__(self, func, types_, args=(), kwargs=None):
        name = str(func)
        if not any(("aten." + v) in name for v in VIEWS):
            flat = [a for a in list(args) + list((kwargs or {}).values()) if isinstance(a, torch.Tensor)]
            if any(a.is_cuda for a in flat) or "zeros" in name or "full" in name or "arange" in name:
                fr = [f for f in traceback.extract_stack() if "semivl_b200/" in f.filename]
                where = f"{fr[-1].filename.split('semivl_b200/')[-1]}:{fr[-1].lineno}" if fr else "?"
                sites[(where, name)] += 1
        return func(*args, **(kwargs or {}))
with Mode():
    step()
torch.cuda.synchronize()
print(f"{sum(sites.values())} kernel-launching aten calls in one step")
for (w, n), k in sorted(sites.items(), key=lambda x: (-x[1], x[0])):
    print(f"{k:4d}  {w:32s} {n}")
