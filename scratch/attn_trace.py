"""Who waits for whom in the tcgen05 attention forward (diag build): cycle totals of CTA (1, 0, 0) over its 17 key steps."""
import os, sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
B, Lq, H = 16, 1025, 12
qkv = (torch.randn(B * Lq, 3 * 768, device="cuda") * 0.5).bfloat16()
for _ in range(3): ops.attention_fwd(qkv, B, Lq, H, False)
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["SVL_ATTN_TRACE"] = str(tr.data_ptr())
ops.attention_fwd(qkv, B, Lq, H, False); torch.cuda.synchronize()
os.environ.pop("SVL_ATTN_TRACE")
t = tr.cpu().tolist()
nt = (Lq + 63) // 64
print(f"MMA warp: {t[0]} cycles for {nt} key steps ({t[0] / nt:.0f} per step); waiting for K {t[1]}, V {t[2]}, P (softmax) {t[3]}, O done (P V complete) {t[4]}; "
      f"issuing / other {t[0] - t[1] - t[2] - t[3] - t[4]}")
print(f"softmax warp 0: waiting for S {t[5]} ({t[5] / nt:.0f} per step)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention_fwd(qkv, B, Lq, H, False)
e1.record(); torch.cuda.synchronize()
print(f"attention forward {e0.elapsed_time(e1) * 100:.1f} us")

att, lse = ops.attention_fwd(qkv, B, Lq, H, False)
datt = torch.randn(B * Lq, 768, device="cuda").bfloat16()
for _ in range(3): ops.attention_bwd(qkv, att, datt, lse, B, Lq, H, False)
tr.zero_()
os.environ["SVL_ATTN_TRACE"] = str(tr.data_ptr())
ops.attention_bwd(qkv, att, datt, lse, B, Lq, H, False); torch.cuda.synchronize()
os.environ.pop("SVL_ATTN_TRACE")
t = tr.cpu().tolist()
nq = (Lq + 63) // 64
print(f"backward kv kernel, MMA warp: {t[8]} cycles for {nq} query steps ({t[8] / nq:.0f} per step); waiting for Q / dO {t[9]}, P (softmax) {t[10]}, "
      f"accumulates of the previous step {t[11]}; issuing / other {t[8] - t[9] - t[10] - t[11]}")
print(f"softmax warp 2: waiting for S {t[12]} ({t[12] / nq:.0f} per step)")
e0.record()
for _ in range(10): ops.attention_bwd(qkv, att, datt, lse, B, Lq, H, False)
e1.record(); torch.cuda.synchronize()
print(f"attention backward {e0.elapsed_time(e1) * 100:.1f} us")
