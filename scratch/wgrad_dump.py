import os, sys, torch
sys.path.insert(0, ".")
from semivl_b200 import lib as L, ops
L.check_device()
nb, h, w, cin, cout = 1, 4, 64, 32, 32
# dy[n, y, x, c] = y * 1000 + x * 4 + c / 8 (distinct, exactly representable in bf16? use small ints): encode y in hundreds, x, and c separately
dy = torch.zeros(nb, h, w, cout); x = torch.zeros(nb, h, w, cin)
for y in range(h):
    for xx in range(w):
        dy[0, y, xx, :] = (y + 1) * 64 + xx       # <= 5*64+63 = 383: needs 9 bits: not exact in bf16 (8 bits) -> use y*64 only for channel 0..: store y in even channels, x in odd
dy = torch.zeros(nb, h, w, cout); 
ys = torch.arange(h).view(1, h, 1, 1).float(); xs = torch.arange(w).view(1, 1, w, 1).float(); cs = torch.arange(cout).view(1, 1, 1, cout).float()
dy = torch.where(cs % 4 == 0, ys + 1, torch.where(cs % 4 == 1, xs, torch.where(cs % 4 == 2, cs, -(ys + 1)))).expand(nb, h, w, cout).contiguous()
xt = dy.clone()
dya, xa = dy.cuda().bfloat16(), xt.cuda().bfloat16()
dump = torch.zeros(16384 + 9216, dtype=torch.uint8, device="cuda")
os.environ["SVL_WGRAD_DUMP"] = str(dump.data_ptr())
filt = [(i - 1, j - 1) for i in range(3) for j in range(3)]
dw = torch.zeros(9, cout, cin, device="cuda")
ops.wgrad(dya, xa, dw, m=cout, n=cin, conv=(nb, h, w), filt=filt)
torch.cuda.synchronize()
raw = dump.cpu().view(torch.bfloat16).float().view(-1, 64)          # 128-byte lines of 64 bf16
def unsw(line_idx, line):       # undo SWIZZLE_128B: 16-byte chunk c stored at c ^ (line & 7)
    out = torch.empty_like(line)
    for c in range(8):
        out[c * 8:(c + 1) * 8] = line[((c ^ (line_idx & 7)) * 8):((c ^ (line_idx & 7)) * 8 + 8)]
    return out
print("A box 0 (expected per pixel line: [row r-1 (=-1: zeros) 32 ch | row r (=0) 32 ch]); first step r = 0, cx = 0")
for li in (0, 1, 2, 9, 63):
    l = unsw(li, raw[li])
    print(li, "first half ch0..7:", l[:8].tolist(), " second half ch0..7:", l[32:40].tolist())
print("A box 1 (rows 1 | 2)")
for li in (0, 1, 63):
    l = unsw(li, raw[64 + li])
    print(li, "first half:", l[:8].tolist(), " second half:", l[32:40].tolist())
print("B strip (rows 0 | 1, pixels -1..64)")
for li in (0, 1, 2, 65):
    l = unsw(li, raw[128 + li])
    print(li, "first half:", l[:8].tolist(), " second half:", l[32:40].tolist())
