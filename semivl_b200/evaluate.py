"""Evaluation of the reference (third_party/unimatch/supervised.py:40-164): `predict` (original / center_crop /
padded_sliding_window / zegclip_sliding_window / sliding_window) and `evaluate` (mIoU with a cross-rank reduction of the
intersection / union / target counts).  Same call signatures and `cfg` keys ('crop_size', 'stride', 'nclass').

The window stitching, the arg-max and the integer histograms are svl_* kernels (csrc/eval.cu); the per-window forward is the
model's own kernel path.  Nothing is staged through the host: the reference moves every prediction to numpy for
`intersectionAndUnion` and back to the GPU for the all-reduce (supervised.py:152-160); here the int64 counts stay on the device
and ONE all-reduce of the [3, K] count matrix per batch replaces the three."""
import numpy as np
import torch
import torch.distributed as dist

from . import lib as L

MODES = ('original', 'center_crop', 'padded_sliding_window', 'zegclip_sliding_window', 'sliding_window')


def _accumulate(final, pred, count, y1, x1, ch, cw, softmax):
    B, N, H, W = final.shape
    h, w = pred.shape[-2:]
    L.call("svl_window_accumulate", final, pred.contiguous(), count, B, N, H, W, h, w, y1, x1, 0, 0, ch, cw, 1 if softmax else 0)


def argmax_classes(x):
    B, N, H, W = x.shape
    out = torch.empty(B, H, W, device=x.device, dtype=torch.int64)
    L.call("svl_argmax_classes", x.contiguous(), out, B, N, H * W)
    return out


def predict(model, img, mask, mode, cfg, return_logits=False):
    """supervised.py:40-130.  img [b,3,h,w] on the GPU; returns int64 labels [b,h,w] (and the stitched scores)."""
    assert mode in MODES, mode
    n = cfg['nclass']
    b, _, h, w = img.shape
    if mode == 'padded_sliding_window':                       # supervised.py:41-65: zero-padded windows, softmax scores summed
        grid, stride = cfg['crop_size'], cfg['stride']
        if stride < 1:
            stride = int(grid * stride)
        final = torch.zeros(b, n, h, w, device=img.device)
        row = 0
        while row < h:
            col = 0
            while col < w:
                y2, x2 = min(h, row + grid), min(w, col + grid)
                ch, cw = y2 - row, x2 - col
                crop = torch.zeros(b, 3, grid, grid, device=img.device)
                crop[:, :, :ch, :cw] = img[:, :, row:y2, col:x2]
                _accumulate(final, model(crop).float(), None, row, col, ch, cw, True)
                col += stride
            row += stride
    elif mode == 'zegclip_sliding_window':                    # supervised.py:67-103: full windows shifted inwards, logits averaged
        hs = ws = cfg['stride']
        hc = wc = cfg['crop_size']
        hg = max(h - hc + hs - 1, 0) // hs + 1
        wg = max(w - wc + ws - 1, 0) // ws + 1
        final = torch.zeros(b, n, h, w, device=img.device)
        count = torch.zeros(b, h, w, device=img.device)
        for hi in range(hg):
            for wi in range(wg):
                y2, x2 = min(hi * hs + hc, h), min(wi * ws + wc, w)
                y1, x1 = max(y2 - hc, 0), max(x2 - wc, 0)
                _accumulate(final, model(img[:, :, y1:y2, x1:x2].contiguous()).float(), count, y1, x1, y2 - y1, x2 - x1, False)
        assert int((count == 0).sum()) == 0
        L.call("svl_divide_count", final, count, b, n, h * w)
        if tuple(mask.shape[-2:]) != (h, w):
            # supervised.py:95-100: the averaged logits are resized to the label size with align_corners=True
            H2, W2 = int(mask.shape[-2]), int(mask.shape[-1])
            resized = torch.empty(b, n, H2, W2, device=img.device)
            L.call("svl_resize_bilinear_ac", final, resized, b * n, h, w, H2, W2)
            final = resized
    elif mode == 'sliding_window':                            # supervised.py:105-117: clipped windows, stride 2/3 grid, softmax summed
        grid = cfg['crop_size']
        final = torch.zeros(b, n, h, w, device=img.device)
        step = int(grid * 2 / 3)
        row = 0
        while row < h:
            col = 0
            while col < w:
                y2, x2 = min(h, row + grid), min(w, col + grid)
                _accumulate(final, model(img[:, :, row:y2, col:x2].contiguous()).float(), None, row, col, y2 - row, x2 - col, True)
                col += step
            row += step
    else:
        if mode == 'center_crop':                             # supervised.py:120-124
            c = cfg['crop_size']
            sh, sw = (h - c) // 2, (w - c) // 2
            img = img[:, :, sh:sh + c, sw:sw + c].contiguous()
        final = model(img).float()
    pred = argmax_classes(final)
    return (pred, final) if return_logits else pred


def crop_mask_for(mode, mask, cfg):
    """the label crop of the 'center_crop' mode (supervised.py:124; the reference crops a local copy that it never uses again)"""
    if mode != 'center_crop':
        return mask
    c = cfg['crop_size']
    h, w = mask.shape[-2:]
    sh, sw = (h - c) // 2, (w - c) // 2
    return mask[:, sh:sh + c, sw:sw + c]


def intersection_union_counts(pred, target, nclass, ignore_index=255, counts=None):
    """int64 [3, K] += (intersection, prediction area, target area) -- third_party/unimatch/util/utils.py:91-103."""
    if counts is None:
        counts = torch.zeros(3, nclass, device=pred.device, dtype=torch.int64)
    assert pred.shape == target.shape, f'{tuple(pred.shape)} != {tuple(target.shape)}'
    L.call("svl_intersection_union", pred.contiguous(), target.contiguous(), pred.numel(), nclass, ignore_index, counts)
    return counts


def reduce_counts(counts):
    """sum over ranks (supervised.py:158-160: three all-reduces of intersection / union / target; here one of the [3, K] matrix)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts)
    return counts


def evaluate(model, loader, mode, cfg):
    """supervised.py:133-164.  Returns (mIoU in percent, per-class IoU in percent as a numpy array)."""
    model.eval()
    assert mode in MODES
    total = None
    with torch.no_grad():
        for img, mask, _ in loader:
            img = img.cuda(non_blocking=True)
            pred = predict(model, img, mask, mode, cfg)
            target = mask.cuda(non_blocking=True).to(torch.int64)
            if pred.shape != target.shape:
                target = crop_mask_for(mode, target, cfg)
            batch = reduce_counts(intersection_union_counts(pred, target, cfg['nclass'], 255))
            total = batch if total is None else total + batch
    inter, area_pred, area_tgt = (total[i].double().cpu().numpy() for i in range(3))
    union = area_pred + area_tgt - inter
    iou_class = inter / (union + 1e-10) * 100.0
    return float(np.mean(iou_class)), iou_class
