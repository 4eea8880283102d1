"""`ResNetV1c` conv encoder of the Cityscapes skr04 model under its mmseg registry name and constructor keywords
(configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:50-60; built by model/vlm.py:50-52) over the B200 conv-encoder engine.

mmsegmentation 0.24.0 (where ResNetV1c lives) is not vendored in the reference; the module restates the one configuration the reference
uses -- depth 101, num_stages=1, out_indices=[0], strides=[1], dilations=[1], style='pytorch', deep stem -- and keeps mmseg's state-dict
names (`stem.0.weight`, `stem.1.running_mean`, `layer1.0.downsample.1.weight`, ...), i.e. the keys of
`pretrained/resnet101_v1c-e67eebb6.pth`.  The torch modules are parameter / buffer containers; the arithmetic runs in
semivl_b200.engine.convenc.  SyncBN: with torch.distributed initialised, the batch statistics of every BatchNorm (and their gradients'
sums) are all-reduced over the ranks, as torch.nn.SyncBatchNorm does after `convert_sync_batchnorm` (semivl.py:136).
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from ..engine.convenc import STEM, ConvEncEngine, no_sync
from ..registry import BACKBONES


def sync_sums(t, count):
    """SyncBN hook: SUM all-reduce of the [2, C] statistics and of the row count over the ranks (NCCL on GPUs, gloo in the CPU tests);
    returns the global row count.  One process: identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return count
    buf = torch.cat((t.reshape(-1), torch.full((1,), float(count), device=t.device, dtype=t.dtype)))     # a fill kernel: capturable in a CUDA graph
    dist.all_reduce(buf)
    t.copy_(buf[:-1].view_as(t))
    return float(buf[-1].item()) if t.device.type == "cpu" else _count_of(buf, count)


def _count_of(buf, local_count):
    # every rank runs the same per-GPU batch in this path (DistributedSampler, drop_last), so the global row count is known without a
    # device -> host read-back in the middle of the step
    return float(local_count) * dist.get_world_size()


class _Bottleneck(nn.Module):
    def __init__(self, cin, planes, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))


class _ConvEncFunction(torch.autograd.Function):
    """Whole-encoder autograd node: forward and backward are the engine's hand-scheduled kernel sequences."""

    @staticmethod
    def forward(ctx, module, img, grad_mode, names, *params):
        p = {k: v.detach() for k, v in zip(names, params)}
        p.update({k: v for k, v in module.named_buffers() if "num_batches_tracked" not in k})
        need_grad = grad_mode and module.training and any(t.requires_grad for t in params)
        feat, ectx = module.engine.forward(img, p, training=module.training, need_grad=need_grad, sync=module.sync)
        ctx.module, ctx.names, ctx.ectx = module, names, ectx
        ctx.req = [t.requires_grad for t in params]
        ctx.save_for_backward(*params)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        if ctx.ectx is None:
            raise RuntimeError("ResNetV1c backward needs a training-mode forward under grad mode")
        params = ctx.saved_tensors
        p = {k: v.detach() for k, v in zip(ctx.names, params)}
        grads = {k: torch.zeros_like(v) for k, v in zip(ctx.names, params)}
        ctx.module.engine.backward(ctx.ectx, dfeat.contiguous(), p, grads, sync=ctx.module.sync)
        ctx.ectx = None
        return (None, None, None, None) + tuple(grads[k] if r else None for k, r in zip(ctx.names, ctx.req))


@BACKBONES.register_module()
class ResNetV1c(nn.Module):
    def __init__(self, depth=101, num_stages=1, out_indices=(0,), dilations=(1,), strides=(1,), norm_cfg=None, style='pytorch',
                 contract_dilation=True, pretrained=None, init_cfg=None, norm_eval=False, precise=False, **unsupported):
        super().__init__()
        if (depth, num_stages, tuple(out_indices), tuple(dilations), tuple(strides), style) != (101, 1, (0,), (1,), (1,), 'pytorch') or unsupported:
            raise NotImplementedError("semivl_b200 implements the conv encoder of the skr04 config: ResNetV1c(depth=101, num_stages=1, "
                                      f"out_indices=[0], strides=[1], dilations=[1], style='pytorch'); got extra / different arguments {unsupported}")
        layers = []
        for _, _, cin, cout, stride in STEM:
            layers += [nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        self.stem = nn.Sequential(*layers)
        self.layer1 = nn.Sequential(_Bottleneck(64, 64, True), _Bottleneck(256, 64, False), _Bottleneck(256, 64, False))
        self.pretrained, self.norm_eval = pretrained, norm_eval
        self.sync_bn = bool(norm_cfg and norm_cfg.get('type') == 'SyncBN')
        self.engine = ConvEncEngine(precise=precise)

    @property
    def sync(self):
        return sync_sums if self.sync_bn else no_sync

    def set_precise(self, precise):
        self.engine.precise = bool(precise)

    def init_weights(self):
        """mmseg ResNet.init_weights: `pretrained` checkpoint by key name, else kaiming-normal convolutions and unit BatchNorms."""
        if isinstance(self.pretrained, str):
            sd = torch.load(self.pretrained, map_location='cpu')
            sd = sd.get('state_dict', sd)
            own = self.state_dict()
            self.load_report = self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)   # deeper stages are not built
            return
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def forward(self, x):
        """x [B,3,H,W] -> (feature [B,256,H/4,W/4],)   (out_indices=[0]: the tuple VLGHead receives as inputs[2])"""
        names = tuple(n for n, _ in self.named_parameters())
        params = tuple(p for _, p in self.named_parameters())
        out = _ConvEncFunction.apply(self, x, torch.is_grad_enabled(), names, *params)
        if self.training:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d) and m.num_batches_tracked is not None:
                    m.num_batches_tracked += 1
        return (out.permute(0, 3, 1, 2),)
