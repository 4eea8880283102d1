"""CPU restatement of the conv encoder of the Cityscapes skr04 model: mmseg's `ResNetV1c(depth=101, num_stages=1, out_indices=[0],
strides=[1], dilations=[1], style='pytorch', norm_cfg=SyncBN)` (configs/_base_/models/vlm-vlg-aspp-s2p4-skr04-ftap-mcvitb.py:50-60) --
TEST INFRASTRUCTURE, not a product path: the checker of semivl_b200/engine/convenc.py.

PARITY UNPINNED for the mmseg-specific part: mmsegmentation 0.24.0 is not vendored under /root/reference and is not installed, so the deep
stem (three 3x3 convs 3->32 (stride 2) ->32->64, each + BN + ReLU, then a 3x3/stride-2 max-pool; V1c = deep_stem=True, avg_down=False) is
restated from the published architecture.  The residual stage IS pinned: `layer1` (3 Bottlenecks, 64 -> 64 -> 256, stride 1, style
'pytorch', 1x1 conv + BN shortcut on the first block) is checked against torchvision's ResNet-101 `layer1` with the same weights
(tests/test_host_cpu.py).  Parameter names follow mmseg's state dict (`stem.0.weight`, `stem.1.*`, ..., `layer1.0.conv1.weight`,
`layer1.0.bn1.*`, `layer1.0.downsample.0.weight`, `layer1.0.downsample.1.*`), which is what `pretrained/resnet101_v1c-e67eebb6.pth` holds.

SyncBN in training mode normalises with the statistics of the GLOBAL batch (all ranks); with one process that is BatchNorm over (B, H, W).
"""
import torch
import torch.nn.functional as F

STEM = ((0, 1, 3, 32, 2), (3, 4, 32, 32, 1), (6, 7, 32, 64, 1))        # (conv index, bn index, cin, cout, stride)


def _bn(x, p, pre, training, eps=1e-5, running=None, momentum=0.1):
    """nn.BatchNorm2d / SyncBN forward: batch statistics (biased variance) in training mode, running statistics otherwise.
    `running` (optional dict): receives the updated running statistics of a training-mode call (momentum 0.1, unbiased variance), computed by
    torch's own batch_norm on clones of the current ones."""
    if training:
        if running is not None:
            rm, rv = p[pre + "running_mean"].detach().clone(), p[pre + "running_var"].detach().clone()
            out = F.batch_norm(x, rm, rv, p[pre + "weight"], p[pre + "bias"], True, momentum, eps)
            running[pre + "running_mean"], running[pre + "running_var"] = rm, rv
            return out
        return F.batch_norm(x, None, None, p[pre + "weight"], p[pre + "bias"], True, 0.0, eps)
    return F.batch_norm(x, p[pre + "running_mean"], p[pre + "running_var"], p[pre + "weight"], p[pre + "bias"], False, 0.0, eps)


def stem_forward(img, p, pre="conv_encoder.", training=True, running=None):
    x = img
    for ci, bi, _, _, stride in STEM:
        x = F.conv2d(x, p[f"{pre}stem.{ci}.weight"], stride=stride, padding=1)
        x = F.relu(_bn(x, p, f"{pre}stem.{bi}.", training, running=running))
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def bottleneck_forward(x, p, pre, training=True, running=None):
    """mmseg Bottleneck, style='pytorch', stride 1, dilation 1: 1x1 -> 3x3 -> 1x1 (x4), BN after each, ReLU after the first two and the sum."""
    out = F.relu(_bn(F.conv2d(x, p[pre + "conv1.weight"]), p, pre + "bn1.", training, running=running))
    out = F.relu(_bn(F.conv2d(out, p[pre + "conv2.weight"], padding=1), p, pre + "bn2.", training, running=running))
    out = _bn(F.conv2d(out, p[pre + "conv3.weight"]), p, pre + "bn3.", training, running=running)
    if pre + "downsample.0.weight" in p:
        x = _bn(F.conv2d(x, p[pre + "downsample.0.weight"]), p, pre + "downsample.1.", training, running=running)
    return F.relu(out + x)


def conv_encoder_forward(img, p, pre="conv_encoder.", training=True, blocks=3, running=None):
    """img [B,3,H,W] -> [feature [B,256,H/4,W/4]]  (out_indices=[0]: the list VLGHead receives as inputs[2])"""
    x = stem_forward(img, p, pre, training, running)
    for i in range(blocks):
        x = bottleneck_forward(x, p, f"{pre}layer1.{i}.", training, running)
    return [x]


def param_shapes(pre="conv_encoder.", blocks=3):
    s = {}
    for ci, bi, cin, cout, _ in STEM:
        s[f"{pre}stem.{ci}.weight"] = (cout, cin, 3, 3)
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            s[f"{pre}stem.{bi}.{leaf}"] = (cout,)
    for i in range(blocks):
        b = f"{pre}layer1.{i}."
        cin = 64 if i == 0 else 256
        for name, shape, c in (("conv1", (64, cin, 1, 1), 64), ("conv2", (64, 64, 3, 3), 64), ("conv3", (256, 64, 1, 1), 256)):
            s[b + name + ".weight"] = shape
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                s[b + "bn" + name[-1] + "." + leaf] = (c,)
        if i == 0:
            s[b + "downsample.0.weight"] = (256, 64, 1, 1)
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                s[b + "downsample.1." + leaf] = (256,)
    return s


def fixture_params(seed=0, pre=""):
    """Seeded test weights: He-scaled convolutions, BatchNorm gains around 1, non-trivial running statistics."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k, s in param_shapes(pre=pre).items():
        if k.endswith("running_var"):
            p[k] = torch.rand(s, generator=g) + 0.5
        elif k.endswith("running_mean"):
            p[k] = torch.randn(s, generator=g) * 0.1
        elif len(s) == 1 and k.endswith("weight"):
            p[k] = 1.0 + 0.2 * torch.randn(s, generator=g)
        elif len(s) == 1:
            p[k] = 0.2 * torch.randn(s, generator=g)
        else:
            fan = s[1] * s[2] * s[3]
            p[k] = torch.randn(s, generator=g) * (2.0 / fan) ** 0.5
    return p
