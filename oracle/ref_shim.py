"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (semivl_b200/).

A minimal stand-in for the un-vendored third-party packages the reference
imports (mmcv-full 1.4.4, mmsegmentation 0.24.0, timm, clip, matplotlib, ...)
so that the UNMODIFIED reference files under /root/reference can be imported
in this authoring container and used to (a) generate golden vectors
(oracle/make_golden.py) and (b) validate the restated oracle
(oracle/semivl_oracle.py).  /root/reference does not exist on the GPU box, so
nothing here may be used by `-m gpu` tests, smoke() or bench.py.

Semantics restated (not copied) from the public behaviour of mmcv 1.4.4 /
mmseg 0.24.0, cross-checked against what the reference itself pins
(state-dict key names in third_party/maskclip/convert_clip_weights.py:27-64,
attribute paths in third_party/maskclip/models/backbones/maskclip_vit.py:112-115,
the corner-pad branch that triggers maskclip_vit.py:448-459).
"""
import importlib.abc
import importlib.machinery
import logging
import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("SEMIVL_REFERENCE_ROOT", "/root/reference")
_STUB_ROOTS = ("mmcv", "mmseg", "timm", "clip", "matplotlib", "tqdm", "tensorboardX", "easydict")


# --------------------------------------------------------------------------- registry
class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        for k, v in default_args.items():
            if v is not None:
                cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


BACKBONES = Registry("backbone")
HEADS = Registry("head")
SEGMENTORS = Registry("segmentor")
LOSSES = Registry("loss")
NECKS = Registry("neck")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    return SEGMENTORS.build(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


# --------------------------------------------------------------------------- mmcv.runner
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self._is_init = False

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _load_checkpoint(path, logger=None, map_location=None):
    return torch.load(path, map_location=map_location)


# --------------------------------------------------------------------------- mmcv.cnn
def build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    typ = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if typ == "LN":
        return "ln" + str(postfix), nn.LayerNorm(num_features, eps=cfg.get("eps", 1e-5))
    if typ in ("BN", "SyncBN"):
        return "bn" + str(postfix), nn.BatchNorm2d(num_features, eps=cfg.get("eps", 1e-5))
    if typ == "GN":
        return "gn" + str(postfix), nn.GroupNorm(cfg["num_groups"], num_features)
    raise KeyError(typ)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def constant_init(module, val, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    if distribution == "uniform":
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x.div(keep) * mask


def build_dropout(cfg):
    if cfg is None:
        return nn.Identity()
    cfg = dict(cfg)
    typ = cfg.pop("type")
    if typ == "DropPath":
        return _DropPath(cfg.get("drop_prob", 0.0))
    if typ == "Dropout":
        return nn.Dropout(cfg.get("drop_prob", cfg.get("p", 0.5)))
    raise KeyError(typ)


class MultiheadAttention(BaseModule):
    """identity + nn.MultiheadAttention(q,k,v)[0], batch_first by transposition."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0,
                 dropout_layer=dict(type="Dropout", drop_prob=0.0), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type="ReLU", inplace=True), ffn_drop=0.0, dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        act = {"GELU": nn.GELU, "ReLU": lambda: nn.ReLU(inplace=True)}[act_cfg["type"]]
        layers = []
        in_ch = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(in_ch, feedforward_channels), act(), nn.Dropout(ffn_drop)))
            in_ch = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


# --------------------------------------------------------------------------- mmseg
class PatchEmbed(BaseModule):
    """corner-pad to a multiple of the stride, Conv2d(k, s), flatten -> (B, L, C)."""

    def __init__(self, in_channels=3, embed_dims=768, conv_type="Conv2d", kernel_size=16, stride=None,
                 padding="corner", dilation=1, bias=True, norm_cfg=None, input_size=None, init_cfg=None):
        super().__init__(init_cfg)
        stride = stride or kernel_size
        self.kernel_size, self.stride, self.pad_mode = kernel_size, stride, padding
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size, stride=stride, bias=bias)
        self.norm = build_norm_layer(norm_cfg, embed_dims)[1] if norm_cfg is not None else None

    def forward(self, x):
        h, w = x.shape[-2:]
        k, s = self.kernel_size, self.stride
        pad_h = max((math.ceil(h / s) - 1) * s + k - h, 0)
        pad_w = max((math.ceil(w / s) - 1) * s + k - w, 0)
        if pad_h or pad_w:
            assert self.pad_mode == "corner"
            x = F.pad(x, [0, pad_w, 0, pad_h])
        x = self.projection(x)
        out_size = (x.shape[2], x.shape[3])
        x = x.flatten(2).transpose(1, 2)
        if self.norm is not None:
            x = self.norm(x)
        return x, out_size


def resize(input, size=None, scale_factor=None, mode="nearest", align_corners=None, warning=True):
    return F.interpolate(input, size, scale_factor, mode, align_corners)


class EncoderDecoder(BaseModule):
    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        if pretrained is not None:
            backbone = dict(backbone)
            backbone["pretrained"] = pretrained
        self.backbone = build_backbone(backbone)
        assert neck is None and auxiliary_head is None
        self.decode_head = build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.train_cfg, self.test_cfg = train_cfg, test_cfg


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_attr(o):
    if isinstance(o, dict):
        return _AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    if isinstance(o, tuple):
        return tuple(_to_attr(v) for v in o)
    return o


class Config:
    @staticmethod
    def fromfile(path):
        g = {}
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), g)
        return _to_attr({k: v for k, v in g.items()
                         if not k.startswith("__") and not isinstance(v, types.ModuleType)})


def get_root_logger(*a, **k):
    return logging.getLogger("mmseg")


# --------------------------------------------------------------------------- auto-stub import finder
_REAL = {
    "mmseg.models.builder": dict(BACKBONES=BACKBONES, HEADS=HEADS, SEGMENTORS=SEGMENTORS, LOSSES=LOSSES,
                                 NECKS=NECKS, build_backbone=build_backbone, build_head=build_head,
                                 build_loss=build_loss, build_segmentor=build_segmentor),
    "mmseg.models": dict(build_segmentor=build_segmentor, BACKBONES=BACKBONES, HEADS=HEADS,
                         SEGMENTORS=SEGMENTORS, LOSSES=LOSSES),
    "mmseg.models.segmentors.encoder_decoder": dict(EncoderDecoder=EncoderDecoder),
    "mmseg.models.utils": dict(PatchEmbed=PatchEmbed),
    "mmseg.ops": dict(resize=resize),
    "mmseg.utils": dict(get_root_logger=get_root_logger),
    "mmcv.runner": dict(BaseModule=BaseModule, ModuleList=ModuleList, Sequential=Sequential,
                        _load_checkpoint=_load_checkpoint),
    "mmcv.cnn": dict(build_norm_layer=build_norm_layer),
    "mmcv.cnn.bricks.transformer": dict(MultiheadAttention=MultiheadAttention, FFN=FFN),
    "mmcv.cnn.utils.weight_init": dict(trunc_normal_=trunc_normal_, constant_init=constant_init,
                                       kaiming_init=kaiming_init),
    "mmcv.utils": dict(Config=Config),
    "timm.models.layers": dict(trunc_normal_=trunc_normal_),
}


class _Dummy(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def _permissive(*a, **k):
    """Decorator / factory that swallows anything (force_fp32(...), register_model, ...)."""
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return _permissive


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        real = _REAL.get(self.__name__, {})
        if name in real:
            return real[name]
        sub = sys.modules.get(self.__name__ + "." + name)
        if sub is not None:
            return sub
        if name[:1].isupper():
            cls = type(name, (_Dummy,), {})
            setattr(self, name, cls)
            return cls
        return _permissive


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Install the stub finder and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT} (only exists in the authoring container)")
    sys.meta_path.insert(0, _Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def available():
    return os.path.isdir(REFERENCE_ROOT)


def default_cfg(dataset="pascal", nclass=21, crop_size=224, model="mmseg.vlm-vlg-aspp-s2p4-sk04-ftap-mcvitb",
                clip_encoder="mcvit16", text="single", mcc_text=None):
    return dict(model=model, nclass=nclass, crop_size=crop_size, dataset=dataset,
                text_embedding_variant=text, mcc_text=mcc_text or text, pl_text=text,
                clip_encoder=clip_encoder, disable_dropout=True, fp_rate=0.5,
                model_args=dict(pretrained=None))


def build_reference_model(cfg):
    """build_model(cfg) of the unmodified reference (model/builder.py:104-159), random init.

    `pretrained` is forced to None so init_weights takes the branch at
    maskclip_vit.py:413-429.  Must run with cwd == REFERENCE_ROOT because the
    reference uses relative config / .npy paths (builder.py:110,126-141).
    """
    install()
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        from model import builder as ref_builder  # noqa
        orig_fromfile = Config.fromfile

        def patched(path):
            c = orig_fromfile(path)
            if "model" in c and isinstance(c["model"], dict):
                c["model"]["pretrained"] = None
                if "backbone" in c["model"]:
                    c["model"]["backbone"]["pretrained"] = None
            if "backbone" in c and isinstance(c["backbone"], dict):
                c["backbone"]["pretrained"] = None
            return c
        ref_builder.Config = types.SimpleNamespace(fromfile=patched)
        model = ref_builder.build_model(cfg)
    finally:
        os.chdir(cwd)
    return model


class in_reference_cwd:
    """Context manager: the reference np.load()s its text table with a relative path on every forward (vlm.py:116)."""

    def __enter__(self):
        self._cwd = os.getcwd()
        os.chdir(REFERENCE_ROOT)

    def __exit__(self, *a):
        os.chdir(self._cwd)
