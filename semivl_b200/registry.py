"""mmseg-style registries and the python-file Config used by the reference's builder (model/builder.py:110-151).

The reference takes these from mmcv 1.4.4 / mmsegmentation 0.24.0 (un-vendored); only the behaviour the hot
path relies on is provided: `@REG.register_module()`, `REG.build(cfg)` (pops 'type', calls cls(**rest)),
`Config.fromfile(path)` (exec a python file into an attribute-accessible dict) and `build_segmentor/backbone/head`.
"""
import copy
import os


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, **default_args):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"{self.name}.build expects a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        typ = args.pop("type")
        cls = typ if isinstance(typ, type) else self.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        return cls(**args)

    def __contains__(self, key):
        return key in self._modules


BACKBONES = Registry("backbone")
HEADS = Registry("head")
SEGMENTORS = Registry("segmentor")
LOSSES = Registry("loss")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    extra = {}
    if train_cfg is not None:
        extra["train_cfg"] = train_cfg
    if test_cfg is not None:
        extra["test_cfg"] = test_cfg
    return SEGMENTORS.build(cfg, **extra)


class ConfigDict(dict):
    """dict with attribute access (nested dicts are wrapped on the way in)."""

    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, list):
            return [ConfigDict._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        scope = {}
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), scope)
        return Config({k: v for k, v in scope.items() if not k.startswith("_") and not callable(v) and not isinstance(v, type(os))})
