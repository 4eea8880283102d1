from .builder import build_model, forward_wrapper  # noqa: F401
from .maskclip_vit import MaskClipVisionTransformer  # noqa: F401
from .resnet import ResNetV1c  # noqa: F401
from .vlg_head import VLGHead  # noqa: F401
from .vlm import VLM  # noqa: F401
