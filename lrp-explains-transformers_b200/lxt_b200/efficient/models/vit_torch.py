"""torchvision ViT patch map (reference lxt/efficient/models/vit_torch.py:7-11): CP-LRP — identity rule on GELU
and LayerNorm, q/k detached in nn.MultiheadAttention."""
from functools import partial

from torch.nn import GELU, LayerNorm, MultiheadAttention
from torchvision.models import vision_transformer

from ..patches import cp_multi_head_attention_forward, layer_norm_forward, non_linear_forward, patch_method

cp_LRP = {
    GELU: partial(patch_method, non_linear_forward, keep_original=True),
    LayerNorm: partial(patch_method, layer_norm_forward),
    MultiheadAttention: partial(patch_method, cp_multi_head_attention_forward, keep_original=True),
}
