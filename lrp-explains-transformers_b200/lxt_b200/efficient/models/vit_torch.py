"""torchvision ViT patch map (reference lxt/efficient/models/vit_torch.py:7-11): CP-LRP — identity rule on GELU
and LayerNorm, q/k detached in nn.MultiheadAttention.  On CUDA bf16 the whole block (in/out projections, attention,
MLP Linears) runs on the B200 kernels; other inputs take the reference-equivalent route."""
from functools import partial

from torch.nn import GELU, LayerNorm, Linear, MultiheadAttention
from torchvision.models import vision_transformer

from ..patches import (b200_cp_multi_head_attention_forward, layer_norm_forward, linear_forward, non_linear_forward,
                       patch_method)

cp_LRP = {
    GELU: partial(patch_method, non_linear_forward, keep_original=True),
    LayerNorm: partial(patch_method, layer_norm_forward),
    MultiheadAttention: partial(patch_method, b200_cp_multi_head_attention_forward, keep_original=True),
    Linear: partial(patch_method, linear_forward, keep_original=True),
}
