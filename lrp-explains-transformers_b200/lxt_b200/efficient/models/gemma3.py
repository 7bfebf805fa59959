"""Gemma-3 patch maps (reference lxt/efficient/models/gemma3.py:11-25): `(1+w)` RMSNorm with the identity rule."""
from functools import partial

from torch.nn import Dropout, Linear
from transformers.models.gemma3 import modeling_gemma3
from transformers.models.gemma3.modeling_gemma3 import Gemma3MLP, Gemma3RMSNorm

from ..patches import (cp_gated_mlp_forward, dropout_forward, gated_mlp_forward, gemma_rms_norm_forward, linear_forward,
                       patch_attention, patch_cp_attention, patch_method)

attnLRP = {
    Gemma3MLP: partial(patch_method, gated_mlp_forward),
    Gemma3RMSNorm: partial(patch_method, gemma_rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_gemma3: patch_attention,
}

cp_LRP = {
    Gemma3MLP: partial(patch_method, cp_gated_mlp_forward),
    Gemma3RMSNorm: partial(patch_method, gemma_rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_gemma3: patch_cp_attention,
}
