"""Qwen3 patch maps (reference lxt/efficient/models/qwen3.py) — same rule set as Llama."""
from functools import partial

from torch.nn import Dropout, Linear
from transformers.models.qwen3 import modeling_qwen3
from transformers.models.qwen3.modeling_qwen3 import Qwen3MLP, Qwen3RMSNorm

from ..patches import (cp_gated_mlp_forward, dropout_forward, gated_mlp_forward, linear_forward, patch_attention,
                       patch_cp_attention, patch_method, rms_norm_forward)

attnLRP = {
    Qwen3MLP: partial(patch_method, gated_mlp_forward),
    Qwen3RMSNorm: partial(patch_method, rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_qwen3: patch_attention,
}

cp_LRP = {
    Qwen3MLP: partial(patch_method, cp_gated_mlp_forward),
    Qwen3RMSNorm: partial(patch_method, rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_qwen3: patch_cp_attention,
}
