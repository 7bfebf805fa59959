"""Llama patch maps (reference lxt/efficient/models/llama.py:9-21) + the nn.Linear GEMM patch."""
from functools import partial

from torch.nn import Dropout, Linear
from transformers.models.llama import modeling_llama
from transformers.models.llama.modeling_llama import LlamaMLP, LlamaRMSNorm

from ..patches import (cp_gated_mlp_forward, dropout_forward, gated_mlp_forward, linear_forward, patch_attention,
                       patch_cp_attention, patch_method, rms_norm_forward)

attnLRP = {
    LlamaMLP: partial(patch_method, gated_mlp_forward),
    LlamaRMSNorm: partial(patch_method, rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_llama: patch_attention,
}

cp_LRP = {
    LlamaMLP: partial(patch_method, cp_gated_mlp_forward),
    LlamaRMSNorm: partial(patch_method, rms_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    Linear: partial(patch_method, linear_forward, keep_original=True),
    modeling_llama: patch_cp_attention,
}
