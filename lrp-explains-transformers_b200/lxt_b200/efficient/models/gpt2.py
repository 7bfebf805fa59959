"""GPT-2 patch maps (reference lxt/efficient/models/gpt2.py:11-32): plain MLP + LayerNorm."""
from functools import partial

from torch.nn import Dropout, LayerNorm
from transformers.models.gpt2 import modeling_gpt2
from transformers.models.gpt2.modeling_gpt2 import GPT2MLP

from ..patches import (dropout_forward, layer_norm_forward, mlp_forward, patch_attention, patch_cp_attention, patch_method)

attnLRP = {
    GPT2MLP: partial(patch_method, mlp_forward),
    LayerNorm: partial(patch_method, layer_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    modeling_gpt2: patch_attention,
}

cp_LRP = {
    GPT2MLP: partial(patch_method, mlp_forward),
    LayerNorm: partial(patch_method, layer_norm_forward),
    Dropout: partial(patch_method, dropout_forward),
    modeling_gpt2: patch_cp_attention,
}
