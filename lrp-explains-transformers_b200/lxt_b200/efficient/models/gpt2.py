"""GPT-2: LayerNorm + plain GELU MLP (reference lxt/efficient/models/gpt2.py)."""
from ._families import plain_decoder_maps

modeling_gpt2, attnLRP, cp_LRP = plain_decoder_maps("gpt2", "GPT2MLP")
