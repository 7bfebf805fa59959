"""Qwen2: same rule set as Llama (reference lxt/efficient/models/qwen2.py)."""
from ._families import gated_decoder_maps

modeling_qwen2, attnLRP, cp_LRP = gated_decoder_maps("qwen2", "Qwen2MLP", "Qwen2RMSNorm")
