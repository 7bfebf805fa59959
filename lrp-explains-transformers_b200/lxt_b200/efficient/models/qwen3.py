"""Qwen3: same rule set as Llama; its per-head q/k RMSNorm is the same class (reference lxt/efficient/models/qwen3.py)."""
from ._families import gated_decoder_maps

modeling_qwen3, attnLRP, cp_LRP = gated_decoder_maps("qwen3", "Qwen3MLP", "Qwen3RMSNorm")
