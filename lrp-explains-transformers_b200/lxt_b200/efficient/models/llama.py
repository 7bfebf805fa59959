"""Llama: `attnLRP` / `cp_LRP` patch maps with the reference's names (lxt/efficient/models/llama.py)."""
from ._families import gated_decoder_maps

modeling_llama, attnLRP, cp_LRP = gated_decoder_maps("llama", "LlamaMLP", "LlamaRMSNorm")
