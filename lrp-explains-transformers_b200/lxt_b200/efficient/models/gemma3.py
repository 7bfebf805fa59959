"""Gemma-3: `(1 + w)` RMSNorm with the identity rule, GELU-tanh gated MLP, sliding-window + global attention
(reference lxt/efficient/models/gemma3.py; the reference patches `Gemma3RMSNorm._norm`, here the whole forward runs in
one kernel with the same semantics)."""
from ..patches import gemma_rms_norm_forward
from ._families import gated_decoder_maps

modeling_gemma3, attnLRP, cp_LRP = gated_decoder_maps("gemma3", "Gemma3MLP", "Gemma3RMSNorm", gemma_rms_norm_forward)
