"""Patch-map factories.  A patch map is `{target: callable(target) -> bool}` (the reference's convention,
lxt/efficient/core.py:36-44); the reference spells one dict per model family by hand, here the families that share a rule
set are generated from their class names."""
import importlib
from functools import partial

from torch.nn import Dropout, LayerNorm, Linear

from .. import patches as P


def _resolve(family: str, *class_names: str):
    mod = importlib.import_module(f"transformers.models.{family}.modeling_{family}")
    return (mod,) + tuple(getattr(mod, n) for n in class_names)


def gated_decoder_maps(family: str, mlp_cls: str, norm_cls: str, norm_forward=P.rms_norm_forward):
    """Decoder families built from RMSNorm + gated MLP + registry attention (Llama, Qwen2/3, Gemma-3):
    returns (modeling module, AttnLRP map, CP-LRP map).  `nn.Linear` additionally goes to the tcgen05 GEMM."""
    modeling, mlp, norm = _resolve(family, mlp_cls, norm_cls)

    def build(mlp_forward, attention_patcher):
        return {
            mlp: partial(P.patch_method, mlp_forward),
            norm: partial(P.patch_method, norm_forward),
            Dropout: partial(P.patch_method, P.dropout_forward),
            Linear: partial(P.patch_method, P.linear_forward, keep_original=True),
            modeling: attention_patcher,
        }

    return modeling, build(P.gated_mlp_forward, P.patch_attention), build(P.cp_gated_mlp_forward, P.patch_cp_attention)


def plain_decoder_maps(family: str, mlp_cls: str):
    """LayerNorm + plain (non-gated) MLP families (GPT-2): identity rule on the activation and on LayerNorm."""
    modeling, mlp = _resolve(family, mlp_cls)

    def build(attention_patcher):
        return {
            mlp: partial(P.patch_method, P.mlp_forward),
            LayerNorm: partial(P.patch_method, P.layer_norm_forward),
            Dropout: partial(P.patch_method, P.dropout_forward),
            modeling: attention_patcher,
        }

    return modeling, build(P.patch_attention), build(P.patch_cp_attention)
