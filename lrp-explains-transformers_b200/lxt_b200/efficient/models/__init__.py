"""Model -> patch-map registry (reference lxt/efficient/models/__init__.py:29-51)."""
import importlib
import warnings

DEFAULT_MAP = {}


def _register(name, module_attr, map_attr):
    try:
        mod = importlib.import_module(f"{__name__}.{name}")
    except Exception as e:  # an optional third-party model family is not installed
        warnings.warn(f"lxt_b200.efficient.models.{name} disabled: {e}")
        return None
    DEFAULT_MAP[getattr(mod, module_attr)] = getattr(mod, map_attr)
    return mod


llama = _register("llama", "modeling_llama", "attnLRP")
qwen2 = _register("qwen2", "modeling_qwen2", "attnLRP")
qwen3 = _register("qwen3", "modeling_qwen3", "attnLRP")
gemma3 = _register("gemma3", "modeling_gemma3", "attnLRP")
gpt2 = _register("gpt2", "modeling_gpt2", "attnLRP")
vit_torch = _register("vit_torch", "vision_transformer", "cp_LRP")


def get_default_map(module):
    """default patch map of a supported modelling module; ValueError listing the supported ones otherwise"""
    if module in DEFAULT_MAP:
        return DEFAULT_MAP[module]
    supported_models = ", ".join(key.__name__ for key in DEFAULT_MAP)
    raise ValueError(
        f"{module.__name__} not yet supported. Supported models are: {supported_models} "
        "Please provide a custom patch_map."
    )
