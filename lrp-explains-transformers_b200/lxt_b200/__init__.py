"""lxt_b200 — B200-native AttnLRP hot path behind the `lxt` rule / patch API.

Sub-modules
  _capi     ctypes binding of liblrp_b200.so (C ABI in include/lrp_b200.h)
  ops       torch-tensor wrappers over the C ABI (device memory + streams only)
  engine    layer-by-layer AttnLRP executor for Llama-family decoders (headline path)
  efficient drop-in mirror of `lxt.efficient` (monkey_patch, rules, patches, model maps)
  explicit  drop-in mirror of `lxt.explicit.functional` / `lxt.explicit.rules`
  dist      one-process-per-GPU batch sharding with a single NCCL gather
"""
__version__ = "0.1.0"

from . import _capi  # noqa: F401


def install_as_lxt() -> None:
    """Register this package under the name `lxt` so that user scripts written against the reference
    (`from lxt.efficient import monkey_patch`) run unchanged."""
    import sys
    from . import efficient, explicit

    import importlib

    sys.modules.setdefault("lxt", sys.modules[__name__])
    sys.modules.setdefault("lxt.efficient", efficient)
    sys.modules.setdefault("lxt.explicit", explicit)
    # every sub-module under its reference name too: `import lxt.explicit.functional as lf` must hand out THIS module object, not
    # a second copy loaded from the package path (module state such as the conservation-check flag would fork)
    for pkg, names in (("explicit", ("functional", "rules", "check", "special", "modules")),
                       ("efficient", ("core", "rules", "patches", "models", "zennit_rules"))):
        for n in names:
            sys.modules.setdefault(f"lxt.{pkg}.{n}", importlib.import_module(f"{__name__}.{pkg}.{n}"))
    sys.modules.setdefault("lxt.efficient.zennit_patches", sys.modules[f"{__name__}.efficient.zennit_rules"])
    try:
        import zennit  # noqa: F401
    except ImportError:
        # the reference's ViT recipe imports `zennit.rules.Gamma` and `zennit.composites.LayerMapComposite` (examples/vit_torch.py:
        # 7-8): stand-ins backed by efficient/zennit_rules.py when the real package is absent
        import types
        from .efficient import zennit_rules
        z, zr, zc = types.ModuleType("zennit"), types.ModuleType("zennit.rules"), types.ModuleType("zennit.composites")
        zr.Gamma = zennit_rules.Gamma
        zc.LayerMapComposite = zennit_rules.LayerMapComposite
        z.rules, z.composites = zr, zc
        sys.modules.setdefault("zennit", z)
        sys.modules.setdefault("zennit.rules", zr)
        sys.modules.setdefault("zennit.composites", zc)
