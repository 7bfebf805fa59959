"""`monkey_patch` — drop-in for `lxt.efficient.core.monkey_patch` (reference lxt/efficient/core.py:20-43)."""
from warnings import warn

from .models import get_default_map


def monkey_patch(module, patch_map=None, verbose=False):
    """Patch the classes / functions of `module` so that a backward pass computes AttnLRP relevance on the B200
    kernels.  `patch_map` maps a target (class or python module) to a `callable(target) -> bool`; a False return
    is reported with a warning and skipped, exactly as the reference does (core.py:39-44)."""
    if patch_map is None:
        patch_map = get_default_map(module)
    for target, patch in patch_map.items():
        if patch(target):
            if verbose:
                print(f"Patched {target.__name__}")
        else:
            warn(f"Failed to patch {target.__name__}. Skipping...")
