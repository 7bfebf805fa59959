"""Drop-in for `lxt.efficient`: `from lxt_b200.efficient import monkey_patch`."""
from .core import monkey_patch  # noqa: F401
from . import rules, patches, models  # noqa: F401
from .zennit_rules import monkey_patch_zennit  # noqa: F401,E402
