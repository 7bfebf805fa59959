"""Drop-in for `lxt.explicit` (functional + rules + modules + special + conservation check; the fx Composite tooling of the
reference, lxt/explicit/core.py, needs `transformers.utils.fx` and is out of scope)."""
from . import functional, rules, check, special, modules  # noqa: F401
