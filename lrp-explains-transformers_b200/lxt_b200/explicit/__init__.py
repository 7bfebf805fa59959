"""Drop-in for `lxt.explicit` (functional + rules + conservation check; the fx Composite tooling of the reference is out of scope)."""
from . import functional, rules, check  # noqa: F401
