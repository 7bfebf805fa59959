"""Drop-in for `lxt.explicit` (functional + rules; the fx Composite tooling of the reference is out of scope)."""
from . import functional, rules  # noqa: F401
