"""`conservation_check()` — drop-in for lxt/explicit/check.py:6-15: inside the context every explicit rule redistributes
relevance uniformly, so `x.grad.sum()` must equal the seeded relevance if every op of the model is wrapped by a rule."""
from contextlib import contextmanager

from .functional import CONSERVATION_CHECK_FLAG


@contextmanager
def conservation_check():
    CONSERVATION_CHECK_FLAG[0] = True
    try:
        yield
    finally:
        CONSERVATION_CHECK_FLAG[0] = False
