# reference model/__init__.py:1-2 star-imports its sub-modules; only the lifting MLP is in scope here.
from .bilinear import *  # noqa: F401,F403
from . import bilinear  # noqa: F401
