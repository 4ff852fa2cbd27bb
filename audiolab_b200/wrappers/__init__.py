"""Host-side mirror of the reference's wrapper plugin API (wrappers/base_wrapper.py, wrappers/separate.py)."""
from .base_wrapper import BaseWrapper, TypedInput  # noqa: F401
from .separate import Separate  # noqa: F401
