"""Analysis tools on the hot path (the reference's tools/__init__.py:5-7 also exports periodic-box and
array-reorder helpers; those are out of scope here)."""
from ._ArrayReorder import ArrayReorder, match_ids  # noqa: F401,E402  (lazy: needs CUDA only when called)
