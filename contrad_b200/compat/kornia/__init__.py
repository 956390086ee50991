"""Import stub for ``kornia`` (absent; reference augment/__init__.py:4 imports
``kornia.filters.get_gaussian_kernel2d`` and ``filter2D`` at module level).  Only GaussianBlur
(not on the round-1 hot path) uses the math; the restatement lives in ``filters``."""
from . import filters  # noqa: F401
