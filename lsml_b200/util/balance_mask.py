"""``balance_mask`` -- mirror of lsml/util/balance_mask.py (host side; RNG order preserved)."""
from ..core.model import balance_mask  # noqa: F401
