from . import masked_gradient  # noqa: F401
from .masked_gradient import gradient_centered, gradient_magnitude_osher_sethian  # noqa: F401
