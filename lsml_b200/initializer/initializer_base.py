"""InitializerBase -- drop-in for lsml/initializer/initializer_base.py.

A subclass provides ``initialize(img, dx, seed) -> bool ndarray``; calling the instance returns
``(u0, dist, mask)`` with ``u0 = 2*init - 1`` and the narrow band of ``u0`` (reference :19-61; the
band comes from the device marcher instead of scikit-fmm).  Argument handling and error messages
are the reference's."""
import abc

import numpy

from ..util.distance_transform import distance_transform


def _spacing(dx, ndim):
    """``dx`` as a float vector of length ``ndim`` (ones when omitted)."""
    if dx is None:
        return numpy.ones(ndim, dtype=float)
    dx = numpy.array(dx, dtype=float)
    if len(dx) != ndim:
        raise ValueError("Number of dx terms ({}) doesn't match dimensions ({})".format(len(dx), ndim))
    return dx


def _checked(init_mask, shape):
    """The user's ``initialize`` result, validated like the reference (:41-53)."""
    if not isinstance(init_mask, numpy.ndarray):
        raise TypeError("Returned initializer was type {} but "
                        "should be numpy.ndarray".format(type(init_mask)))
    if init_mask.dtype != bool:
        raise TypeError("Returned initializer was dtype {} but should be bool".format(init_mask.dtype))
    if init_mask.shape != shape:
        raise ValueError("Returned initializer was shape {} but should be {}".format(
            init_mask.shape, shape))
    return init_mask


class InitializerBase(abc.ABC):

    def __init__(self):
        pass

    def __call__(self, img, band=0, dx=None, seed=None):
        dx = _spacing(dx, img.ndim)
        seed = numpy.array(img.shape) / 2 if seed is None else numpy.array(seed)
        inside = _checked(self.initialize(img=img, dx=dx, seed=seed), img.shape)
        u = 2 * inside.astype(float) - 1
        dist, mask = distance_transform(arr=u, band=band, dx=dx)
        return u, dist, mask

    @abc.abstractmethod
    def initialize(self, img, dx, seed):
        raise NotImplementedError
