"""InitializerBase -- mirror of lsml/initializer/initializer_base.py."""
import abc

import numpy

from ..util.distance_transform import distance_transform


class InitializerBase(abc.ABC):
    """ ``initializer(img, band, dx=, seed=) -> (u0, dist, mask)`` (reference :19-61) """

    def __init__(self):
        pass

    def __call__(self, img, band=0, dx=None, seed=None):
        if dx is None:
            dx = numpy.ones(img.ndim, dtype=float)
        else:
            dx = numpy.array(dx, dtype=float)
            if len(dx) != img.ndim:
                msg = "Number of dx terms ({}) doesn't match dimensions ({})"
                raise ValueError(msg.format(len(dx), img.ndim))
        if seed is None:
            seed = numpy.array(img.shape) / 2
        else:
            seed = numpy.array(seed)
        init_mask = self.initialize(img=img, dx=dx, seed=seed)
        if not isinstance(init_mask, numpy.ndarray):
            msg = "Returned initializer was type {} but should be numpy.ndarray"
            raise TypeError(msg.format(type(init_mask)))
        if init_mask.dtype != bool:
            msg = "Returned initializer was dtype {} but should be bool"
            raise TypeError(msg.format(init_mask.dtype))
        if init_mask.shape != img.shape:
            msg = "Returned initializer was shape {} but should be {}"
            raise ValueError(msg.format(init_mask.shape, img.shape))
        u = 2 * init_mask.astype(float) - 1
        dist, mask = distance_transform(arr=u, band=band, dx=dx)
        return u, dist, mask

    @abc.abstractmethod
    def initialize(self, img, dx, seed):
        raise NotImplementedError
