"""Feature base classes -- host-side mirror of lsml/feature/base_feature.py.

Same class names, constructor arguments, abstract properties and ``__call__`` validation
(reference :67-139).  The difference is where the numbers come from: every provided feature
exposes ``spec()``, a plain tuple that lsml_b200.core.engine compiles into a column of the
device feature program; ``compute_feature`` of the provided classes runs that program on the
GPU.  A user subclass that only implements a Python ``compute_feature`` has no device kernel:
using it in a FeatureMap raises ``NotImplementedError`` (no CPU fallback).
"""
import abc

import numpy

IMAGE_FEATURE_TYPE = 'image'
SHAPE_FEATURE_TYPE = 'shape'
LOCAL_FEATURE_TYPE = 'local'
GLOBAL_FEATURE_TYPE = 'global'


class BaseFeature(abc.ABC):
    """ The abstract base class for all features (reference :13-139) """

    #: number of feature columns this class produces
    size = 1

    @property
    @abc.abstractmethod
    def name(self):
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def type(self):
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def locality(self):
        raise NotImplementedError

    def __eq__(self, other):
        return isinstance(other, type(self)) and self.name == other.name

    def __hash__(self):
        return hash(self.name)

    def __str__(self):
        return self.name

    def __repr__(self):
        return self.__str__()

    def __init__(self, ndim=2):
        self.ndim = ndim

    def spec(self):
        """Tuple describing the device kernel column(s) of this feature."""
        raise NotImplementedError(
            "feature {!r} ({}) has no device kernel registered; lsml_b200 has no CPU "
            "fallback for Python-level compute_feature implementations".format(
                self.name, type(self).__name__))

    def __call__(self, u, img=None, dist=None, mask=None, dx=None):
        """ Validates the inputs (reference :67-139) and computes the feature on the GPU """
        not_ndarray = "`{var}` must be a numpy array"
        if not isinstance(u, numpy.ndarray):
            raise TypeError(not_ndarray.format(var='u'))
        shape = u.shape
        shape_mismatch = ("`{{var}}` has shape {{shape}}, "
                          "but must be shape {correct_shape}")
        shape_mismatch = shape_mismatch.format(correct_shape=shape)

        if isinstance(self, BaseImageFeature):
            if img is None:
                raise ValueError("`img` must be present for image features")
            if not isinstance(img, numpy.ndarray):
                raise TypeError(not_ndarray.format(var='img'))
            if img.shape != shape:
                raise ValueError(shape_mismatch.format(var='img', shape=img.shape))

        if dist is not None:
            if not isinstance(dist, numpy.ndarray):
                raise TypeError(not_ndarray.format(var='dist'))
            if dist.shape != shape:
                raise ValueError(shape_mismatch.format(var='img', shape=dist.shape))

        if dx is None:
            dx = numpy.ones(self.ndim, dtype=float)
        else:
            dx = numpy.array(dx, dtype=float)
            if len(dx) != self.ndim:
                msg = "Number of dx terms ({}) doesn't match dimensions ({})"
                raise ValueError(msg.format(len(dx), self.ndim))

        if mask is None:
            mask = numpy.ones(shape, dtype=bool)
        if not isinstance(mask, numpy.ndarray):
            raise TypeError(not_ndarray.format(var='mask'))
        if mask.shape != shape:
            raise ValueError(shape_mismatch.format(var='mask', shape=mask.shape))
        if mask.dtype != bool:
            raise TypeError("`mask` dtype ({}) was not of type bool".format(mask.dtype))

        if not mask.any():
            return numpy.empty(u.shape + (self.size,))

        if isinstance(self, BaseImageFeature):
            return self.compute_feature(u=u, img=img, dist=dist, mask=mask, dx=dx)
        elif isinstance(self, BaseShapeFeature):
            return self.compute_feature(u=u, dist=dist, mask=mask, dx=dx)

    def compute_feature(self, u, img=None, dist=None, mask=None, dx=None):
        """ Dense result ``u.shape`` (``u.shape + (size,)`` for multi-column features), defined
        on ``mask`` (zero elsewhere), computed by the device feature program. """
        from .feature_map import compute_dense
        out = compute_dense([self], u, img, mask, dx)
        return out if self.size > 1 else out[..., 0]


class BaseImageFeature(BaseFeature):
    type = IMAGE_FEATURE_TYPE

    def __init__(self, ndim=2, sigma=3):
        super(BaseImageFeature, self).__init__(ndim)
        self.sigma = sigma


class BaseShapeFeature(BaseFeature):
    type = SHAPE_FEATURE_TYPE
