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


def _require_array(value, var):
    if not isinstance(value, numpy.ndarray):
        raise TypeError("`{var}` must be a numpy array".format(var=var))


def _require_shape(value, shape, var):
    if value.shape != shape:
        raise ValueError("`{var}` has shape {shape}, but must be shape {correct_shape}".format(
            var=var, shape=value.shape, correct_shape=shape))


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
        """ Validates the inputs in the reference's order and with its messages (reference
        :67-139), then computes the feature on the GPU. """
        _require_array(u, 'u')
        image_feature = isinstance(self, BaseImageFeature)
        if image_feature:
            if img is None:
                raise ValueError("`img` must be present for image features")
            _require_array(img, 'img')
            _require_shape(img, u.shape, 'img')
        if dist is not None:
            _require_array(dist, 'dist')
            _require_shape(dist, u.shape, 'img')      # (the reference names `img` here too)
        if dx is None:
            dx = numpy.ones(self.ndim, dtype=float)
        else:
            dx = numpy.array(dx, dtype=float)
            if len(dx) != self.ndim:
                raise ValueError("Number of dx terms ({}) doesn't match dimensions ({})".format(
                    len(dx), self.ndim))
        if mask is None:
            mask = numpy.ones(u.shape, dtype=bool)
        _require_array(mask, 'mask')
        _require_shape(mask, u.shape, 'mask')
        if mask.dtype != bool:
            raise TypeError("`mask` dtype ({}) was not of type bool".format(mask.dtype))
        if not mask.any():
            return numpy.empty(u.shape + (self.size,))
        if image_feature:
            return self.compute_feature(u=u, img=img, dist=dist, mask=mask, dx=dx)
        if isinstance(self, BaseShapeFeature):
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
