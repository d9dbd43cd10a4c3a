"""Provided shape features -- mirror of lsml/feature/provided/shape.py (device kernels in
lsml_b200/csrc/features.cu)."""
from ..base_feature import BaseShapeFeature, GLOBAL_FEATURE_TYPE, LOCAL_FEATURE_TYPE


class Size(BaseShapeFeature):
    """ Length / area / volume of {u > 0} (reference :9-32) """
    locality = GLOBAL_FEATURE_TYPE

    @property
    def name(self):
        if self.ndim == 1:
            return 'Length'
        elif self.ndim == 2:
            return 'Area'
        elif self.ndim == 3:
            return 'Volume'
        else:
            return 'Hyper-volume'

    def spec(self):
        return ("size",)


class BoundarySize(BaseShapeFeature):
    """ Size of the zero level set (:35-89): curve length in 2-D (marching squares), surface area
    in 3-D (marching cubes as a reduction, csrc/mcubes.cu; the reference's scikit-image Lewiner
    variant is a third-party algorithm: parity unpinned beyond test_shape.py:59-107). """
    locality = GLOBAL_FEATURE_TYPE

    def __init__(self, ndim=2):
        if ndim < 2 or ndim > 3:
            msg = ("Boundary size is only defined for dimensions 2 and 3; "
                   "ndim provided = {}")
            raise ValueError(msg.format(ndim))
        super(BoundarySize, self).__init__(ndim)

    @property
    def name(self):
        if self.ndim == 2:
            return 'Curve length'
        elif self.ndim == 3:
            return 'Surface area'

    def spec(self):
        return ("boundary_size",)


class IsoperimetricRatio(BaseShapeFeature):
    """ Circularity 4 pi A / L^2 (2-D) or sphericity 36 pi V^2 / A^3 (3-D) (:92-153) """
    locality = GLOBAL_FEATURE_TYPE

    @property
    def name(self):
        if self.ndim == 2:
            return 'Circularity'
        else:
            return 'Sphericity'

    def __init__(self, ndim=2):
        if ndim < 2 or ndim > 3:
            msg = ("Isoperimetric ratio defined for dimensions 2 and 3; "
                   "ndim provided = {}")
            raise ValueError(msg.format(ndim))
        super(IsoperimetricRatio, self).__init__(ndim)

    def spec(self):
        return ("isoperimetric",)


class Moments(BaseShapeFeature):
    """ Normalised statistical moments of {u > 0} along given axes (:156-238) """
    locality = GLOBAL_FEATURE_TYPE

    @property
    def name(self):
        return "Moments (axes={}; orders={})".format(self.axes, self.orders)

    @property
    def size(self):
        return len(self.axes) * len(self.orders)

    def __init__(self, ndim=2, axes=(0, 1), orders=(1, 2)):
        super(Moments, self).__init__(ndim)
        for axis in axes:
            if axis < 0 or axis > ndim - 1:
                msg = "axis provided ({}) must be one of 0 ... {}"
                raise ValueError(msg.format(axis, ndim - 1))
        for order in orders:
            if order < 1:
                raise ValueError("Moments order should be greater than or equal to 1")
        self.axes = axes
        self.orders = orders

    def spec(self):
        return ("moments", tuple(self.axes), tuple(self.orders))


class DistanceToCenterOfMass(BaseShapeFeature):
    """ Distance of each band voxel to the centre of mass of {u > 0} (:241-267) """
    locality = LOCAL_FEATURE_TYPE

    @property
    def name(self):
        return "Distance to center of mass"

    def spec(self):
        return ("dist_to_com",)


def get_basic_shape_features(ndim=2, moment_orders=[1, 2]):
    """ Basic shape features (:270-297; like the reference, `moment_orders` is ignored and
    Moments keeps its default axes (0, 1) even for ndim == 3) """
    feature_classes = [BoundarySize, DistanceToCenterOfMass, IsoperimetricRatio, Moments, Size]
    return [feature_class(ndim=ndim) for feature_class in feature_classes]
