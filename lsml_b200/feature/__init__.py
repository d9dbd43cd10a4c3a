# flake8: noqa
from .base_feature import BaseFeature, BaseImageFeature, BaseShapeFeature
from .feature_map import FeatureMap
from .provided.image import (
    get_basic_image_features,
    COMRaySample,
    ImageEdgeSample,
    ImageSample,
    InteriorImageAverage,
    InteriorImageVariation,
)
from .provided.shape import (
    BoundarySize,
    DistanceToCenterOfMass,
    get_basic_shape_features,
    IsoperimetricRatio,
    Moments,
    Size,
)
from .user import PrecomputedImageFeature, PrevIterFeature, SmoothedLevelSet
