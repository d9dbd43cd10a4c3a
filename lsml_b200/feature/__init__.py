"""``lsml_b200.feature`` -- the names ``lsml.feature`` exports (feature/__init__.py:3-18), plus the
user-feature registry of lsml_b200/feature/user.py."""
from . import base_feature as _base
from . import feature_map as _map
from . import user as _user
from .provided import image as _image
from .provided import shape as _shape

BaseFeature, BaseImageFeature, BaseShapeFeature = (
    _base.BaseFeature, _base.BaseImageFeature, _base.BaseShapeFeature)
FeatureMap = _map.FeatureMap

_IMAGE_NAMES = ("ImageSample", "ImageEdgeSample", "InteriorImageAverage", "InteriorImageVariation",
                "COMRaySample", "get_basic_image_features")
_SHAPE_NAMES = ("Size", "BoundarySize", "IsoperimetricRatio", "Moments", "DistanceToCenterOfMass",
                "get_basic_shape_features")
_USER_NAMES = ("SmoothedLevelSet", "PrevIterFeature", "PrecomputedImageFeature")
for _module, _names in ((_image, _IMAGE_NAMES), (_shape, _SHAPE_NAMES), (_user, _USER_NAMES)):
    for _name in _names:
        globals()[_name] = getattr(_module, _name)

__all__ = ["BaseFeature", "BaseImageFeature", "BaseShapeFeature", "FeatureMap",
           *_IMAGE_NAMES, *_SHAPE_NAMES, *_USER_NAMES]
