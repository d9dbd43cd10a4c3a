"""User-feature registry: features beyond ``lsml.feature.provided`` that still run on the device.

The reference lets users subclass ``BaseFeature`` and compute anything in Python
(examples/gestalt2d/corner_feature.py, prev_iter_feature.py).  lsml_b200 has no CPU fallback on
the hot path, so a user feature has to map onto device work.  Two families cover the reference's
own examples:

* ``SmoothedLevelSet`` -- ``scipy.ndimage.gaussian_filter(u, sigma)`` sampled on the band
  (the example's ``PrevIterFeature``): recomputed from ``u`` on the device every iteration with
  the separable Gaussian kernel that also builds the image planes (scipy's arithmetic order).
* ``PrecomputedImageFeature`` -- any u-INDEPENDENT image feature (the example's ``CornerFeature``:
  a Harris response of the smoothed image).  The user implements ``compute_plane(img, dx)`` in
  Python; it runs ONCE per image when the image is loaded, the resulting plane stays in HBM and is
  sampled on the band every iteration like ``ImageSample``.  Nothing of it is on the per-iteration
  path.

Anything else (a u-dependent feature with arbitrary Python code) raises ``NotImplementedError``
from ``BaseFeature.spec``.
"""
from .base_feature import BaseImageFeature, BaseShapeFeature, LOCAL_FEATURE_TYPE


class SmoothedLevelSet(BaseShapeFeature):
    """ The level set of the previous iteration smoothed by a Gaussian of parameter ``sigma``
    (in voxels, not divided by dx): examples/gestalt2d/prev_iter_feature.py:6-22, same name. """
    locality = LOCAL_FEATURE_TYPE

    @property
    def name(self):
        return "smooth-ls-sigma={:.2f}".format(self.sigma)

    def __init__(self, ndim=2, sigma=2):
        super().__init__(ndim)
        self.sigma = sigma

    def spec(self):
        return ("smooth_u", float(self.sigma))


#: the name the reference's example uses
PrevIterFeature = SmoothedLevelSet


class PrecomputedImageFeature(BaseImageFeature):
    """ Base class for u-independent image features computed by user code once per image.

    Subclasses implement ``name`` and ``compute_plane(img, dx) -> float64 array of img.shape``
    (``img`` is the normalised image, as ``compute_feature`` receives it in the reference). """
    locality = LOCAL_FEATURE_TYPE

    def compute_plane(self, img, dx):
        raise NotImplementedError

    def spec(self):
        return ("user_plane", self.name, self.compute_plane)
