"""Provided image features -- mirror of lsml/feature/provided/image.py (device kernels in
lsml_b200/csrc/features.cu)."""
from ..base_feature import BaseImageFeature, GLOBAL_FEATURE_TYPE, LOCAL_FEATURE_TYPE


class ImageSample(BaseImageFeature):
    """ The gaussian-smoothed image sampled locally (reference :10-33) """
    locality = LOCAL_FEATURE_TYPE

    @property
    def name(self):
        return "Image sample (σ = {:.3f})".format(self.sigma)

    def spec(self):
        return ("image_sample", float(self.sigma))


class ImageEdgeSample(BaseImageFeature):
    """ The gaussian-smoothed image edge (gradient magnitude) sampled locally (:36-66) """
    locality = LOCAL_FEATURE_TYPE

    @property
    def name(self):
        return "Image edge sample (σ = {:.3f})".format(self.sigma)

    def spec(self):
        return ("image_edge", float(self.sigma))


class InteriorImageAverage(BaseImageFeature):
    """ Mean of the (smoothed) image over the band mask (:69-93; despite the name) """
    locality = GLOBAL_FEATURE_TYPE

    @property
    def name(self):
        return "Interior image average (σ = {:.3f})".format(self.sigma)

    def spec(self):
        return ("band_mean", float(self.sigma))


class InteriorImageVariation(BaseImageFeature):
    """ Population std of the (smoothed) image over the band mask (:96-121) """
    locality = GLOBAL_FEATURE_TYPE

    @property
    def name(self):
        return "Interior image variation (σ = {:.3f})".format(self.sigma)

    def spec(self):
        return ("band_std", float(self.sigma))


class COMRaySample(BaseImageFeature):
    """ The RAW image sampled along the segment from a voxel to the centre of mass of the
    current iterate and symmetrically outwards (:124-171); ``sigma`` only enters the name, as
    in the reference.  Device kernel: csrc/com_ray.cuh through ``assemble_kernel``. """
    locality = LOCAL_FEATURE_TYPE

    @property
    def name(self):
        return "COM Ray (N={}, σ = {:.3f})".format(self.n_samples, self.sigma)

    @property
    def size(self):
        return 2 * self.n_samples

    def __init__(self, ndim=2, sigma=3, n_samples=10):
        if ndim not in (2, 3):
            raise ValueError("`ndim` must be either 2 or 3")
        super().__init__(ndim, sigma)
        self.n_samples = n_samples

    def spec(self):
        return ("com_ray", int(self.n_samples))


def get_basic_image_features(ndim=2, sigmas=[0, 2]):
    """ Basic image features at multiple sigma values (:174-200) """
    feature_classes = [ImageSample, ImageEdgeSample, InteriorImageAverage,
                       InteriorImageVariation]
    return [feature_class(ndim=ndim, sigma=sigma)
            for feature_class in feature_classes
            for sigma in sigmas]
