"""``ThresholdInitializer`` -- mirror of lsml/initializer/provided/threshold.py.

Host side, once per image.  The reference takes the threshold from
``skimage.filters.thresholding.threshold_otsu`` (scikit-image: third party, unpinned, not
installed here); ``otsu_threshold`` below restates the published algorithm as scikit-image
implements it (256-bin histogram over [min, max], threshold = centre of the bin that maximises
the between-class variance).  PARITY UNPINNED beyond the reference's own test
(initializer/provided/test/test_threshold.py:8-26).
"""
import numpy
from scipy.ndimage import gaussian_filter

from ..initializer_base import InitializerBase


def otsu_threshold(image, nbins=256):
    """ Otsu's threshold of a float image (see the module docstring). """
    image = numpy.asarray(image, dtype=float)
    first = image.ravel()[0]
    if (image == first).all():
        return float(first)
    counts, edges = numpy.histogram(image.ravel(), bins=nbins,
                                    range=(image.min(), image.max()))
    centers = (edges[:-1] + edges[1:]) / 2.0
    counts = counts.astype(float)
    weight1 = numpy.cumsum(counts)
    weight2 = numpy.cumsum(counts[::-1])[::-1]
    with numpy.errstate(invalid="ignore", divide="ignore"):
        mean1 = numpy.cumsum(counts * centers) / weight1
        mean2 = (numpy.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return float(centers[numpy.nanargmax(variance12)])


class ThresholdInitializer(InitializerBase):
    """ The Otsu threshold of the Gaussian-smoothed image (reference :8-40) """

    def __init__(self, sigma):
        if sigma < 0:
            raise ValueError("sigma ({}) was negative".format(sigma))
        self.sigma = sigma

    def initialize(self, img, dx, seed):
        blur = img.copy() if self.sigma == 0 else gaussian_filter(img, self.sigma)
        return blur >= otsu_threshold(blur)
