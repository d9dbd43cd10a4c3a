"""``distance_transform`` -- mirror of lsml/util/distance_transform.py on the device.

The reference wraps ``skfmm.distance(arr, narrow=band, dx=dx)`` (:49) and special-cases
all-zero (:36-39) and single-sign (:42-47) inputs; ``band == 0`` means "no narrow band"
(:54-57).  Here the narrow-band fast-marching solution is computed by
lsml_b200/csrc/fmm.cu; the edge cases are decided on the device from exact counts.
"""
import numpy

from ..core.engine import BandEngine


def distance_transform(arr, band, dx):
    """ Returns ``(dist, mask)`` as numpy arrays, like the reference. """
    arr = numpy.ascontiguousarray(arr, dtype=numpy.float64)
    dx = numpy.asarray(dx, dtype=numpy.float64)
    eng = BandEngine(arr.shape, 1, dx=dx, band=band, specs=[])
    eng.set_level_sets(arr[None])
    dist = eng.distance()[0].cpu().numpy()
    mask = eng.mask[0].cpu().numpy().astype(bool)
    if not mask.any():
        # single-sign input: +/- inf everywhere (reference :42-47)
        sign = numpy.sign(arr.ravel()[0])
        dist = sign * numpy.full(arr.shape, numpy.inf)
    return dist, mask
