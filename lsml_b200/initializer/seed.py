"""Seed functions -- mirror of lsml/initializer/seed.py (host side, once per example)."""
import numpy


def center_of_mass_seeder(dataset_example):
    """ The centre of mass of ``dataset_example.seg`` in physical units (reference :5-31);
    ``dataset_example`` is anything with ``seg`` and ``dx`` attributes. """
    seg = dataset_example.seg
    indices = numpy.indices(seg.shape, dtype=float)
    all_slice = slice(None, None, None)
    new_axes_slices = (None,) * seg.ndim
    indices *= numpy.asarray(dataset_example.dx, dtype=float)[(all_slice,) + new_axes_slices]
    total = seg.sum()
    return numpy.array([(ind * seg).sum() / total for ind in indices])
