"""Overlap scores -- drop-in for lsml/score_functions.py (numpy inputs; the batched device
counters behind ``fit`` are in lsml_b200.core.model.overlap_scores).

``jaccard(u, seg, threshold=0.0)`` is |{u > threshold} & seg| / |{u > threshold} | seg| with the
reference's convention that two empty sets overlap perfectly (:4-23); ``dice`` is derived from it
as 2J / (J + 1) (:26-30)."""
import numpy


def _overlap_counts(u, seg, threshold):
    if seg.dtype != bool:
        raise ValueError("`seg` dtype ({}) was not of type bool".format(seg.dtype))
    inside = u > threshold
    return float(numpy.count_nonzero(inside & seg)), float(numpy.count_nonzero(inside | seg))


def jaccard(u, seg, threshold=0.0):
    both, either = _overlap_counts(u, seg, threshold)
    return both / either if either else 1.0


def dice(u, seg, threshold=0.0):
    j = jaccard(u, seg, threshold)
    return 2.0 * j / (j + 1.0)
