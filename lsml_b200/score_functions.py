"""Scores -- mirror of lsml/score_functions.py (numpy inputs; the batched device counters are
in lsml_b200.core.engine.BandEngine.overlap_scores)."""
import numpy


def jaccard(u, seg, threshold=0.0):
    if seg.dtype != bool:
        msg = "`seg` dtype ({}) was not of type bool"
        raise ValueError(msg.format(seg.dtype))
    thresholded = u > threshold
    intersection = float((thresholded & seg).sum())
    union = float((thresholded | seg).sum())
    if union == 0:
        return 1.0
    return intersection / union


def dice(u, seg, threshold=0.0):
    jaccard_score = jaccard(u, seg, threshold)
    return 2.0 * jaccard_score / (jaccard_score + 1.0)
