"""``on_iterate`` helpers for ``LevelSetMachineLearning.segment`` -- mirror of
lsml/util/on_iterate.py (``plot_contours`` is visualisation and not reproduced)."""


def collect_scores(seg, score_func):
    """ A callback ``f(i, u)`` that appends ``score_func(u, seg)`` to ``f.scores``
    (reference :6-23). """
    scores = []

    def on_iterate(i, u):
        scores.append(score_func(u, seg))

    on_iterate.scores = scores
    return on_iterate
