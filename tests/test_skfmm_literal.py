"""oracle/skfmm_literal.c (the upstream-literal restatement of scikit-fmm used as the measuring
stick of tests/tools/band_gap_report.py) against the reference's own known answers
(/root/reference/lsml/util/test/test_distance_transform.py:11-53), and against the product's
marcher where the two must agree: inputs on which no quadratic is non-causal."""
import numpy as np

from oracle import oracle as O


def test_reference_known_answer_1d():          # test_distance_transform.py:11-21
    arr = np.r_[-1, -1, 1, -1, -1.]
    dist, mask = O.distance_transform(arr, 1, [1.], literal=True)
    assert (dist == np.r_[0., -0.5, 0.5, -0.5, 0.]).all()
    assert (mask == np.r_[False, True, True, True, False]).all()


def test_reference_edge_cases():               # test_distance_transform.py:23-53
    d, m = O.distance_transform(np.ones((4, 5, 6)), 0, [1, 1, 1.], literal=True)
    assert m.sum() == 0 and (d == np.inf).all()
    d, m = O.distance_transform(-np.ones((4, 5, 6)), 0, [1, 1, 1.], literal=True)
    assert m.sum() == 0 and (d == -np.inf).all()
    arr = np.ones((4, 5, 6)); arr[2] = -1
    d, m = O.distance_transform(arr, 0, [1, 1, 1.], literal=True)
    assert m.sum() == arr.size
    d, m = O.distance_transform(np.zeros((4, 5, 6)), 0, [1, 1, 1.], literal=True)
    assert m.sum() == 24 * 5 and (d == 0).all()


def test_literal_equals_product_marcher_on_smooth_inputs():
    """On a smooth signed-distance-like input the two marchers agree in the band mask; the
    distances differ slightly where the product marcher's causality rule demotes a second-order
    contributor (a few 1e-3 at most), and not at all with that rule switched back."""
    n = 40
    g = np.indices((n, n, n)).astype(float) - n / 2 + 0.25
    u = 11.3 - np.sqrt((g ** 2).sum(0))
    a = O.distance_transform(u, 3.0, np.ones(3))
    b = O.distance_transform(u, 3.0, np.ones(3), literal=True)
    assert np.array_equal(a[1], b[1])
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-2)
    try:
        O._lib().orc_fmm_set_ablation(2)
        c = O.distance_transform(u, 3.0, np.ones(3))
    finally:
        O._lib().orc_fmm_set_ablation(0)
    assert np.array_equal(c[1], b[1]) and np.array_equal(c[0], b[0])


def test_all_four_departures_switched_back_is_the_literal_marcher():
    """orc_fmm_distance with its ablation flags all set is the same algorithm as the literal
    file up to heap tie mechanics -- on a noisy input (no exact ties) bit for bit."""
    rs = np.random.RandomState(3)
    n = 32
    g = np.indices((n, n, n)).astype(float) - n / 2
    u = 9.0 - np.sqrt((g ** 2).sum(0)) + 0.4 * rs.standard_normal((n, n, n))
    try:
        O._lib().orc_fmm_set_ablation(15)
        a = O.distance_transform(u, 3.0, np.ones(3))
    finally:
        O._lib().orc_fmm_set_ablation(0)
    b = O.distance_transform(u, 3.0, np.ones(3), literal=True)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
