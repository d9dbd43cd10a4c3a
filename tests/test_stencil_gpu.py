"""GPU parity: dense masked stencils (a1, a2) and band compaction (a3) against the oracle.

Mirrors lsml/gradient/test/test_masked_gradient.py (same seed, dims 50-100, anisotropic dx,
mask = arr > 0, normalize) but demands BIT-EXACT equality with the CPU path instead of the
reference's 1e-8 mean-abs bound, because the band mask downstream is judged bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(seed=1234):
    rs = np.random.RandomState(seed)
    for ndim in (1, 2, 3):
        dims = rs.randint(50, 101, size=ndim)
        arr = rs.randn(*dims)
        nu = rs.randn(*dims)
        yield ndim, arr, nu, None, None
        yield ndim, arr, nu, arr > 0, None
        yield ndim, arr, nu, rs.rand(*dims) > 0.9, rs.rand(ndim) + 0.25
        yield ndim, arr, nu, None, np.array([0.5, 2.0, 4.0][:ndim])


def test_gmag_os_host_pointer_dropin_bit_exact(oracle):
    from lsml_b200.gradient import masked_gradient as mg
    for ndim, arr, nu, mask, dx in _cases():
        got = mg.gradient_magnitude_osher_sethian(arr, nu, mask=mask, dx=dx)
        want = oracle.gmag_os(arr, nu, mask, dx)
        assert got.dtype == np.float64 and got.shape == arr.shape
        assert np.array_equal(got, want), (ndim, np.abs(got - want).max())
        if oracle.reference_lib() is not None:
            assert np.array_equal(got, oracle.gmag_os(arr, nu, mask, dx, use_reference=True))


def test_gradient_centered_host_pointer_dropin_bit_exact(oracle):
    from lsml_b200.gradient import masked_gradient as mg
    for ndim, arr, nu, mask, dx in _cases(4321):
        for normalize in (False, True):
            grads, gmag = mg.gradient_centered(arr, mask=mask, dx=dx, normalize=normalize)
            wg, wm = oracle.gradient_centered(arr, mask, dx, normalize)
            assert np.array_equal(gmag, wm)
            for a in range(ndim):
                assert np.array_equal(grads[a], wg[a])


def test_reference_test_suite_properties():
    """The reference's own assertions (test_masked_gradient.py:13-120) against numpy."""
    from lsml_b200.gradient import masked_gradient as mg
    rs = np.random.RandomState(1234)
    for ndim in (1, 2, 3):
        dims = rs.randint(50, 101, size=ndim)
        arr = rs.randn(*dims)
        dx = rs.rand(ndim)
        grads, gmag = mg.gradient_centered(arr, dx=dx)
        ng = np.gradient(arr, *dx)
        ng = [ng] if ndim == 1 else ng
        for i in range(ndim):
            assert np.abs(grads[i] - ng[i]).mean() <= 1e-8
        assert np.abs(gmag - np.linalg.norm(ng, axis=0)).mean() <= 1e-8
        unit = mg.gradient_centered(arr, normalize=True, return_gradient_magnitude=False)
        assert np.abs(np.linalg.norm(unit, axis=0) - 1).mean() <= 1e-8
    arr = rs.randn(141, 112)
    nu = rs.randn(141, 112)
    dx = rs.rand(2)
    di = np.diff(arr, axis=0) / dx[0]
    dj = np.diff(arr, axis=1) / dx[1]
    fwdi, bcki = np.vstack([di, di[-1]]), np.vstack([di[0], di])
    fwdj, bckj = np.c_[dj, dj[:, -1]], np.c_[dj[:, 0], dj]
    plus = np.sqrt(np.maximum(bcki, 0) ** 2 + np.minimum(fwdi, 0) ** 2 +
                   np.maximum(bckj, 0) ** 2 + np.minimum(fwdj, 0) ** 2)
    minus = np.sqrt(np.minimum(bcki, 0) ** 2 + np.maximum(fwdi, 0) ** 2 +
                    np.minimum(bckj, 0) ** 2 + np.maximum(fwdj, 0) ** 2)
    want = np.zeros_like(arr)
    want[nu > 0] = minus[nu > 0]
    want[nu < 0] = plus[nu < 0]
    got = mg.gradient_magnitude_osher_sethian(arr, nu, dx=dx)
    assert np.abs(got - want).mean() <= 1e-8


def test_tma_pipeline_on_batches(oracle):
    """Items of a batch are the 4th tensor-map coordinate: halos never leak across items."""
    import torch
    from lsml_b200 import _lib
    from lsml_b200.gradient import masked_gradient as mg
    assert _lib.lib.lsml_stencil_use_tma(-1) == 1
    rs = np.random.RandomState(17)
    arr = rs.randn(3, 37, 26, 80)
    nu = rs.randn(*arr.shape)
    t = lambda a: torch.from_numpy(a).cuda()
    for mask in (None, rs.rand(*arr.shape) > 0.4):
        got = mg.gradient_magnitude_osher_sethian(t(arr), t(nu), mask=None if mask is None else t(mask),
                                                  dx=np.array([1.0, 2.0, 0.5]), batched=True)
        for b in range(arr.shape[0]):
            want = oracle.gmag_os(arr[b], nu[b], None if mask is None else mask[b],
                                  np.array([1.0, 2.0, 0.5]))
            assert np.array_equal(got[b].cpu().numpy(), want), b


def test_device_tensor_path_and_batch(oracle):
    import torch
    from lsml_b200.gradient import masked_gradient as mg
    rs = np.random.RandomState(7)
    arr = rs.randn(5, 33, 47, 65)
    nu = rs.randn(*arr.shape)
    mask = rs.rand(*arr.shape) > 0.5
    dx = np.array([1.0, 0.7, 1.3])
    t = lambda a: torch.from_numpy(a).cuda()
    got = mg.gradient_magnitude_osher_sethian(t(arr), t(nu), mask=t(mask), dx=dx, batched=True)
    grads, gm = mg.gradient_centered(t(arr), mask=t(mask), dx=dx, batched=True)
    for b in range(arr.shape[0]):
        assert np.array_equal(got[b].cpu().numpy(), oracle.gmag_os(arr[b], nu[b], mask[b], dx))
        wg, wm = oracle.gradient_centered(arr[b], mask[b], dx)
        assert np.array_equal(gm[b].cpu().numpy(), wm)
        for a in range(3):
            assert np.array_equal(grads[a][b].cpu().numpy(), wg[a])
    # float32 variant: same expressions evaluated in fp32, compared at fp32 tolerance
    got32 = mg.gradient_magnitude_osher_sethian(t(arr).float(), t(nu).float(), mask=t(mask),
                                                dx=dx, batched=True)
    want = np.stack([oracle.gmag_os(arr[b], nu[b], mask[b], dx) for b in range(arr.shape[0])])
    assert np.allclose(got32.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


def test_outputs_untouched_outside_mask():
    """Contract of the C ABI: only masked voxels are written (masked_gradient.c `continue`)."""
    import ctypes
    from lsml_b200 import _lib
    from lsml_b200.gradient import masked_gradient as mg
    rs = np.random.RandomState(3)
    arr, nu = rs.randn(40, 50), rs.randn(40, 50)
    mask = rs.rand(40, 50) > 0.5
    out = np.full_like(arr, 7.25)
    mg._host_gmag_os_func(2)(40, 50, arr, mask, nu, out, 1.0, 1.0)
    assert (out[~mask] == 7.25).all() and (out[mask] != 7.25).any()


@pytest.mark.parametrize("shape,batch,density", [((51, 51), 7, 0.1), ((64, 64, 64), 3, 0.02),
                                                 ((5,), 1, 0.5), ((4099,), 2, 0.0),
                                                 ((37, 11, 13), 2, 1.0)])
def test_band_compaction_matches_numpy_order(shape, batch, density):
    import torch
    from lsml_b200.core import band
    rs = np.random.RandomState(11)
    mask = rs.rand(batch, *shape) < density
    idx, offsets = band.compact(torch.from_numpy(mask).cuda())
    want = np.flatnonzero(mask.ravel())
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), want)
    per_item = mask.reshape(batch, -1).sum(axis=1)
    assert np.array_equal(offsets.cpu().numpy(), np.r_[0, np.cumsum(per_item)])


@pytest.fixture(params=["tma", "march"])
def generation(request):
    """Both generations of the dense float64 3-D kernel: the TMA plane pipeline and the
    register-marching kernels it falls back to (lsml_stencil_use_tma)."""
    from lsml_b200 import _lib
    prev = _lib.lib.lsml_stencil_use_tma(1 if request.param == "tma" else 0)
    yield request.param
    _lib.lib.lsml_stencil_use_tma(prev)


@pytest.mark.parametrize("shape,dx", [((20, 17, 64), None), ((19, 24, 128), (1.0, 1.0, 1.0)),
                                      ((18, 9, 96), (0.5, 2.0, 4.0)), ((21, 16, 72), (0.7, 1.3, 1.1)),
                                      ((35, 8, 64), None), ((2, 2, 64), None),
                                      ((70, 21, 130), None), ((66, 19, 128), (1.0, 0.5, 2.0))])
def test_vectorised_march_kernel_bit_exact_all_inputs(oracle, shape, dx, generation):
    """The 16-byte vectorised 3-D kernels (even last axis >= 64; TMA-staged or register-marching)
    evaluate the upwind sum in a scaled form; they must still be bit-identical to the reference
    for ordinary values, with and without masks, on partial tiles and across plane chunks, and for
    the values where the scaled form is not exact by itself (subnormal squares, infinities, NaNs,
    exact zeros), which take the kernel's exact slow path."""
    import torch
    from lsml_b200.gradient import masked_gradient as mg
    rs = np.random.RandomState(11)
    arr = rs.randn(*shape)
    nu = rs.randn(*shape)
    dxa = None if dx is None else np.asarray(dx)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for mask in (None, rs.rand(*shape) > 0.3, rs.rand(*shape) > 0.97):
        got = mg.gradient_magnitude_osher_sethian(t(arr), t(nu), mask=None if mask is None else t(mask),
                                                  dx=dxa).cpu().numpy()
        want = oracle.gmag_os(arr, nu, mask, dxa, use_reference=oracle.reference_lib() is not None)
        assert np.array_equal(got, want), np.abs(got - want).max()
        # the centered gradient runs through the same plane pipeline (masked_gradient.c:13-75)
        for normalize in (False, True):
            grads, gm = mg.gradient_centered(t(arr), mask=None if mask is None else t(mask), dx=dxa,
                                             normalize=normalize)
            wg, wm = oracle.gradient_centered(arr, mask, dxa, normalize,
                                              use_reference=oracle.reference_lib() is not None)
            assert np.array_equal(gm.cpu().numpy(), wm)
            for a in range(3):
                assert np.array_equal(grads[a].cpu().numpy(), wg[a]), (a, normalize)
    # special values
    sp = arr.copy()
    flat = sp.ravel()
    idx = rs.choice(flat.size, size=400, replace=flat.size < 400)
    flat[idx[:80]] *= 1e-170                  # squares underflow / are subnormal
    flat[idx[80:160]] *= 1e-200
    flat[idx[160:200]] = 0.0
    flat[idx[200:240]] = np.inf
    flat[idx[240:280]] = -np.inf
    flat[idx[280:320]] = np.nan
    flat[idx[320:360]] *= 1e160               # squares overflow
    sp[..., :8] = 3.0                         # flat region: exact zeros of the gradient
    tiny = arr * 1e-165                       # everything tiny: sum in the subnormal range
    for a in (sp, tiny):
        with np.errstate(all="ignore"):
            got = mg.gradient_magnitude_osher_sethian(t(a), t(nu), dx=dxa).cpu().numpy()
            want = oracle.gmag_os(a, nu, None, dxa, use_reference=oracle.reference_lib() is not None)
        assert np.array_equal(got, want, equal_nan=True), \
            np.flatnonzero(~((got == want) | (np.isnan(got) & np.isnan(want))).ravel())[:10]
    # float32 build of the same kernel (last axis % 4 == 0 and >= 128 takes the vector path)
    if shape[2] % 4 == 0:
        a32 = np.tile(arr, (1, 1, 2)).astype(np.float32)
        n32 = np.tile(nu, (1, 1, 2)).astype(np.float32)
        got = mg.gradient_magnitude_osher_sethian(t(a32), t(n32), dx=dxa).cpu().numpy()
        want = oracle.gmag_os(a32.astype(np.float64), n32.astype(np.float64), None, dxa)
        assert np.allclose(got, want, rtol=2e-6, atol=1e-6)


def test_host_pointer_dropin_at_volume_size(oracle):
    """The six reference symbols on HOST buffers at a size where the
    staging copies matter: 160 x 144 x 192 float64 (35 MB per array), banded and full masks."""
    from lsml_b200.gradient import masked_gradient as mg
    rs = np.random.RandomState(21)
    shape = (160, 144, 192)
    arr = rs.randn(*shape)
    nu = rs.randn(*shape)
    dx = np.array([1.0, 0.5, 2.0])
    for mask in (None, np.abs(arr) < 0.1):
        got = mg.gradient_magnitude_osher_sethian(arr, nu, mask=mask, dx=dx)
        want = oracle.gmag_os(arr, nu, mask, dx, use_reference=oracle.reference_lib() is not None)
        assert np.array_equal(got, want)
    grads, gmag = mg.gradient_centered(arr, mask=np.abs(arr) < 0.5, dx=dx, normalize=True)
    wg, wm = oracle.gradient_centered(arr, np.abs(arr) < 0.5, dx, True,
                                      use_reference=oracle.reference_lib() is not None)
    assert np.array_equal(gmag, wm)
    for a in range(3):
        assert np.array_equal(grads[a], wg[a])
