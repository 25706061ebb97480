"""Helpers shared by the GPU parity test modules (test infrastructure)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _slab(eng, S_np, dtype):
    """Upload a host [(N+1), M] array into a padded step-major slab; returns the [:, :M] view."""
    N1, M = S_np.shape
    slab = eng.alloc_slab(M, N1 - 1, "f64" if dtype == torch.float64 else "f32")
    slab[:, :M] = _dev(S_np, dtype)
    return slab[:, :M]


def _check_sweep(res, ref, price_rtol, beta_rtol=1e-6, boundary_rtol=0.0):
    assert res.price == pytest.approx(ref.price, rel=price_rtol)
    # one-pass variance: absolute floor for the degenerate "all cash-flows equal" case
    assert res.stderr == pytest.approx(ref.stderr, rel=max(price_rtol, 1e-9), abs=1e-7 * max(1.0, abs(ref.price)))
    np.testing.assert_array_equal(res.n_itm, ref.n_itm)
    np.testing.assert_array_equal(res.ex_count, ref.ex_count)
    np.testing.assert_array_equal(np.isnan(res.boundary), np.isnan(ref.boundary))
    # the boundary is one of the input path values: exact when kernel and oracle saw the same array
    np.testing.assert_allclose(np.nan_to_num(res.boundary), np.nan_to_num(ref.boundary), rtol=boundary_rtol, atol=0)
    np.testing.assert_array_equal(np.isnan(res.betas), np.isnan(ref.betas))
    # betas are ill-conditioned individually; compare the fitted continuation over the ITM range instead
    x = np.linspace(0.8, 1.0, 9)
    for t in range(res.betas.shape[0]):
        if not np.isnan(ref.betas[t, 0]):
            p = ref.betas.shape[1]
            fit_ref = sum(ref.betas[t, i] * x**i for i in range(p))
            fit_gpu = sum(res.betas[t, i] * x**i for i in range(p))
            np.testing.assert_allclose(fit_gpu, fit_ref, rtol=beta_rtol, atol=beta_rtol)


def _oracle_price_philox(eng, mods, model_fn, S0, K, T, N, M, ot, seed, stream, dtype="f64", semantics="reference"):
    """Oracle sweep on the kernel's own Philox normals for one option."""
    L, E, orc = mods
    rng = E.RngSpec(seed=seed, stream=stream)
    z1 = eng.philox_normals(L.MODEL_HESTON, M, N, 0, dtype, rng).double().cpu().numpy()
    z2 = eng.philox_normals(L.MODEL_HESTON, M, N, 1, dtype, rng).double().cpu().numpy()
    S = orc.heston_paths_antithetic(S0, 0.05, T, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, z1, z2)
    return orc.lsm_sweep(S, K, 0.05, T, ot, semantics=semantics)
