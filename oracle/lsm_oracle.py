"""numpy restatement of the reference's path simulation + LSM backward induction.

TEST INFRASTRUCTURE ONLY (checker, CPU baseline).  Never imported by the product.

Parity status
-------------
The reference (Levicoz/Options-model) ships NO golden vectors or known-answer tests
for this path.  This restatement is therefore pinned against
outputs of the reference's own functions, generated in the build container by
``oracle/gen_golden.py`` (which imports the real reference from /root/reference with
its I/O-only dependencies stubbed) and committed under ``tests/golden/``:

* path schemes, features, Welford, RNG seed tree, European streaming, calibrator
  scheme: pinned bit-for-bit / to 1e-15 against the real reference functions;
* the LSM loop skeleton (discount order, sticky mask, strict ``>``, N-1 discounts,
  global normalisation): pinned against the real ``price_american_enhanced_lsm`` run
  with its network class swapped for a deterministic stand-in;
* the network pricers end to end -- global ``SingleLSMNet`` (om3:439-651) and per-date ``ContNet``
  (om2:216-330): ``price_american_enhanced_lsm_nn`` / ``price_american_om2_nn`` consume numpy's and
  torch's generators in the reference's order and reproduce prices of the REAL pricers bit for bit
  (``oracle/gen_golden_gnet.py`` -> ``tests/golden/ref_gnet_prices.json``);
* local volatility (om3:263-333): pinned against the real ``IVModel`` /
  ``simulate_local_vol_paths_antithetic`` (``oracle/gen_golden_localvol.py`` ->
  ``tests/golden/ref_localvol.npz``);
* the polynomial regressor itself does not exist in the reference (``lsm_poly_degree``
  is a dead parameter, options_model_2.py:176-180): its definition is SURVEY.md
  section 8(c) and it is pinned only by this file's own goldens -- PARITY UNPINNED for that
  regressor (the loop around it is pinned as above);
* Andersen QE and full truncation are north-star schemes absent from the reference: restated from
  the published algorithms, checked against the semi-analytic Heston price.

All ``file:line`` citations are relative to /root/reference.
om3 = options_model_3/options_model_3.py, om3gpu = options_model_3/option_model_3_gpu.py,
om2 = options_model_2.py, hc = options_model_3/heston_calibration.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np
from numpy.random import default_rng

# --------------------------------------------------------------------------------------
# RNG seed tree  (om3:69-79, identical in om3gpu:101-111)
# --------------------------------------------------------------------------------------


class RNGManager:
    """om3:69-79.  master = default_rng(seed); child seed = master.integers(0, 2**31-1)."""

    def __init__(self, master_seed: int = 42):
        self.master_rng = default_rng(master_seed)
        self.master_seed = master_seed

    def get_child_rng(self) -> np.random.Generator:
        return default_rng(self.master_rng.integers(0, 2**31 - 1))

    def get_child_seed(self) -> int:
        return self.master_rng.integers(0, 2**31 - 1)


# --------------------------------------------------------------------------------------
# Welford streaming estimator  (om3:33-63)
# --------------------------------------------------------------------------------------


def welford_batch_update(mean, m2, n, batch):
    """Chan merge of (mean, M2, n) with one batch.  om3:33-49."""
    batch = np.asarray(batch, dtype=np.float64)
    b_n = batch.size
    if b_n == 0:
        return mean, m2, n
    batch_mean = batch.mean()
    batch_m2 = ((batch - batch_mean) ** 2).sum()
    delta = batch_mean - mean
    new_n = n + b_n
    if new_n == 0:
        return mean, m2, n
    mean_new = mean + delta * (b_n / new_n)
    m2_new = m2 + batch_m2 + delta**2 * n * b_n / new_n
    return mean_new, m2_new, new_n


def monte_carlo_price_streaming(simulator_func, total_paths, chunk_size, *a, **kw):
    """om3:51-63."""
    n_done, mean, m2 = 0, 0.0, 0.0
    while n_done < total_paths:
        batch = min(chunk_size, total_paths - n_done)
        payoffs = simulator_func(batch, *a, **kw)
        mean, m2, n_done = welford_batch_update(mean, m2, n_done, payoffs)
    variance = m2 / (n_done - 1) if n_done > 1 else 0.0
    stderr = np.sqrt(variance / n_done) if n_done > 0 else 0.0
    return mean, stderr, n_done


# --------------------------------------------------------------------------------------
# Payoff and features
# --------------------------------------------------------------------------------------


def payoff(S, K, option_type):
    """om3:376-380."""
    if option_type == "call":
        return np.maximum(S - K, 0)
    return np.maximum(K - S, 0)


def features_ref7(S, K, r, T, t_current):
    """[1, x, x^2, x^3, max(x-1,0), sqrt(tau), x*sqrt(tau)].  om3:105-121."""
    x = S / K
    tau = T - t_current
    tau_sqrt = float(np.sqrt(max(tau, 1e-6)))
    tau_col = np.full_like(x, tau_sqrt, dtype=np.float64)
    return np.column_stack(
        [
            np.ones_like(x, dtype=np.float64),
            x.astype(np.float64),
            (x**2).astype(np.float64),
            (x**3).astype(np.float64),
            np.maximum(x - 1, 0).astype(np.float64),
            tau_col,
            (x * tau_col).astype(np.float64),
        ]
    )


# --------------------------------------------------------------------------------------
# Path simulation (CPU fp64 schemes)
# --------------------------------------------------------------------------------------


def draw_gbm_normals(rng, N, M):
    """One call standard_normal((N, M//2)).  om3:475."""
    return rng.standard_normal((N, M // 2))


def gbm_paths_antithetic(S0, r, sigma, T, M, N, Z_half):
    """Antithetic multiplicative log-Euler GBM, step-major S[(N+1), M].  om3:473-480.

    Columns [0, M/2) use +Z, columns [M/2, M) use -Z (om3:476).
    """
    dt = T / N
    drift = (r - 0.5 * sigma**2) * dt
    diffusion = sigma * np.sqrt(dt)
    Z = np.concatenate([Z_half, -Z_half], axis=1)
    S = np.zeros((N + 1, M), dtype=np.float64)
    S[0] = S0
    for t in range(1, N + 1):
        S[t] = S[t - 1] * np.exp(drift + diffusion * Z[t - 1])
    return S


def draw_heston_normals(rng, N, M):
    """Per step: z1_half then z2_half, each standard_normal(M//2).  om3:222-224."""
    Z1 = np.empty((N, M // 2))
    Z2 = np.empty((N, M // 2))
    for t in range(N):
        Z1[t] = rng.standard_normal(M // 2)
        Z2[t] = rng.standard_normal(M // 2)
    return Z1, Z2


def heston_paths_antithetic(S0, r, T, v0, kappa, theta, xi, rho, M, N, Z1_half, Z2_half, return_v=False):
    """Absorption-Euler Heston, antithetic, step-major.  om3:211-233 (even M only).

    v_prev = max(v,0); v' = max(v_prev + kappa(theta-v_prev)dt + xi sqrt(v_prev dt) w2, 0);
    S' = S exp((r - v_prev/2)dt + sqrt(v_prev dt) w1);  w1=z1, w2=rho z1 + sqrt(1-rho^2) z2.
    """
    dt = T / N
    S = np.zeros((N + 1, M), dtype=np.float64)
    v = np.zeros((N + 1, M), dtype=np.float64)
    S[0] = S0
    v[0] = v0
    for t in range(1, N + 1):
        z1 = np.concatenate([Z1_half[t - 1], -Z1_half[t - 1]])
        z2 = np.concatenate([Z2_half[t - 1], -Z2_half[t - 1]])
        w1 = z1
        w2 = rho * z1 + np.sqrt(1 - rho**2) * z2
        v_prev = np.maximum(v[t - 1], 0)
        v[t] = v_prev + kappa * (theta - v_prev) * dt + xi * np.sqrt(v_prev * dt) * w2
        v[t] = np.maximum(v[t], 0)
        S[t] = S[t - 1] * np.exp((r - 0.5 * v_prev) * dt + np.sqrt(v_prev * dt) * w1)
    return (S, v) if return_v else S


def heston_paths_full_truncation(S0, r, T, v0, kappa, theta, xi, rho, M, N, Z1_half, Z2_half):
    """Lord-Koekkoek-van Dijk full truncation (NOT in the reference; north-star scheme).

    The state v may go negative; only v+ = max(v,0) enters drift and diffusion.
    """
    dt = T / N
    S = np.zeros((N + 1, M), dtype=np.float64)
    S[0] = S0
    v = np.full(M, v0, dtype=np.float64)
    for t in range(1, N + 1):
        z1 = np.concatenate([Z1_half[t - 1], -Z1_half[t - 1]])
        z2 = np.concatenate([Z2_half[t - 1], -Z2_half[t - 1]])
        w2 = rho * z1 + np.sqrt(1 - rho**2) * z2
        vp = np.maximum(v, 0)
        sq = np.sqrt(vp * dt)
        S[t] = S[t - 1] * np.exp((r - 0.5 * vp) * dt + sq * z1)
        v = v + kappa * (theta - vp) * dt + xi * sq * w2
    return S


def _norm_cdf(z):
    from math import erfc
    return 0.5 * np.vectorize(erfc, otypes=[np.float64])(-z * 0.70710678118654752440)


def heston_paths_qe(S0, r, T, v0, kappa, theta, xi, rho, M, N, Z1_half, Z2_half, return_v=False):
    """Andersen's quadratic-exponential scheme (NOT in the reference; north-star scheme).

    L. Andersen, "Simple and efficient simulation of the Heston stochastic volatility model", J. Comp. Finance
    11(3), 2008: variance by moment matching (sec. 3.2: squared Gaussian for psi <= 1.5, mass at zero + exponential
    tail otherwise), log-asset by the central discretisation of sec. 4.2 with gamma1 = gamma2 = 1/2, no martingale
    correction.  z1 drives the asset, z2 the variance; the exponential branch uses U = Phi(z2).
    """
    dt = T / N
    E = np.exp(-kappa * dt)
    c1 = xi * xi * E * (1 - E) / kappa
    c2 = theta * xi * xi * (1 - E) ** 2 / (2 * kappa)
    a_ = kappa * rho / xi - 0.5
    k0r = r * dt - rho * kappa * theta * dt / xi
    k1 = 0.5 * dt * a_ - rho / xi
    k2 = 0.5 * dt * a_ + rho / xi
    k3 = 0.5 * dt * (1 - rho * rho)
    S = np.zeros((N + 1, M), dtype=np.float64)
    V = np.zeros((N + 1, M), dtype=np.float64)
    S[0] = S0
    V[0] = v0
    v = np.full(M, v0, dtype=np.float64)
    for t in range(1, N + 1):
        z1 = np.concatenate([Z1_half[t - 1], -Z1_half[t - 1]])
        z2 = np.concatenate([Z2_half[t - 1], -Z2_half[t - 1]])
        m = theta * (1 - E) + v * E
        s2 = v * c1 + c2
        with np.errstate(divide="ignore", invalid="ignore"):
            psi = s2 / (m * m)
            ip = 2.0 / psi
            b2 = ip - 1 + np.sqrt(ip) * np.sqrt(np.maximum(ip - 1, 0))
            vq = m / (1 + b2) * (np.sqrt(b2) + z2) ** 2
            p = (psi - 1) / (psi + 1)
            beta = (1 - p) / m
            ve = np.where(_norm_cdf(z2) <= p, 0.0, np.log((1 - p) / _norm_cdf(-z2)) / beta)
        vn = np.where(m > 0, np.where(psi <= 1.5, vq, ve), 0.0)
        S[t] = S[t - 1] * np.exp(k0r + k1 * v + k2 * vn + np.sqrt(k3 * v + k3 * vn) * z1)
        v = vn
        V[t] = v
    return (S, V) if return_v else S


def heston_european_analytic(S0, K, r, T, v0, kappa, theta, xi, rho, option_type="call", n=4000, umax=200.0):
    """Semi-analytic Heston price (characteristic function in the 'little trap' form of Albrecher et al. 2007,
    Gil-Pelaez inversion, composite midpoint rule).  Test infrastructure for the non-parity schemes."""
    u = (np.arange(n) + 0.5) * (umax / n)
    x = np.log(S0)

    def cf(w):
        d = np.sqrt((rho * xi * 1j * w - kappa) ** 2 + xi * xi * (1j * w + w * w))
        g = (kappa - rho * xi * 1j * w - d) / (kappa - rho * xi * 1j * w + d)
        ed = np.exp(-d * T)
        C = r * 1j * w * T + kappa * theta / xi**2 * ((kappa - rho * xi * 1j * w - d) * T - 2 * np.log((1 - g * ed) / (1 - g)))
        D = (kappa - rho * xi * 1j * w - d) / xi**2 * (1 - ed) / (1 - g * ed)
        return np.exp(C + D * v0 + 1j * w * x)

    k = np.log(K)
    p1 = 0.5 + np.sum(np.real(np.exp(-1j * u * k) * cf(u - 1j) / (1j * u * cf(-1j)))) * (umax / n) / np.pi
    p2 = 0.5 + np.sum(np.real(np.exp(-1j * u * k) * cf(u) / (1j * u))) * (umax / n) / np.pi
    call = S0 * p1 - K * np.exp(-r * T) * p2
    return call if option_type == "call" else call - S0 + K * np.exp(-r * T)


# --------------------------------------------------------------------------------------
# Local volatility: the IV network inside the step (om3:263-333; network nniv:109-155)
# --------------------------------------------------------------------------------------
def _erf32(x):
    from math import erf
    return np.vectorize(erf, otypes=[np.float64])(x.astype(np.float64)).astype(np.float32)


def ivnet_forward(net, X):
    """ImprovedIVNetwork.forward in eval mode (nniv:146-155), fp32: h = GELU(W_in x + b_in); L times
    h = h + GELU(LayerNorm(W h + b)) (dropout is the identity); out = clamp(W_out h + b_out, min=epsilon).
    ``net`` = dict(hidden, layers, weights (state_dict order, flat fp32), m_scale, tau_scale, epsilon)."""
    H, L = int(net["hidden"]), int(net["layers"])
    w = np.asarray(net["weights"], dtype=np.float32)
    gelu = lambda v: (np.float32(0.5) * v * (np.float32(1) + _erf32(v * np.float32(0.70710678118654752440)))).astype(np.float32)  # noqa: E731
    o = 0
    W_in = w[o:o + 2 * H].reshape(H, 2); o += 2 * H
    b_in = w[o:o + H]; o += H
    h = gelu((X.astype(np.float32) @ W_in.T + b_in).astype(np.float32))
    for _ in range(L):
        W = w[o:o + H * H].reshape(H, H); o += H * H
        b = w[o:o + H]; o += H
        g = w[o:o + H]; o += H
        be = w[o:o + H]; o += H
        z = (h @ W.T + b).astype(np.float32)
        mu = z.mean(axis=1, keepdims=True, dtype=np.float32)
        var = ((z - mu) ** 2).mean(axis=1, keepdims=True, dtype=np.float32)
        zn = ((z - mu) / np.sqrt(var + np.float32(1e-5)) * g + be).astype(np.float32)
        h = (h + gelu(zn)).astype(np.float32)
    w_out = w[o:o + H]; o += H
    out = (h @ w_out + w[o]).astype(np.float32)
    return np.maximum(out, np.float32(net["epsilon"]))


def ivnet_sigma(net, K, S_batch, tau):
    """IVModel.get_volatility_batch (om3:277-298): moneyness scaled by the scaler's std but NOT centred (App. B-7)."""
    tau = max(float(tau), 1e-6)
    S_batch = np.asarray(S_batch, dtype=np.float64)
    m = np.log(np.maximum(K, 1e-8) / np.maximum(S_batch, 1e-8))
    X = np.column_stack([m / net["m_scale"], np.full_like(m, tau / net["tau_scale"])])
    return np.maximum(ivnet_forward(net, X.astype(np.float32)), np.float32(1e-6)).astype(np.float64)


def localvol_paths_antithetic(S0, r, T, M, N, net, K, Z_half):
    """simulate_local_vol_paths_antithetic (om3:300-320, even M): Z = [Z_half, -Z_half]."""
    dt = T / N
    S = np.zeros((N + 1, M), dtype=np.float64)
    S[0] = S0
    Z = np.concatenate([Z_half, -Z_half], axis=1)
    for t in range(1, N + 1):
        tau_t = max(T - (t - 1) * dt, 1e-6)
        sig = ivnet_sigma(net, K, S[t - 1], tau_t)
        S[t] = S[t - 1] * np.exp((r - 0.5 * sig**2) * dt + sig * np.sqrt(dt) * Z[t - 1])
    return S


# --------------------------------------------------------------------------------------
# Path simulation (torch fp32 variants, restated in numpy float32)
# --------------------------------------------------------------------------------------


def bs_paths_fp32(S0, r, T, sigma, M, N, Z_half32):
    """om3gpu:117-138 in float32 arithmetic (even M)."""
    f = np.float32
    dt = T / N
    drift = f((r - 0.5 * sigma**2) * dt)
    diffusion = f(sigma) * np.sqrt(f(dt))  # sigma * torch.sqrt(torch.tensor(dt))
    Z = np.concatenate([Z_half32, -Z_half32], axis=1).astype(f)
    S = np.full((N + 1, M), f(S0), dtype=f)
    for t in range(1, N + 1):
        S[t] = S[t - 1] * np.exp(drift + diffusion * Z[t - 1])
    return S


def bs_paths_logspace_fp32(S0, r, T, sigma, M, N, Z32):
    """om3gpu:150-185 ("bandwidth optimized"): no antithetic, cumulative sum in log space."""
    f = np.float32
    dt = T / max(1, N)
    drift = f((r - 0.5 * sigma * sigma) * dt)
    diffusion = f(sigma * math.sqrt(dt))
    logS = np.empty((N + 1, M), dtype=f)
    logS[0] = f(math.log(max(S0, 1e-12)))
    inc = drift + diffusion * Z32.astype(f)
    for t in range(1, N + 1):
        logS[t] = logS[t - 1] + inc[t - 1]
    return np.exp(logS)


def heston_paths_fp32(S0, r, T, v0, kappa, theta, xi, rho, M, N, Z1_half32, Z2_half32):
    """om3gpu:187-225 in float32 arithmetic (even M)."""
    f = np.float32
    dt = f(T / N)
    kappa, theta, xi, rho = f(kappa), f(theta), f(xi), f(rho)
    sqrt_1_rho2 = np.sqrt(f(1) - rho * rho)
    S = np.full((N + 1, M), f(S0), dtype=f)
    v = np.full((N + 1, M), f(v0), dtype=f)
    rr = f(r)
    for t in range(1, N + 1):
        z1 = np.concatenate([Z1_half32[t - 1], -Z1_half32[t - 1]]).astype(f)
        z2 = np.concatenate([Z2_half32[t - 1], -Z2_half32[t - 1]]).astype(f)
        w2 = rho * z1 + sqrt_1_rho2 * z2
        v_prev = np.maximum(v[t - 1], f(0))
        sq = np.sqrt(v_prev * dt)
        v[t] = np.maximum(v_prev + kappa * (theta - v_prev) * dt + xi * sq * w2, f(0))
        S[t] = S[t - 1] * np.exp((rr - f(0.5) * v_prev) * dt + sq * z1)
    return S


# --------------------------------------------------------------------------------------
# Calibrator scheme (hc.HestonPricer)
# --------------------------------------------------------------------------------------


def hc_draw_normals(rng, n_paths, n_steps, antithetic=True):
    """Z1 then Z2_indep, each one standard_normal((n_sim, n_steps)) call.  hc:224-238."""
    n_sim = n_paths // 2 if antithetic else n_paths
    Z1 = rng.standard_normal((n_sim, n_steps))
    Z2i = rng.standard_normal((n_sim, n_steps))
    return Z1, Z2i


def hc_simulate_paths(kappa, theta, sigma, rho, v0, S0, T, r, n_paths, n_steps, Z1, Z2i, antithetic=True):
    """Arithmetic-Euler S, floored-variance Heston; path-major (n_paths, n_steps+1).  hc:204-257."""
    dt = T / n_steps
    S = np.zeros((n_paths, n_steps + 1))
    V = np.zeros((n_paths, n_steps + 1))
    S[:, 0] = S0
    V[:, 0] = v0
    Z2 = rho * Z1 + np.sqrt(1 - rho**2) * Z2i
    if antithetic:
        Z1 = np.vstack([Z1, -Z1])
        Z2 = np.vstack([Z2, -Z2])
    sqrt_dt = np.sqrt(dt)
    for t in range(n_steps):
        V_pos = np.maximum(V[:, t], 1e-8)
        sqrt_V = np.sqrt(V_pos)
        dV = kappa * (theta - V_pos) * dt + sigma * sqrt_V * sqrt_dt * Z2[:, t]
        V[:, t + 1] = np.maximum(V_pos + dV, 1e-8)
        dS = r * S[:, t] * dt + sqrt_V * S[:, t] * sqrt_dt * Z1[:, t]
        S[:, t + 1] = S[:, t] + dS
    return S, V


def hc_price_european(S_T, K, T, r, option_type="call"):
    """exp(-rT) * mean(payoff(S_T)).  hc:265-277."""
    if option_type.lower() == "call":
        pay = np.maximum(S_T - K, 0)
    else:
        pay = np.maximum(K - S_T, 0)
    return float(np.exp(-r * T) * np.mean(pay))


# --------------------------------------------------------------------------------------
# Regressors for the per-date LSM loop
# --------------------------------------------------------------------------------------

PIVOT_RTOL = 1e-14  # pivot < 1e-14 * trace(G) -> "no exercise at this date"


def hc_objective(x, S0, r, K, T, sigma_iv, n_mc_paths, n_time_steps, seed=42, use_vega_weighting=True, min_vega_weight=0.01):
    """``HestonCalibrator._objective_function`` (hc:404-472) on a fresh calibrator: one simulation per surface row from the
    pricer's persistent generator (hc:202, 226-227), vega-weighted RMSE of ln(P_heston / P_bs(sigma_market)) plus the
    Feller penalty.  Returns (objective, per-row Heston prices).  Reproduces the real reference bit for bit
    (tests/golden/ref_gnet_prices.json, key "hc_objective")."""
    from math import erf, exp, log, pi, sqrt

    kappa, theta, sigma, rho, v0 = (float(v) for v in x)
    rng = default_rng(seed)
    ncdf = lambda z: 0.5 * (1.0 + erf(z / sqrt(2.0)))  # noqa: E731
    prices, tot, wsum = [], 0.0, 0.0
    for k, t, iv in zip(K, T, sigma_iv):
        Z1, Z2i = hc_draw_normals(rng, n_mc_paths, n_time_steps)
        S, _ = hc_simulate_paths(kappa, theta, sigma, rho, v0, S0, t, r, n_mc_paths, n_time_steps, Z1, Z2i)
        p = hc_price_european(S[:, -1], k, t, r, "call")
        prices.append(p)
        if np.isnan(p) or p <= 1e-8:
            continue
        d1 = (log(S0 / k) + (r + 0.5 * iv**2) * t) / (iv * sqrt(t))
        d2 = d1 - iv * sqrt(t)
        bs = S0 * ncdf(d1) - k * exp(-r * t) * ncdf(d2)
        if bs <= 1e-8:
            continue
        vega = max(S0 * exp(-0.5 * d1 * d1) / sqrt(2.0 * pi) * sqrt(t), 1e-8)
        w = max(vega / 100.0, min_vega_weight) if use_vega_weighting else 1.0
        tot += w * log(p / bs) ** 2
        wsum += w
    if wsum == 0:
        return 1e6, prices
    penalty = 0.0 if 2 * kappa * theta >= sigma**2 else 100.0 * abs(2 * kappa * theta - sigma**2)
    return sqrt(tot / wsum) + penalty, prices


def basis_matrix(S_itm, K, basis, T=None, t_current=None, r=None):
    x = np.asarray(S_itm, dtype=np.float64) / K
    if basis == "poly2":
        return np.column_stack([np.ones_like(x), x, x * x])
    if basis == "poly3":
        return np.column_stack([np.ones_like(x), x, x * x, x * x * x])
    if basis == "ref7":
        return features_ref7(np.asarray(S_itm, dtype=np.float64), K, r, T, t_current)
    raise ValueError(basis)


def cholesky_solve_guarded(G, g, rtol=PIVOT_RTOL):
    """LDL^T (no pivoting) with the pivot guard of SURVEY.md 8(c).

    Returns beta, or None when a pivot d_k <= rtol * trace(G) (degenerate system).
    The CUDA kernel runs exactly this recurrence in fp64.
    """
    p = G.shape[0]
    tr = float(np.trace(G))
    L = np.eye(p)
    d = np.zeros(p)
    for k in range(p):
        s = G[k, k]
        for j in range(k):
            s -= L[k, j] * L[k, j] * d[j]
        if not (s > rtol * tr):
            return None
        d[k] = s
        for i in range(k + 1, p):
            u = G[i, k]
            for j in range(k):
                u -= L[i, j] * L[k, j] * d[j]
            L[i, k] = u / s
    # forward: L z = g ; diag: w = z/d ; back: L^T beta = w
    z = np.array(g, dtype=np.float64).copy()
    for i in range(p):
        for j in range(i):
            z[i] -= L[i, j] * z[j]
    w = z / d
    beta = w.copy()
    for i in range(p - 1, -1, -1):
        for j in range(i + 1, p):
            beta[i] -= L[j, i] * beta[j]
    return beta


class PolyRegressor:
    """Per-date least squares on basis columns; in-sample continuation."""

    def __init__(self, K, basis="poly2", T=None, r=None):
        self.K, self.basis, self.T, self.r = K, basis, T, r
        self.p = {"poly2": 3, "poly3": 4, "ref7": 7}[basis]

    def __call__(self, t, t_current, S_itm, Y):
        Phi = basis_matrix(S_itm, self.K, self.basis, self.T, t_current, self.r)
        n = Phi.shape[0]
        if n < self.p:
            return None, None
        G = Phi.T @ Phi
        g = Phi.T @ np.asarray(Y, dtype=np.float64)
        beta = cholesky_solve_guarded(G, g)
        if beta is None:
            return None, None
        return Phi @ beta, beta


class ContNetRegressor:
    """The reference's per-date network fit (om2:289-306 = om15:159-181 = om1:121-145): standardise the ITM
    prices (population std), train a fresh ContNet(1 -> H -> H -> 1, ReLU; om2:114-126) with `epochs` full-batch
    Adam steps on the raw discounted cash-flows, return its in-sample prediction.  Runs on torch CPU in fp32 like
    the reference.  `init_fn(t)` supplies the flat initial parameters [w1[H] b1[H] W2[H][H] b2[H] w3[H] b3] of date
    t (the reference uses torch's default initialisation; parity tests pass the engine's Philox weights)."""

    p = 1

    def __init__(self, init_fn, hidden=32, epochs=10, lr=1e-3):
        self.init_fn, self.H, self.epochs, self.lr = init_fn, hidden, epochs, lr
        self.last_loss = {}

    def __call__(self, t, t_current, S_itm, Y):
        import torch
        from torch import nn, optim

        H = self.H
        X = np.asarray(S_itm, dtype=np.float64)
        Xs = (X - X.mean()) / X.std() if X.std() > 0 else X - X.mean()  # om2:289
        net = nn.Sequential(nn.Linear(1, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, 1))
        if self.init_fn is None:  # torch's default initialisation from the global generator, as `ContNet()` (om2:291)
            p0 = None
        else:
            p0 = np.asarray(self.init_fn(t), dtype=np.float32)
        if p0 is not None:
            with torch.no_grad():
                net[0].weight.copy_(torch.from_numpy(p0[0:H].reshape(H, 1)))
                net[0].bias.copy_(torch.from_numpy(p0[H:2 * H]))
                net[2].weight.copy_(torch.from_numpy(p0[2 * H:2 * H + H * H].reshape(H, H)))
                net[2].bias.copy_(torch.from_numpy(p0[2 * H + H * H:3 * H + H * H]))
                net[4].weight.copy_(torch.from_numpy(p0[3 * H + H * H:4 * H + H * H].reshape(1, H)))
                net[4].bias.copy_(torch.from_numpy(p0[4 * H + H * H:4 * H + H * H + 1]))
        opt = optim.Adam(net.parameters(), lr=self.lr)
        Xt = torch.from_numpy(Xs.reshape(-1, 1)).float()
        Yt = torch.from_numpy(np.asarray(Y, dtype=np.float64).reshape(-1, 1)).float()
        for _ in range(self.epochs):
            loss = nn.MSELoss()(net(Xt), Yt)
            opt.zero_grad()
            loss.backward()
            opt.step()
            self.last_loss[t] = float(loss.detach())
        with torch.no_grad():
            cont = net(Xt).numpy().flatten()
        return cont.astype(np.float64), None


class SingleLSMNetDateRegressor:
    """Per-date fit of om3's regressor -- the loop of om2:277-310 (= om15:145-186) with `SingleLSMNet(7, hidden, 3)`
    (om3:85-103) in place of ContNet (BASELINE config 3, "tensor-core fit per date"; the reference ships the two pieces,
    not this combination): at date t the seven reference features (om3:105-121) of the live rows are z-scored with the
    date's own moments (std == 0 -> 1, om3:561), the current cash-flows with their mean / std (om3:550-556; ddof as
    `single_lsm_net_fit`'s variant), a FRESH network is trained by :func:`single_lsm_net_fit` (om3:565-613 /
    om3gpu:740-798) and predicts in-sample (dropout active iff ``inference_dropout``).  Streams: torch's RNG seeded with
    ``seed + t`` -- comparable with the engine statistically."""

    p = 1

    def __init__(self, K, r, T, variant="gpu", hidden=128, epochs=10, lr=1e-3, batch=None, dropout=0.1, seed=0,
                 inference_dropout=False, target_ddof=None, patience=None):
        self.K, self.r, self.T = K, r, T
        self.kw = dict(variant=variant, hidden=hidden, epochs=epochs, lr=lr, batch=batch, dropout=dropout,
                       inference_dropout=inference_dropout, patience=patience)
        self.seed = seed
        self.ddof = (0 if variant == "cpu" else 1) if target_ddof is None else target_ddof
        self.loss = {}

    def __call__(self, t, t_current, S_itm, Y):
        F = features_ref7(np.asarray(S_itm, dtype=np.float64), self.K, self.r, self.T, t_current)
        f_mean, f_std = F.mean(axis=0), F.std(axis=0)
        f_std[f_std == 0] = 1
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        y_mean = Y.mean()
        y_std = Y.std(ddof=self.ddof) if Y.size > self.ddof else 0.0
        if not y_std > 0:
            y_std = 1.0
        log = []
        Xn = (F - f_mean) / f_std
        predict = single_lsm_net_fit(seed=self.seed + t, log=log, **self.kw)(Xn, (Y - y_mean) / y_std)
        self.loss[t] = min(log) if log else float("nan")
        return np.asarray(predict(Xn)).reshape(-1) * y_std + y_mean, None


class FixedPolicyRegressor:
    """Out-of-sample exercise: the continuation value comes from coefficients fitted on OTHER paths
    (``betas[t]`` as returned by :func:`lsm_sweep`; a NaN row means "no exercise at that date").  Plugged into
    :func:`lsm_sweep` it prices that fixed policy with the same loop semantics."""

    def __init__(self, K, betas):
        self.K, self.betas, self.p = float(K), np.asarray(betas, dtype=np.float64), np.asarray(betas).shape[1]

    def __call__(self, t, t_current, S_itm, Y):
        b = self.betas[t]
        if np.isnan(b).any():
            return None, None
        x = np.asarray(S_itm, dtype=np.float64) / self.K
        return sum(b[i] * x**i for i in range(self.p)), b


@dataclass
class SweepResult:
    price: float
    stderr: float
    betas: np.ndarray  # [N+1, p], NaN rows where no regression happened
    boundary: np.ndarray  # [N+1], put: max exercised S; call: min exercised S; NaN if none
    ex_count: np.ndarray  # [N+1] int64, paths newly exercised at date t
    n_itm: np.ndarray  # [N+1] int64, regression rows at date t
    cashflows: Optional[np.ndarray] = field(default=None, repr=False)
    exercised: Optional[np.ndarray] = field(default=None, repr=False)


def lsm_sweep(S, K, r, T, option_type="put", regressor: Optional[Callable] = None, basis="poly2",
              semantics="reference", keep_state=False) -> SweepResult:
    """The reference's backward loop with a pluggable per-date regressor.

    semantics="reference" follows om3:616-651 (= om3:485-500, om2:278-310):
      cf = payoff(S[N]); for t = N-1..1: cf *= exp(-r dt) (ALL paths, before the mask, om3:620);
      itm = (payoff(S[t]) > 0) & ~exercised (om3:621); skip if none (om3:623);
      exercise where payoff > continuation, strict (om3:643-644); cf[ex] = payoff, exercised[ex] = True
      (sticky, om3:646-649); price = mean(cf) after N-1 discounts (om3:651).
    semantics="textbook": no sticky mask (a later-date exercise is overwritten by an earlier one),
      and one more discount so the value is at time 0.
    """
    S = np.asarray(S)
    N = S.shape[0] - 1
    M = S.shape[1]
    dt = T / N
    discount = np.exp(-r * dt)
    if regressor is None:
        regressor = PolyRegressor(K, basis, T=T, r=r)
    p = getattr(regressor, "p", 1)
    sticky = semantics == "reference"

    cf = payoff(S[-1], K, option_type).astype(np.float64)
    exercised = np.zeros(M, dtype=bool)
    betas = np.full((N + 1, p), np.nan)
    boundary = np.full(N + 1, np.nan)
    ex_count = np.zeros(N + 1, dtype=np.int64)
    n_itm = np.zeros(N + 1, dtype=np.int64)

    for t in range(N - 1, 0, -1):
        cf *= discount
        pay_t = payoff(S[t], K, option_type)
        itm = pay_t > 0
        if sticky:
            itm &= ~exercised
        n_itm[t] = int(itm.sum())
        if not np.any(itm):
            continue
        X = S[t, itm]
        cont, beta = regressor(t, t * dt, X, cf[itm])
        if cont is None:
            continue
        if beta is not None:
            betas[t, : len(beta)] = beta
        immediate = pay_t[itm]
        to_ex = immediate > cont
        idx = np.where(itm)[0][to_ex]
        cf[idx] = immediate[to_ex]
        exercised[idx] = True
        ex_count[t] = idx.size
        if idx.size:
            boundary[t] = X[to_ex].max() if option_type == "put" else X[to_ex].min()

    scale = 1.0 if sticky else discount
    price = float(cf.mean() * scale)
    stderr = float(cf.std(ddof=1) / math.sqrt(M) * scale) if M > 1 else 0.0
    return SweepResult(price, stderr, betas, boundary, ex_count, n_itm,
                       cf * scale if keep_state else None, exercised if keep_state else None)


# --------------------------------------------------------------------------------------
# v3 "global regressor" structure (om3:482-651) with a pluggable fit
# --------------------------------------------------------------------------------------


def lsm_global(S, K, r, T, option_type, fit: Callable, target_ddof=0):
    """om3:482-651 with the network replaced by ``fit``.

    Pass 1 (om3:485-516): exercised is never set, so the targets are discounted TERMINAL payoffs of
    every ITM path at every date.  Normalisation om3:550-563 (population std; om3gpu:731 uses the
    sample std -> target_ddof=1).  ``fit(Xn, Ys) -> predict(Xn) -> Ys_hat``.  Pass 2 (om3:615-651)
    applies the sticky-mask loop with continuation = predict(norm(features)) * Y_std + Y_mean.
    """
    S = np.asarray(S, dtype=np.float64)
    N = S.shape[0] - 1
    M = S.shape[1]
    dt = T / N
    discount = np.exp(-r * dt)
    cf = payoff(S[-1], K, option_type).astype(np.float64)
    feats, targs = [], []
    for t in range(N - 1, 0, -1):
        cf *= discount
        itm = payoff(S[t], K, option_type) > 0
        if not np.any(itm):
            continue
        feats.append(features_ref7(S[t, itm], K, r, T, t * dt))
        targs.append(cf[itm].reshape(-1, 1))
    if not feats:
        return float(cf.mean()), None
    X_all = np.vstack(feats)
    Y_all = np.vstack(targs)
    Y_mean = Y_all.mean()
    Y_std = Y_all.std(ddof=target_ddof)
    if Y_std > 0:
        Ys = (Y_all - Y_mean) / Y_std
    else:
        Ys = Y_all - Y_mean
        Y_std = 1.0
    f_mean = X_all.mean(axis=0)
    f_std = X_all.std(axis=0)
    f_std[f_std == 0] = 1
    Xn = (X_all - f_mean) / f_std
    predict = fit(Xn, Ys)

    cf = payoff(S[-1], K, option_type).astype(np.float64)
    exercised = np.zeros(M, dtype=bool)
    ex_count = np.zeros(N + 1, dtype=np.int64)
    boundary = np.full(N + 1, np.nan)
    for t in range(N - 1, 0, -1):
        cf *= discount
        itm = (payoff(S[t], K, option_type) > 0) & (~exercised)
        if not np.any(itm):
            continue
        X = S[t, itm]
        fn = (features_ref7(X, K, r, T, t * dt) - f_mean) / f_std
        cont = np.asarray(predict(fn)).reshape(-1) * Y_std + Y_mean
        immediate = payoff(X, K, option_type)
        to_ex = immediate > cont
        idx = np.where(itm)[0][to_ex]
        cf[idx] = immediate[to_ex]
        exercised[idx] = True
        ex_count[t] = idx.size
        if idx.size:
            boundary[t] = X[to_ex].max() if option_type == "put" else X[to_ex].min()
    stats = dict(Y_mean=float(Y_mean), Y_std=float(Y_std), f_mean=f_mean, f_std=f_std, n_rows=int(X_all.shape[0]),
                 ex_count=ex_count, boundary=boundary, stderr=float(cf.std(ddof=1) / math.sqrt(M)) if M > 1 else 0.0)
    return float(cf.mean()), stats


def single_lsm_net_fit(variant="cpu", hidden=128, epochs=25, lr=1e-3, batch=None, dropout=0.1, seed=0,
                       inference_dropout=True, log=None, reference_streams=False, patience=None):
    """``fit`` for :func:`lsm_global` restating the reference's global network regression in torch (CPU, fp32).

    variant "cpu" = om3:565-613: SingleLSMNet(7, hidden, 3) (om3:85-103), DataLoader(batch 256, shuffle), Adam(lr,
    weight_decay 1e-5), ReduceLROnPlateau(patience 5, factor 0.5, min_lr 1e-6) on the mean batch loss, best-weights
    snapshot on an improvement > 1e-6, early stop after 8 epochs without one, best weights restored.
    variant "gpu" = om3gpu:740-798: batch min(8192, n), AdamW(weight_decay 1e-4), no scheduler, patience 3.
    The returned predictor keeps dropout ACTIVE when ``inference_dropout`` (the reference never calls net.eval(),
    SURVEY App. A).  Streams come from torch's RNG seeded with ``seed``: comparable with the engine statistically.
    ``reference_streams=True`` (variant "cpu") consumes torch's GLOBAL generator in exactly the reference's order --
    no re-seeding here (the caller did ``torch.manual_seed(rng_manager.get_child_seed())``, om3:455), network built
    like ``SingleLSMNet.__init__``, shuffling through ``DataLoader(shuffle=True)`` (om3:577) -- so that the real
    ``price_american_enhanced_lsm`` is reproduced bit for bit (tests/golden/ref_gnet_prices.json).
    """
    import copy

    import torch
    from torch import nn, optim

    def fit(Xn, Ys):
        if not reference_streams:
            torch.manual_seed(seed)
        layers = [nn.Linear(Xn.shape[1], hidden), nn.ReLU(), nn.Dropout(dropout)]
        for _ in range(2):
            layers += [nn.Linear(hidden, hidden), nn.ReLU(), nn.Dropout(dropout)]
        layers += [nn.Linear(hidden, 1)]
        net = nn.Sequential(*layers)
        X = torch.from_numpy(np.asarray(Xn)).float()
        Y = torch.from_numpy(np.asarray(Ys)).float().reshape(-1, 1)
        n = X.shape[0]
        cpu = variant == "cpu"
        B = int(batch if batch is not None else (min(256, n) if cpu else min(8192, n)))
        opt = optim.Adam(net.parameters(), lr=lr, weight_decay=1e-5) if cpu else optim.AdamW(net.parameters(), lr=lr, weight_decay=1e-4)
        sched = optim.lr_scheduler.ReduceLROnPlateau(opt, patience=5, factor=0.5, min_lr=1e-6) if cpu else None
        best, best_sd, bad = float("inf"), None, 0
        loader = None
        if reference_streams:
            from torch.utils.data import DataLoader, TensorDataset

            loader = DataLoader(TensorDataset(X, Y), batch_size=min(256, n), shuffle=True)  # om3:573-577
        for ep in range(epochs):
            tot, nb = 0.0, 0
            if loader is not None:
                batches = loader
            else:
                perm = torch.randperm(n)
                batches = ((X[perm[b0:b0 + B]], Y[perm[b0:b0 + B]]) for b0 in range(0, n, B))
            for bx, by in batches:
                loss = nn.MSELoss()(net(bx), by)
                opt.zero_grad()
                loss.backward()
                opt.step()
                tot += float(loss.detach())
                nb += 1
            avg = tot / nb
            if sched is not None:
                sched.step(avg)
            if log is not None:
                log.append(avg)
            if avg < best - 1e-6:
                best, best_sd, bad = avg, copy.deepcopy(net.state_dict()), 0
            else:
                bad += 1
                if bad >= ((8 if cpu else 3) if patience is None else patience):  # om3:605 / om3gpu:786
                    break
        if best_sd is not None:
            net.load_state_dict(best_sd)
        net.train(bool(inference_dropout))

        def predict(fn):
            with torch.no_grad():
                return net(torch.from_numpy(np.asarray(fn)).float()).numpy().reshape(-1).astype(np.float64)

        return predict

    return fit


def price_american_om2_nn(S0, K, r, sigma, T, option_type, num_simulations, num_time_steps, seed=42, nn_hidden=32, nn_epochs=10,
                          nn_lr=1e-3):
    """``om2.OptionPricer(...).price_american_option`` (om2:216-310, GBM), stream for stream: ``np.random.seed`` /
    ``torch.manual_seed`` (om2:241-242), one ``standard_normal((N, M/2))`` call, a fresh default-initialised ContNet per
    date.  -> (mean, population std, P(cash-flow == 0)).  Reproduces the real reference bit for bit on the same numpy /
    torch versions (tests/golden/ref_gnet_prices.json)."""
    import torch

    torch.manual_seed(seed)
    M, N = num_simulations // 2 * 2, num_time_steps
    Z = np.random.RandomState(seed).standard_normal((N, M // 2))
    S = gbm_paths_antithetic(S0, r, sigma, T, M, N, Z)
    res = lsm_sweep(S, K, r, T, option_type, ContNetRegressor(None, nn_hidden, nn_epochs, nn_lr), semantics="reference",
                    keep_state=True)
    cf = res.cashflows
    return float(cf.mean()), float(cf.std()), float(np.mean(cf == 0))


def price_american_enhanced_lsm_nn(S0, K, r, T, option_type, num_simulations, num_time_steps, master_seed, sigma=None,
                                   heston_params=None, nn_hidden=128, nn_epochs=25, nn_lr=1e-3):
    """``AdvancedOptionPricer(...).price_american_enhanced_lsm`` (om3:439-651) with its own network, stream for stream:
    child generator + ``torch.manual_seed`` from the RNGManager (om3:454-455), numpy paths, global SingleLSMNet fit
    and the decision pass with torch's global generator consumed in the reference's order.  Reproduces the real
    reference bit for bit on the same numpy / torch versions (tests/golden/ref_gnet_prices.json)."""
    import torch

    m = RNGManager(master_seed)
    rng = m.get_child_rng()
    torch.manual_seed(int(m.get_child_seed()))
    M, N = num_simulations // 2 * 2, num_time_steps
    if heston_params is not None:
        hp = heston_params
        Z1, Z2 = draw_heston_normals(rng, N, M)
        S = heston_paths_antithetic(S0, r, T, hp["v0"], hp["kappa"], hp["theta"], hp["xi"], hp["rho"], M, N, Z1, Z2)
    else:
        S = gbm_paths_antithetic(S0, r, sigma, T, M, N, draw_gbm_normals(rng, N, M))
    fit = single_lsm_net_fit("cpu", hidden=nn_hidden, epochs=nn_epochs, lr=nn_lr, reference_streams=True, inference_dropout=True)
    price, stats = lsm_global(S, K, r, T, option_type, fit, target_ddof=0)
    return price, stats


def linear_fit(Xn, Ys):
    """Least squares with intercept on the (z-scored) reference features: the regressor optmc_lsm_global
    implements.  Minimum-norm solution, so all-zero / duplicated columns are harmless."""
    A = np.column_stack([np.ones(len(Xn)), np.asarray(Xn, dtype=np.float64)])
    w, *_ = np.linalg.lstsq(A, np.asarray(Ys, dtype=np.float64).ravel(), rcond=None)
    return lambda fn: np.column_stack([np.ones(len(fn)), np.asarray(fn, dtype=np.float64)]) @ w


# --------------------------------------------------------------------------------------
# European estimators
# --------------------------------------------------------------------------------------


def price_european_streaming(K, r, sigma, option_type, rng_manager, S0, T, num_simulations=10000,
                             num_time_steps=50, chunk_size=500, heston_params=None):
    """om3:382-437: per chunk a fresh child rng, a full path array, payoff of S[-1], Welford merge."""
    discount_factor = np.exp(-r * T)

    def simulator(batch_size):
        rng = rng_manager.get_child_rng()
        N = num_time_steps
        if heston_params is not None:
            M = batch_size // 2 * 2
            Z1, Z2 = draw_heston_normals(rng, N, M)
            hp = heston_params
            S = heston_paths_antithetic(S0, r, T, hp["v0"], hp["kappa"], hp["theta"], hp["xi"], hp["rho"], M, N, Z1, Z2)
            if batch_size % 2 != 0:  # om3:235-249
                dt = T / N
                S_odd = np.zeros((N + 1, 1)); v_odd = np.zeros((N + 1, 1))
                S_odd[0] = S0; v_odd[0] = hp["v0"]
                for t in range(1, N + 1):
                    z1 = rng.standard_normal(1); z2 = rng.standard_normal(1)
                    w2 = hp["rho"] * z1 + np.sqrt(1 - hp["rho"] ** 2) * z2
                    vp = np.maximum(v_odd[t - 1], 0)
                    v_odd[t] = np.maximum(vp + hp["kappa"] * (hp["theta"] - vp) * dt + hp["xi"] * np.sqrt(vp * dt) * w2, 0)
                    S_odd[t] = S_odd[t - 1] * np.exp((r - 0.5 * vp) * dt + np.sqrt(vp * dt) * z1)
                S = np.concatenate([S, S_odd], axis=1)
        else:
            M = batch_size // 2 * 2
            Zh = draw_gbm_normals(rng, N, M)
            S = gbm_paths_antithetic(S0, r, sigma, T, M, N, Zh)
            if batch_size % 2 != 0:  # om3:417-423
                dt = T / N
                drift = (r - 0.5 * sigma**2) * dt
                diffusion = sigma * np.sqrt(dt)
                S_odd = np.zeros((N + 1, 1)); S_odd[0] = S0
                Z_odd = rng.standard_normal((N, 1))
                for t in range(1, N + 1):
                    S_odd[t] = S_odd[t - 1] * np.exp(drift + diffusion * Z_odd[t - 1])
                S = np.concatenate([S, S_odd], axis=1)
        return (payoff(S[-1], K, option_type) * discount_factor).astype(np.float64)

    return monte_carlo_price_streaming(simulator, num_simulations, chunk_size)


def european_from_paths(S_T, K, r, T, option_type):
    """Discounted terminal payoff mean / stderr (a12) on a given terminal slab."""
    pay = payoff(np.asarray(S_T, dtype=np.float64), K, option_type) * np.exp(-r * T)
    n = pay.size
    return float(pay.mean()), float(pay.std(ddof=1) / math.sqrt(n)) if n > 1 else 0.0


# --------------------------------------------------------------------------------------
# End-to-end restatement of price_american_enhanced_lsm with the polynomial regressor
# --------------------------------------------------------------------------------------


def price_american_lsm(S0, K, r, T, option_type, num_simulations, num_time_steps, rng_manager, sigma=None,
                       heston_params=None, basis="poly2", semantics="reference", return_paths=False):
    """om3:439-480 (validation, RNG order, path model routing) + lsm_sweep."""
    if S0 <= 0 or K <= 0 or T <= 0:
        raise ValueError("S0, K, T must be positive.")
    if r < 0:
        raise ValueError("r must be non-negative.")
    if num_simulations <= 0 or num_time_steps <= 0:
        raise ValueError("num_simulations and num_time_steps must be positive integers.")
    rng = rng_manager.get_child_rng()
    rng_manager.get_child_seed()  # om3:455 consumes a second master draw (torch.manual_seed)
    M = num_simulations // 2 * 2
    N = num_time_steps
    if heston_params is not None:
        hp = heston_params
        Z1, Z2 = draw_heston_normals(rng, N, M)
        S = heston_paths_antithetic(S0, r, T, hp["v0"], hp["kappa"], hp["theta"], hp["xi"], hp["rho"], M, N, Z1, Z2)
        normals = (Z1, Z2)
    else:
        if sigma is None:
            raise ValueError("sigma is None: provide sigma, iv_model, or heston configuration")
        Zh = draw_gbm_normals(rng, N, M)
        S = gbm_paths_antithetic(S0, r, sigma, T, M, N, Zh)
        normals = (Zh,)
    res = lsm_sweep(S, K, r, T, option_type, basis=basis, semantics=semantics)
    if return_paths:
        return res, S, normals
    return res


# ----------------------------------------------------------------------------------------------------------------
# Quasi-Monte-Carlo draws: no counterpart in the reference (PCG64 normals, om3:223-224,475).  Restated
# from the published constructions: Sobol' points with Joe-Kuo (2008) direction numbers -- scipy.stats.qmc.Sobol ships
# the same table and serves as the independent implementation of the point set -- and the Brownian-bridge ordering of
# Jaeckel, "Monte Carlo Methods in Finance" (2002), sec. 10.8.3.
# ----------------------------------------------------------------------------------------------------------------
def brownian_bridge_schedule(N):
    """Construction order on the unit grid t = 1..N: step s sets W(idx+1) = wl W(left) + wr W(right+1) + sd z_s (W(0) = 0)."""
    idx, left, right = (np.zeros(N, dtype=np.int64) for _ in range(3))
    wl, wr, sd = (np.zeros(N) for _ in range(3))
    done = np.zeros(N, dtype=bool)
    done[N - 1] = True
    idx[0], right[0], sd[0] = N - 1, N - 1, np.sqrt(N)
    j = 0
    for i in range(1, N):
        while done[j]:
            j += 1
        k = j
        while not done[k]:
            k += 1
        l = j + ((k - 1 - j) >> 1)
        done[l] = True
        idx[i], left[i], right[i] = l, j, k
        wl[i], wr[i] = (k - l) / (k + 1 - j), (l + 1 - j) / (k + 1 - j)
        sd[i] = np.sqrt((l + 1 - j) * (k - l) / (k + 1 - j))
        j = k + 1
        if j >= N:
            j = 0
    return idx, left, right, wl, wr, sd


def sobol_bridge_normals(M, N, factors=1, bridge=True, shift=None, pair_offset=0):
    """Step-major normals [factors][N][M/2] of antithetic pairs pair_offset .. pair_offset + M/2 - 1: Sobol' point
    (pair index + 1) in factors * N dimensions, optional digital shift (uint32 per dimension), u = (x + 1/2) 2^-32,
    z = Phi^-1(u); with ``bridge`` the dimensions of a factor drive the Brownian-bridge construction and the normals
    are the increments of the constructed path."""
    from scipy.special import ndtri
    from scipy.stats import qmc

    D, Mh = factors * N, M // 2
    eng = qmc.Sobol(d=D, scramble=False, bits=32)
    eng.fast_forward(1 + int(pair_offset))
    x = np.rint(eng.random(Mh) * 4294967296.0).astype(np.uint64).astype(np.uint32)  # [Mh, D], exact multiples of 2^-32
    if shift is not None:
        x = x ^ np.asarray(shift, dtype=np.uint32)[None, :]
    z = ndtri((x.astype(np.float64) + 0.5) * 2.3283064365386963e-10)          # [Mh, D]
    out = np.empty((factors, N, Mh))
    sched = brownian_bridge_schedule(N) if bridge else None
    for f in range(factors):
        zf = z[:, f * N:(f + 1) * N]
        if not bridge:
            out[f] = zf.T
            continue
        idx, left, right, wl, wr, sd = sched
        W = np.zeros((N, Mh))
        W[idx[0]] = sd[0] * zf[:, 0]
        for s in range(1, N):
            lo = W[left[s] - 1] if left[s] > 0 else 0.0
            W[idx[s]] = wl[s] * lo + wr[s] * W[right[s]] + sd[s] * zf[:, s]
        out[f, 0] = W[0]
        out[f, 1:] = W[1:] - W[:-1]
    return out
