"""Drop-in mirror of the reference's Python call signatures for the hot path.

Same names, argument meaning and error behaviour as the reference; the bodies route to liboptmc.so.
om3 = options_model_3/options_model_3.py, om3gpu = options_model_3/option_model_3_gpu.py,
om2 = options_model_2.py, om1 = Options_model.py, hc = options_model_3/heston_calibration.py.

What differs from the reference, by design:
  * The continuation regressor defaults to the polynomial least-squares fit of SURVEY.md 8(c) (``lsm_regressor``
    keyword: "poly2" | "poly3"): deterministic and one persistent launch.  ``lsm_regressor="nn"`` runs the
    reference's own algorithm -- one SingleLSMNet(7, 128, 3) trained on the rows of all dates with the
    reference's optimiser settings, on the tensor cores (optmc_lsm_gnet); ``nn_epochs``, ``nn_lr``,
    ``nn_dropout`` then mean what they mean in the reference.
  * Random numbers: functions that receive an explicit numpy ``Generator`` / rely on torch's global
    generator (simulate_heston_paths_antithetic, simulate_*_torch, HestonPricer.simulate_paths) draw
    from it exactly as the reference does and feed the draws to the kernels, so results match the
    reference draw-for-draw.  The pricer methods use in-kernel Philox keyed by the RNGManager child
    seed; they consume the same number of master draws as the reference (om3:454-455, om3:392).
  * iv_model (local volatility, om3:263-333): ``IVModel`` flattens the reference's trained network once; the kernel
    evaluates it inside every path step (price_american_enhanced_lsm / simulate_local_vol_paths_antithetic).
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass, field, make_dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
from numpy.random import default_rng

from . import _lib as L
from . import engine as E


# ---------------------------------------------------------------------------------------------------
# RNG seed tree (om3:69-79) -- host integers only
# ---------------------------------------------------------------------------------------------------
class RNGManager:
    def __init__(self, master_seed: int = 42):
        self.master_rng = default_rng(master_seed)
        self.master_seed = master_seed

    def get_child_rng(self) -> np.random.Generator:
        child_seed = self.master_rng.integers(0, 2**31 - 1)
        return default_rng(child_seed)

    def get_child_seed(self) -> int:
        return self.master_rng.integers(0, 2**31 - 1)


# ---------------------------------------------------------------------------------------------------
# closed forms used by the control variate (om3:150-159) -- scalar host math, not on the hot path
# ---------------------------------------------------------------------------------------------------
def _ncdf(x: float) -> float:
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


class BlackScholesGreeks:
    @staticmethod
    def black_scholes_price(S, K, T, r, sigma, option_type="call"):
        d1 = (math.log(S / K) + (r + 0.5 * sigma**2) * T) / (sigma * math.sqrt(T))
        d2 = d1 - sigma * math.sqrt(T)
        if option_type == "call":
            return S * _ncdf(d1) - K * math.exp(-r * T) * _ncdf(d2)
        return K * math.exp(-r * T) * _ncdf(-d2) - S * _ncdf(-d1)


# ---------------------------------------------------------------------------------------------------
# free path-simulation functions
# ---------------------------------------------------------------------------------------------------
def _engine(device=0) -> E.Engine:
    return E.default_engine(device)


def _dev_index(device) -> int:
    import torch

    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type != "cuda":
        raise L.OptmcError(f"device {d} requested; options_model_b200 runs on CUDA devices only (no CPU fallback)")
    return d.index if d.index is not None else torch.cuda.current_device()


def simulate_heston_paths_antithetic(S0: float, r: float, T: float, v0: float, kappa: float, theta: float,
                                     xi: float, rho: float, num_simulations: int, num_time_steps: int,
                                     rng: np.random.Generator) -> np.ndarray:
    """om3:211-251.  Draws z1_half, z2_half per step from ``rng`` exactly like the reference (om3:223-224),
    runs the absorption-Euler kernel in fp64 on those draws, returns S[(N+1), num_simulations] (host)."""
    import torch

    eng = _engine()
    N = int(num_time_steps)
    M = num_simulations // 2 * 2
    out = np.zeros((N + 1, num_simulations), dtype=np.float64)
    model = E.heston(S0, r, T, v0, kappa, theta, xi, rho)
    if M > 0:
        Z1 = np.empty((N, M // 2)); Z2 = np.empty((N, M // 2))
        for t in range(N):
            Z1[t] = rng.standard_normal(M // 2)
            Z2[t] = rng.standard_normal(M // 2)
        z1 = torch.from_numpy(Z1).to(eng.tdev); z2 = torch.from_numpy(Z2).to(eng.tdev)
        S = eng.paths(model, M, N, "f64", E.RngSpec(z1=z1, z2=z2))
        out[:, :M] = S.cpu().numpy()
    if num_simulations % 2 != 0:  # om3:235-249: one extra, non-antithetic path
        Zo = np.empty((2, N, 1))
        for t in range(N):
            Zo[0, t] = rng.standard_normal(1)
            Zo[1, t] = rng.standard_normal(1)
        z = torch.from_numpy(Zo).to(eng.tdev)
        S = eng.paths(model, 1, N, "f64", E.RngSpec(z1=z[0].contiguous(), z2=z[1].contiguous(), antithetic=False))
        out[:, M:] = S.cpu().numpy()
    return out


class IVModel:
    """om3:263-298.  Wraps the reference's trained ``ImprovedIVNetwork`` (any torch module with the same state_dict
    layout and a fitted ``.scaler``); the weights are flattened once and the network is evaluated by the CUDA
    kernels -- inside the path step (``simulate_local_vol_paths_antithetic``) or for a batch of spots
    (``get_volatility_batch``)."""

    def __init__(self, nn_model):
        self.model = nn_model.eval()
        if hasattr(nn_model, "scaler") and nn_model.scaler is not None:
            self.m_scale = nn_model.scaler.m_scale
            self.tau_scale = nn_model.scaler.tau_scale
        else:
            raise ValueError("Model does not have a fitted scaler")
        sd = nn_model.state_dict()
        H = int(sd["input_proj.weight"].shape[0])
        layers = len({k.split(".")[1] for k in sd if k.startswith("layers.")})
        weights = np.concatenate([v.detach().cpu().numpy().reshape(-1) for v in sd.values()]).astype(np.float32)
        eps = float(getattr(getattr(nn_model, "config", None), "epsilon", 1e-4))
        self.net = dict(hidden=H, layers=layers, weights=weights, m_scale=float(self.m_scale), tau_scale=float(self.tau_scale),
                        epsilon=eps)

    def get_volatility_batch_torch(self, K: float, S_batch, tau: float):
        """om3gpu:498-519: tensor in, fp32 tensor out on the engine's device."""
        import torch

        if K <= 0:
            raise ValueError(f"K must be positive, got {K}")
        if torch.any(S_batch <= 0):
            raise ValueError("All S_batch values must be positive")
        return _engine().ivnet_sigma(self.net, K, S_batch.reshape(-1), tau).reshape(S_batch.shape).float()

    def get_volatility_batch(self, K: float, S_batch: np.ndarray, tau: float) -> np.ndarray:
        S_batch = np.asarray(S_batch, dtype=np.float64)
        if K <= 0:
            raise ValueError(f"K must be positive, got {K}")
        if np.any(S_batch <= 0):
            raise ValueError("All S_batch values must be positive")
        return _engine().ivnet_sigma(self.net, K, S_batch.ravel(), tau).cpu().numpy().reshape(S_batch.shape)


def simulate_local_vol_paths_antithetic(S0: float, r: float, T: float, num_simulations: int, num_time_steps: int,
                                        iv_model: "IVModel", K: float, rng: np.random.Generator) -> np.ndarray:
    """om3:300-333.  Draws Z_half from ``rng`` exactly like the reference (one (N, M/2) call, then (N, 1) for an odd
    path), runs the local-volatility kernel in fp64 on those draws, returns S[(N+1), num_simulations] (host)."""
    import torch

    eng = _engine()
    N = int(num_time_steps)
    M = num_simulations // 2 * 2
    out = np.zeros((N + 1, num_simulations), dtype=np.float64)
    out[0] = S0
    if M > 0:
        Zh = rng.standard_normal((N, M // 2))
        S = eng.paths_localvol(S0, r, T, iv_model.net, K, M, N, "f64", E.RngSpec(z1=torch.from_numpy(Zh).to(eng.tdev)))
        out[:, :M] = S.cpu().numpy()
    if num_simulations % 2 != 0:
        Zo = rng.standard_normal((N, 1))
        S = eng.paths_localvol(S0, r, T, iv_model.net, K, 1, N, "f64",
                               E.RngSpec(z1=torch.from_numpy(Zo).to(eng.tdev), antithetic=False))
        out[:, M:] = S.cpu().numpy()
    return out


def simulate_local_vol_paths_torch(S0: float, r: float, T: float, num_simulations: int, num_time_steps: int, iv_model: "IVModel",
                                   K: float, device, philox_seed: Optional[int] = None):
    """om3gpu:250-298.  fp32 paths on ``device``; Z_half = torch.randn(N, M//2, device=device) exactly as the reference
    draws it (torch.manual_seed reproduces its stream), or in-register Philox with ``philox_seed``.  Returns the
    [(N+1), num_simulations] tensor; an odd path count adds one non-antithetic column (om3gpu:283-296)."""
    import torch

    eng = _engine(_dev_index(device))
    N, M = int(num_time_steps), num_simulations // 2 * 2
    cols = []
    if M > 0:
        rng = E.RngSpec(seed=int(philox_seed)) if philox_seed is not None else \
            E.RngSpec(z1=torch.randn(N, M // 2, device=eng.tdev))
        cols.append(eng.paths_localvol(S0, r, T, iv_model.net, K, M, N, "f32", rng))
    if num_simulations % 2 != 0:
        rng = E.RngSpec(seed=int(philox_seed) + 1, antithetic=False) if philox_seed is not None else \
            E.RngSpec(z1=torch.randn(N, 1, device=eng.tdev), antithetic=False)
        cols.append(eng.paths_localvol(S0, r, T, iv_model.net, K, 1, N, "f32", rng))
    return torch.cat(cols, dim=1) if len(cols) > 1 else cols[0]


def simulate_bs_paths_torch(S0: float, r: float, T: float, sigma: float, num_simulations: int, num_time_steps: int,
                            device, philox_seed: Optional[int] = None):
    """om3gpu:117-148.  fp32, antithetic.  Default: Z_half = torch.randn(N, M//2, device=device) as the
    reference draws it (so torch.manual_seed reproduces the reference's paths); philox_seed=int switches
    to in-register Philox (nothing but S touches HBM)."""
    import torch

    eng = _engine(_dev_index(device))
    M = num_simulations // 2 * 2
    N = int(num_time_steps)
    model = E.gbm(S0, r, T, sigma)
    if philox_seed is not None:
        S = eng.paths(model, M, N, "f32", E.RngSpec(seed=philox_seed))
    else:
        Zh = torch.randn(N, M // 2, device=eng.tdev)
        S = eng.paths(model, M, N, "f32", E.RngSpec(z1=Zh))
    if num_simulations % 2 != 0:  # om3gpu:141-146
        Zo = torch.randn(N, 1, device=eng.tdev)
        So = eng.paths(model, 1, N, "f32", E.RngSpec(z1=Zo, antithetic=False))
        S = torch.cat([S, So], dim=1)
    return S


def simulate_bs_paths_torch_bandwidth_optimized(S0: float, r: float, T: float, sigma: float, num_simulations: int,
                                                num_time_steps: int, device, philox_seed: Optional[int] = None):
    """om3gpu:150-185: no antithetic, cumulative sum in log space, exp at the end (one launch here)."""
    import torch

    eng = _engine(_dev_index(device))
    M, N = int(num_simulations), int(num_time_steps)
    model = E.gbm(S0, r, T, sigma, scheme=L.SCHEME_GBM_LOGSPACE)
    if philox_seed is not None:
        return eng.paths(model, M, N, "f32", E.RngSpec(seed=philox_seed, antithetic=False))
    Z = torch.randn(N, M, device=eng.tdev, dtype=torch.float32)
    return eng.paths(model, M, N, "f32", E.RngSpec(z1=Z, antithetic=False))


def simulate_heston_paths_torch(S0: float, r: float, T: float, v0: float, kappa: float, theta: float, xi: float,
                                rho: float, num_simulations: int, num_time_steps: int, device,
                                philox_seed: Optional[int] = None):
    """om3gpu:187-248.  fp32 absorption Euler; per step z1_half then z2_half from torch.randn (om3gpu:209-210)."""
    import torch

    eng = _engine(_dev_index(device))
    M = num_simulations // 2 * 2
    N = int(num_time_steps)
    model = E.heston(S0, r, T, v0, kappa, theta, xi, rho)
    if philox_seed is not None:
        S = eng.paths(model, M, N, "f32", E.RngSpec(seed=philox_seed))
    else:
        Z1 = torch.empty(N, M // 2, device=eng.tdev); Z2 = torch.empty(N, M // 2, device=eng.tdev)
        for t in range(N):
            Z1[t] = torch.randn(M // 2, device=eng.tdev)
            Z2[t] = torch.randn(M // 2, device=eng.tdev)
        S = eng.paths(model, M, N, "f32", E.RngSpec(z1=Z1, z2=Z2))
    if num_simulations % 2 != 0:  # om3gpu:228-246
        Zo = torch.empty(2, N, 1, device=eng.tdev)
        for t in range(N):
            Zo[0, t] = torch.randn(1, device=eng.tdev)
            Zo[1, t] = torch.randn(1, device=eng.tdev)
        So = eng.paths(model, 1, N, "f32", E.RngSpec(z1=Zo[0].contiguous(), z2=Zo[1].contiguous(), antithetic=False))
        S = torch.cat([S, So], dim=1)
    return S


def create_regression_features(S, K, r, T, t_current):
    """om3:105-121: numpy in, numpy [n, 7] fp64 out (computed by the features kernel)."""
    import torch

    eng = _engine()
    Sd = torch.as_tensor(np.ascontiguousarray(np.asarray(S, dtype=np.float64))).to(eng.tdev)
    return eng.features_ref7(Sd, K, r, T, t_current).cpu().numpy()


def create_regression_features_torch(S_t, K: float, r: float, T: float, t_current: float):
    """om3gpu:342-359: device tensor in, [n, 7] device tensor out, dtype preserved."""
    eng = _engine(S_t.device.index or 0)
    return eng.features_ref7(S_t.contiguous(), K, r, T, t_current)


create_regression_features_torch_vectorized = create_regression_features_torch  # om3gpu:361-377 (same 7 columns)


# ---------------------------------------------------------------------------------------------------
# AdvancedOptionPricer (om3:339-713; GPU variant om3gpu:547-904)
# ---------------------------------------------------------------------------------------------------
class AdvancedOptionPricer:
    def __init__(self, K: float, r: float, sigma: Optional[float], option_type: str = "call",
                 rng_manager: Optional[RNGManager] = None, use_heston: bool = False,
                 heston_params: Optional[Dict[str, Any]] = None, nn_hidden: int = 128, nn_epochs: int = 25,
                 nn_lr: float = 1e-3, verbose: bool = False, iv_model=None, use_streaming: bool = True,
                 chunk_size: int = 500, european_approximation: bool = False, use_control_variate: bool = True,
                 nn_layers: int = 3, nn_dropout: float = 0.10,
                 # engine extensions (not in the reference)
                 lsm_regressor: str = "poly2", semantics: str = "reference", dtype: str = "f32", device: int = 0,
                 gpu_reference_quirks: bool = False, batched: bool = True, out_of_sample: bool = False,
                 control_variate_same_paths: bool = False, path_shard: Optional[Tuple[int, int]] = None,
                 qmc: bool = False):
        self.K = K
        self.r = r
        self.sigma = sigma
        self.option_type = option_type
        self.rng_manager = rng_manager or RNGManager()
        self.use_heston = use_heston
        self.heston_params = heston_params
        self.nn_hidden, self.nn_epochs, self.nn_lr = nn_hidden, nn_epochs, nn_lr
        self.nn_layers, self.nn_dropout = nn_layers, nn_dropout
        self.verbose = verbose
        self.iv_model = iv_model
        self.use_streaming = use_streaming
        self.chunk_size = chunk_size
        self.european_approximation = european_approximation
        self.use_control_variate = use_control_variate
        self.lsm_regressor = lsm_regressor
        self.semantics = semantics
        self.dtype = dtype
        self.device = device
        self.gpu_reference_quirks = gpu_reference_quirks
        self.batched = batched
        self.out_of_sample = out_of_sample  # fit the polynomial on one path set, exercise on an independent one
        self.control_variate_same_paths = control_variate_same_paths  # European leg on the American paths
        # (rank, world) of a path-sharded pricer: price_american_grid prices num_simulations paths PER RANK
        # of every option, the sweep exchanges its totals over NVLink (sharded.init_peer_exchange first)
        self.path_shard = path_shard
        # Sobol' + Brownian-bridge draws with a random digital shift keyed by the pricing's child seed
        # instead of pseudo-random normals (per-point pricing with the polynomial regressor; steps x factors <= 512)
        self.qmc = qmc
        self.last_result: Optional[E.SweepResult] = None
        self._lsm_net = None      # om3gpu:596: the torch-GPU file caches its network across pricing calls
        self._nn_variant = "cpu"  # training defaults of om3:565-613; the *_gpu entry point switches to om3gpu:740-798

    # om3:461-472 model routing
    def _model(self, S0: float, T: float) -> E.ModelSpec:
        if self.iv_model is not None:
            raise NotImplementedError("iv_model prices through price_american_enhanced_lsm (slab + sweep); the fused "
                                      "European / batched entry points take GBM or Heston models")
        if self.use_heston and self.heston_params is not None:
            hp = self.heston_params
            return E.heston(S0, self.r, T, hp["v0"], hp["kappa"], hp["theta"], hp["xi"], hp["rho"])
        if self.sigma is None:
            raise ValueError("sigma is None: provide sigma, iv_model, or heston configuration")
        return E.gbm(S0, self.r, T, self.sigma)

    def price_american_enhanced_lsm(self, S0: float, T: float, num_simulations: int = 10000,
                                    num_time_steps: int = 50) -> float:
        """om3:439-651 with the polynomial regressor.  One path launch + one persistent sweep launch."""
        if S0 <= 0 or self.K <= 0 or T <= 0:
            raise ValueError("S0, K, T must be positive.")
        if self.r < 0:
            raise ValueError("r must be non-negative.")
        if num_simulations <= 0 or num_time_steps <= 0:
            raise ValueError("num_simulations and num_time_steps must be positive integers.")
        seed = int(self.rng_manager.master_rng.integers(0, 2**31 - 1))  # om3:454: the child generator's seed
        torch_seed = self.rng_manager.get_child_seed()                  # om3:455: torch.manual_seed draw
        M = num_simulations // 2 * 2
        if M == 0:
            return float("nan")  # the reference averages an empty cash-flow vector
        if self.iv_model is not None:  # om3:461-462: local-volatility paths, then the same sweep
            eng = _engine(self.device)
            S = eng.paths_localvol(S0, self.r, T, self.iv_model.net, self.K, M, int(num_time_steps), self.dtype, E.RngSpec(seed=seed))
            if self.lsm_regressor == "nn":
                out = eng.lsm_gnet(S, self.K, self.r, T, self.option_type, self.semantics, variant=self._nn_variant,
                                   epochs=self.nn_epochs, lr=self.nn_lr, dropout=self.nn_dropout, seed=int(torch_seed or 0),
                                   arrays=self.verbose)
                self.last_result = out
                return float(out["price"])
            res = eng.lsm(S, self.K, self.r, T, self.option_type, self.lsm_regressor, self.semantics, arrays=self.verbose)
            self.last_result = res
            return float(res.price)
        model = self._model(S0, T)
        if self.lsm_regressor in ("nn", "nn_per_date"):
            # "nn": the reference's own regressor, one SingleLSMNet for all dates (om3:482-651); "nn_per_date": a fresh
            # SingleLSMNet per exercise date (the om2:277-310 loop with om3's network; BASELINE config 3)
            if self.nn_hidden != 128 or self.nn_layers != 3:
                raise ValueError("lsm_regressor='nn' is built for SingleLSMNet(7, 128, 3) (nn_hidden=128, nn_layers=3)")
            eng = _engine(self.device)
            S = eng.paths(model, M, int(num_time_steps), self.dtype, E.RngSpec(seed=seed))
            per_date = self.lsm_regressor == "nn_per_date"
            warm = self.gpu_reference_quirks and self._nn_variant == "gpu" and not per_date  # om3gpu:741-748: one cached network per pricer
            out = eng.lsm_gnet(S, self.K, self.r, T, self.option_type, self.semantics, variant=self._nn_variant,
                               epochs=self.nn_epochs, lr=self.nn_lr, dropout=self.nn_dropout, seed=int(torch_seed or 0),
                               arrays=self.verbose, init_params=self._lsm_net if warm else None, return_params=warm,
                               per_date=1 if per_date else 0)
            if warm:
                self._lsm_net = out["params"]
            self.last_result = out
            return float(out["price"])
        eng = _engine(self.device)
        if self.qmc:
            heston = self.use_heston and self.heston_params is not None
            Z = eng.qmc_normals(M, int(num_time_steps), factors=2 if heston else 1, bridge=True, dtype=self.dtype, shift_seed=seed)
            rs = E.RngSpec(z1=Z[0], z2=Z[1]) if heston else E.RngSpec(z1=Z)
            S = eng.paths(model, M, int(num_time_steps), self.dtype, rs)
            res = eng.lsm(S, self.K, self.r, T, self.option_type, self.lsm_regressor, self.semantics, arrays=self.verbose)
            self.last_result = res
            return float(res.price)
        # Philox key = the master seed, stream = this pricing's child seed: the convention of price_american_grid, so
        # the batched curve and the per-point loop return identical prices point by point
        res = eng.price_american(model, M, int(num_time_steps), self.K, self.option_type, self.dtype,
                                 E.RngSpec(seed=self.rng_manager.master_seed, stream=seed), basis=self.lsm_regressor,
                                 semantics=self.semantics, arrays=self.verbose or self.out_of_sample)
        if self.out_of_sample:  # the fitted policy priced on fresh paths (the next Philox key, same stream)
            S_new = eng.paths(model, M, int(num_time_steps), self.dtype,
                              E.RngSpec(seed=self.rng_manager.master_seed + 1, stream=seed))
            res = eng.lsm_apply_policy(S_new, res.betas, self.K, self.r, T, self.option_type, self.lsm_regressor,
                                       self.semantics, arrays=self.verbose)
        self.last_result = res
        return float(res.price)

    def price_american_enhanced_lsm_gpu(self, S0: float, T: float, num_simulations: int = 10000,
                                        num_time_steps: int = 50) -> float:
        """om3gpu:655-839.  With gpu_reference_quirks=True also applies om3gpu:667-669 (step override for
        T < 10 days) and om3gpu:675 (50 000-path cap); both are memory work-arounds of the reference."""
        if S0 <= 0 or self.K <= 0 or T <= 0:
            raise ValueError("S0, K, T must be positive.")
        if self.gpu_reference_quirks:
            if T < 10 / 365.0:
                num_time_steps = max(10, min(25, int(T * 365 * 2)))
            num_simulations = min(num_simulations, 50000)
        prev, self._nn_variant = self._nn_variant, "gpu"
        try:
            return self.price_american_enhanced_lsm(S0, T, num_simulations, num_time_steps)
        finally:
            self._nn_variant = prev

    def price_european_streaming(self, S0: float, T: float, num_simulations: int = 10000,
                                 num_time_steps: int = 50) -> float:
        """om3:382-437.  One fused no-store launch replaces the 500-path chunk loop; the master generator
        advances once per reference chunk (om3:392) so later calls see the same seed tree."""
        n_chunks = max(1, -(-int(num_simulations) // int(self.chunk_size)))
        seeds = [int(self.rng_manager.get_child_seed()) for _ in range(n_chunks)]
        n = int(num_simulations)
        anti = n % 2 == 0
        if self.iv_model is not None:  # om3:394-395: local-volatility paths, then the payoff of the terminal row
            eng = _engine(self.device)
            M = n // 2 * 2
            S = eng.paths_localvol(S0, self.r, T, self.iv_model.net, self.K, M, int(num_time_steps), self.dtype,
                                   E.RngSpec(seed=seeds[0], stream=0x45))
            mean, se = eng.european_from_slab(S[int(num_time_steps)].contiguous(), self.K, self.r, T, self.option_type)
            if self.verbose:
                print(f"European streaming MC: {mean:.4f} ± {se:.4f} (n={M})")
            return float(mean)
        model = self._model(S0, T)
        mean, se = _engine(self.device).price_european_batch(model, n, int(num_time_steps), [self.K], [T],
                                                             [1 if self.option_type == "put" else 0], self.dtype,
                                                             E.RngSpec(seed=self.rng_manager.master_seed,
                                                                       stream=0x45 + (seeds[0] & 0x7FFFFFFF), antithetic=anti))
        if self.verbose:
            print(f"European streaming MC: {mean[0]:.4f} ± {se[0]:.4f} (n={n})")
        return float(mean[0])

    price_european_gpu = price_european_streaming  # om3gpu:605-653

    def price_american_with_control_variate(self, S0: float, T: float, num_simulations: int = 10000,
                                            num_time_steps: int = 50) -> float:
        """om3:653-677: american + 1.0 * (BS_analytic - european_MC), European leg on independent paths."""
        if self.control_variate_same_paths and self.use_control_variate and self.sigma is not None \
                and self.iv_model is None and not (self.use_heston and self.heston_params is not None) \
                and not self.lsm_regressor.startswith("nn"):
            # the European payoff of the SAME paths as the control -- one slab, one sweep, one reduction of
            # its terminal row (the reference simulates the European leg independently, which adds its variance)
            if S0 <= 0 or self.K <= 0 or T <= 0:
                raise ValueError("S0, K, T must be positive.")
            seed = int(self.rng_manager.master_rng.integers(0, 2**31 - 1))
            self.rng_manager.get_child_seed()
            M, N = num_simulations // 2 * 2, int(num_time_steps)
            eng = _engine(self.device)
            S = eng.paths(self._model(S0, T), M, N, self.dtype, E.RngSpec(seed=self.rng_manager.master_seed, stream=seed))
            res = eng.lsm(S, self.K, self.r, T, self.option_type, self.lsm_regressor, self.semantics, arrays=False)
            eu_mc, _ = eng.european_from_slab(S[N].contiguous(), self.K, self.r, T, self.option_type)
            eu_exact = BlackScholesGreeks.black_scholes_price(S0, self.K, T, self.r, self.sigma, self.option_type)
            self.last_result = res
            return float(res.price + (eu_exact - eu_mc))
        american_price = self.price_american_enhanced_lsm(S0, T, num_simulations, num_time_steps)
        if not self.use_control_variate or self.sigma is None:
            return american_price
        european_mc = self.price_european_streaming(S0, T, num_simulations, num_time_steps)
        european_analytical = BlackScholesGreeks.black_scholes_price(S0, self.K, T, self.r, self.sigma, self.option_type)
        american_cv = american_price + 1.0 * (european_analytical - european_mc)
        if self.verbose:
            print(f"American: {american_price:.4f}, European MC: {european_mc:.4f}, "
                  f"European Analytical: {european_analytical:.4f}, CV Adjusted: {american_cv:.4f}")
        return american_cv

    def price_american_option(self, S0: float, T: float, num_simulations: int = 10000, num_time_steps: int = 50,
                              plot_paths: bool = False) -> float:
        """om3:679-695 routing."""
        if self.use_streaming and self.european_approximation:
            if self.verbose:
                print("WARNING: Using European approximation for American option (streaming mode)")
            return self.price_european_streaming(S0, T, num_simulations, num_time_steps)
        if self.use_control_variate and self.sigma is not None:
            return self.price_american_with_control_variate(S0, T, num_simulations, num_time_steps)
        return self.price_american_enhanced_lsm(S0, T, num_simulations, num_time_steps)

    def price_american_grid(self, S0, T, num_time_steps, num_simulations: int = 10000, K=None, european: bool = False,
                            seeds=None):
        """Engine extension: price a whole grid (arrays S0 / T / steps / K broadcast against each other) with
        optmc_price_american_batch -- what the reference does with one price_american_enhanced_lsm call per
        grid point (om3:706-712).  The master generator advances exactly as in the per-point loop (om3:454-455);
        point i draws from Philox stream = its child seed."""
        S0a, Ta, Na, Ka = np.broadcast_arrays(np.asarray(S0, dtype=np.float64), np.asarray(T, dtype=np.float64),
                                              np.asarray(num_time_steps, dtype=np.int64),
                                              np.asarray(self.K if K is None else K, dtype=np.float64))
        shape = S0a.shape
        S0a, Ta, Na, Ka = (np.atleast_1d(a).ravel() for a in (S0a, Ta, Na, Ka))
        if np.any(S0a <= 0) or np.any(Ka <= 0) or np.any(Ta <= 0):
            raise ValueError("S0, K, T must be positive.")
        if self.r < 0:
            raise ValueError("r must be non-negative.")
        if num_simulations <= 0 or np.any(Na <= 0):
            raise ValueError("num_simulations and num_time_steps must be positive integers.")
        if seeds is None:
            seeds = []
            for _ in range(S0a.size):
                seeds.append(int(self.rng_manager.master_rng.integers(0, 2**31 - 1)))  # om3:454
                self.rng_manager.get_child_seed()                                      # om3:455
        M = num_simulations // 2 * 2
        model = self._model(float(S0a[0]), float(Ta[0]))
        rng = E.RngSpec(seed=self.rng_manager.master_seed)
        M_total = 0
        if self.path_shard is not None:
            rank, world = self.path_shard
            M_total = M * world
            rng = E.RngSpec(seed=self.rng_manager.master_seed, pair_offset=rank * (M // 2))
        out = _engine(self.device).price_american_batch(
            model, M, S0a, Ka, Ta, Na, 1 if self.option_type == "put" else 0, self.dtype, rng,
            basis=self.lsm_regressor, semantics=self.semantics, streams=seeds, european=european, M_total=M_total)
        if european or M_total:
            price, se, extras = out
            if european:
                return price.reshape(shape), se.reshape(shape), extras["european"][:, 0].reshape(shape)
        else:
            price, se = out
        return price.reshape(shape), se.reshape(shape)

    def compute_curve_for_S0(self, S0: float, intervals_per_day: int, total_points: int, num_simulations: int,
                             plot_paths: bool) -> List[Dict[str, Any]]:
        """om3:697-713.  The whole curve is one batched engine call (two with the independent control-variate leg):
        price_american_option's default route (om3:692-693) is american + (BS - european_MC); the European leg is either
        reduced from each option's OWN terminal row inside the grouped sweep launch (control_variate_same_paths) or
        priced on independent paths by one fused optmc_price_european_grid launch, as the reference's
        price_european_streaming call per grid point does.  The generators advance exactly as in the per-point loop."""
        cv = self.use_control_variate and self.sigma is not None
        eu = self.use_streaming and self.european_approximation
        grid_ok = (self.batched and not eu and total_points > 0 and self.iv_model is None and not self.lsm_regressor.startswith("nn")
                   and not self.qmc)
        if grid_ok:
            days = np.array([i / intervals_per_day for i in range(total_points, 0, -1)])
            steps = np.maximum(10, np.minimum(130, np.ceil(days))).astype(np.int64)
            T = days / 365
            if not cv:
                prices, _ = self.price_american_grid(S0, T, steps, num_simulations)
            else:
                n_chunks = max(1, -(-int(num_simulations) // int(self.chunk_size)))
                am_seeds, eu_seeds = [], []
                for _ in range(total_points):  # per point: the American draws (om3:454-455), then the European chunks (om3:392)
                    am_seeds.append(int(self.rng_manager.master_rng.integers(0, 2**31 - 1)))
                    self.rng_manager.get_child_seed()
                    if not self.control_variate_same_paths:
                        eu_seeds.append([int(self.rng_manager.get_child_seed()) for _ in range(n_chunks)][0] & 0x7FFFFFFF)
                put = 1 if self.option_type == "put" else 0
                if self.control_variate_same_paths:
                    am, _, eu_mc = self.price_american_grid(S0, T, steps, num_simulations, european=True, seeds=am_seeds)
                else:
                    am, _ = self.price_american_grid(S0, T, steps, num_simulations, seeds=am_seeds)
                    n = int(num_simulations)
                    eu_mc, _ = _engine(self.device).price_european_grid(
                        self._model(S0, float(T[0])), n, S0, self.K, T, steps, put, self.dtype,
                        E.RngSpec(seed=self.rng_manager.master_seed, stream=0x45, antithetic=n % 2 == 0),
                        stream_id=np.asarray(eu_seeds, dtype=np.int64).astype(np.int32))
                exact = np.array([BlackScholesGreeks.black_scholes_price(S0, self.K, float(t), self.r, self.sigma,
                                                                         self.option_type) for t in T])
                prices = am + 1.0 * (exact - eu_mc)
            return [{"S0": S0, "Days to Expiry": float(d), "Option Value": float(p)} for d, p in zip(days, prices)]
        records = []
        for i in range(total_points, 0, -1):
            d = i / intervals_per_day
            T = d / 365
            steps = max(10, min(130, int(np.ceil(d))))
            est_price = self.price_american_option(S0, T, num_simulations, steps, plot_paths)
            records.append({"S0": S0, "Days to Expiry": d, "Option Value": est_price})
        return records


def compute_curve_worker_enhanced(S0, K, r, sigma, option_type, worker_seed, intervals_per_day, total_points,
                                  num_simulations, plot_paths, use_heston, heston_params, nn_hidden=128, nn_epochs=25,
                                  nn_lr=1e-3, verbose=False, european_approximation=False, use_control_variate=True):
    """om3:719-739: errors are logged and an empty list is returned, as in the reference."""
    try:
        pricer = AdvancedOptionPricer(K, r, sigma, option_type, RNGManager(worker_seed), use_heston, heston_params,
                                      nn_hidden=nn_hidden, nn_epochs=nn_epochs, nn_lr=nn_lr, verbose=verbose,
                                      european_approximation=european_approximation,
                                      use_control_variate=use_control_variate)
        return pricer.compute_curve_for_S0(S0, intervals_per_day, total_points, num_simulations, plot_paths)
    except Exception as e:  # noqa: BLE001 -- reference behaviour (om3:737-739)
        logging.error(f"Error in enhanced worker for S0={S0}: {e}")
        return []


compute_curve_worker_gpu = compute_curve_worker_enhanced  # om3gpu:910-932 (same argument list)


def compute_multiple_S0_gpu_batch(s0_list, K, r, sigma, option_type, intervals_per_day, total_points,
                                  num_simulations, nn_hidden=128, nn_epochs=25, nn_lr=1e-3, verbose=False,
                                  european_approximation=False, use_control_variate=True, seed=42):
    """om3gpu:934-956: one pricer reused across the S0 list."""
    pricer = AdvancedOptionPricer(K, r, sigma, option_type, RNGManager(seed), nn_hidden=nn_hidden,
                                  nn_epochs=nn_epochs, nn_lr=nn_lr, verbose=verbose,
                                  european_approximation=european_approximation,
                                  use_control_variate=use_control_variate)
    records = []
    for S0 in s0_list:
        records.extend(pricer.compute_curve_for_S0(S0, intervals_per_day, total_points, num_simulations, False))
    return records


# ---------------------------------------------------------------------------------------------------
# older API the Streamlit UIs import (om2:176-457, om1:44-211)
# ---------------------------------------------------------------------------------------------------
class OptionPricer:
    """om2:176-355.  regressor="nn" (default, the reference's behaviour, om2:277-310): a fresh ContNet per exercise
    date trained with nn_epochs full-batch Adam steps (optmc_lsm_mlp; nn_hidden 32 -> fp32 CUDA cores, 128 -> bf16
    tcgen05 tensor cores).  regressor="poly": the polynomial of SURVEY.md 8(c) with lsm_poly_degree 2 or 3 (the
    reference validates lsm_poly_degree and then ignores it, om2:176-180)."""

    def __init__(self, K: float, r: float, sigma: Optional[float], option_type: str = "call",
                 lsm_poly_degree: int = 2, seed: int = 42, use_heston: bool = False,
                 heston_params: Optional[Dict[str, Any]] = None, nn_hidden: int = 32, nn_epochs: int = 10,
                 nn_lr: float = 1e-3, verbose: bool = False, regressor: str = "nn"):
        self.K, self.r, self.sigma, self.option_type = K, r, sigma, option_type
        self.lsm_poly_degree, self.seed = lsm_poly_degree, seed
        self.use_heston, self.heston_params = use_heston, heston_params
        self.nn_hidden, self.nn_epochs, self.nn_lr, self.verbose = nn_hidden, nn_epochs, nn_lr, verbose
        self.regressor = regressor

    def price_american_option(self, S0: float, T: float, num_simulations: int = 10000, num_time_steps: int = 50,
                              plot_paths: bool = False) -> float:
        if S0 <= 0 or self.K <= 0 or T <= 0 or (self.sigma is None and not self.use_heston):
            raise ValueError("S0, K, T, and sigma must be positive.")
        if self.r < 0:
            raise ValueError("r must be non-negative.")
        if num_simulations <= 0 or num_time_steps <= 0:
            raise ValueError("num_simulations and num_time_steps must be positive integers.")
        if self.lsm_poly_degree < 0 or not isinstance(self.lsm_poly_degree, int):
            raise ValueError("lsm_poly_degree must be a non-negative integer.")
        if self.option_type not in ("call", "put"):
            raise ValueError("option_type must be 'call' or 'put'.")
        M = num_simulations // 2 * 2
        if self.use_heston and self.heston_params is not None:
            hp = self.heston_params
            model = E.heston(S0, self.r, T, hp["v0"], hp["kappa"], hp["theta"], hp["xi"], hp["rho"])
        else:
            model = E.gbm(S0, self.r, T, self.sigma)
        if self.regressor == "nn":
            eng = _engine()
            S = eng.paths(model, M, int(num_time_steps), "f32", E.RngSpec(seed=int(self.seed)))
            res = eng.lsm_mlp(S, self.K, self.r, T, self.option_type, "reference", hidden=int(self.nn_hidden),
                              epochs=int(self.nn_epochs), lr=float(self.nn_lr), seed=int(self.seed), arrays=False)
            return float(res.price)
        basis = "poly3" if self.lsm_poly_degree >= 3 else "poly2"
        res = _engine().price_american(model, M, int(num_time_steps), self.K, self.option_type, "f32",
                                       E.RngSpec(seed=int(self.seed)), basis=basis)
        return float(res.price)

    def compute_curve_for_S0(self, S0, intervals_per_day, total_points, num_simulations, plot_paths):
        records = []
        for i in range(total_points, 0, -1):  # om2:346-354
            d = i / intervals_per_day
            T = d / 365
            steps = max(10, min(130, int(np.ceil(d))))
            records.append({"S0": S0, "Days to Expiry": d,
                            "Option Value": self.price_american_option(S0, T, num_simulations, steps, plot_paths)})
        return records


class om1:
    """Options_model.py (om1:44-211), the API `options_ui.py:13` imports: module-level functions returning
    (mean, std, probability of expiring worthless).  Bind them on the reference side with
    ``price_american_option = om1.price_american_option`` etc."""

    @staticmethod
    def price_american_option(S0, K, T, r, sigma, num_simulations=10000, num_time_steps=50, option_type="call",
                              lsm_poly_degree=2, plot_paths=False, seed=42):
        """om1:44-170: GBM antithetic paths from ``np.random.seed(seed)`` draws (replayed exactly), per-date ContNet
        regression (10 full-batch Adam steps, lr 1e-3), sticky mask, N-1 discounts.  -> (mean, std, zero_prob)."""
        import torch

        if S0 <= 0 or K <= 0 or T <= 0 or sigma <= 0:
            raise ValueError("S0, K, T, and sigma must be positive.")
        if r < 0:
            raise ValueError("r must be non-negative.")
        if num_simulations <= 0 or num_time_steps <= 0:
            raise ValueError("num_simulations and num_time_steps must be positive integers.")
        if lsm_poly_degree < 0 or not isinstance(lsm_poly_degree, int):
            raise ValueError("lsm_poly_degree must be a non-negative integer.")
        if option_type not in ("call", "put"):
            raise ValueError("option_type must be 'call' or 'put'.")
        eng = _engine()
        M, N = num_simulations // 2 * 2, int(num_time_steps)
        if M == 0:
            return float("nan"), float("nan"), float("nan")
        Z = np.random.RandomState(seed).standard_normal((N, M // 2))  # = np.random.seed(seed); standard_normal (om1:73-82)
        S = eng.paths(E.gbm(S0, r, T, sigma), M, N, "f64", E.RngSpec(z1=torch.from_numpy(Z).to(eng.tdev)))
        res = eng.lsm_mlp(S, K, r, T, option_type, "reference", hidden=32, epochs=10, lr=1e-3, seed=int(seed), arrays=False)
        std = res.stderr * math.sqrt(M) * math.sqrt((M - 1) / M) if M > 1 else 0.0  # np.std: population
        return float(res.price), float(std), eng.lsm_zero_cashflows() / M

    @staticmethod
    def compute_curve_for_S0(S0, K, r, sigma, num_simulations, intervals_per_day, total_points, option_type,
                             lsm_poly_degree, plot_paths, seed):
        """om1:190-211."""
        records = []
        for i in range(total_points, 0, -1):
            d = i / intervals_per_day
            T = d / 365
            steps = max(10, min(130, int(np.ceil(d))))
            est, std, zp = om1.price_american_option(S0, K, T, r, sigma, num_simulations, steps, option_type,
                                                     lsm_poly_degree, plot_paths, seed)
            records.append({"S0": S0, "Days to Expiry": d, "Option Value": est, "Std Dev": std, "Zero Prob": zp})
        return records


def compute_curve_worker(S0, K, r, sigma, option_type, lsm_poly_degree, seed, intervals_per_day, total_points,
                         num_simulations, plot_paths, use_heston, heston_params, nn_hidden=32, nn_epochs=10,
                         nn_lr=1e-3, verbose=False, regressor="nn"):
    """om2:443-457."""
    try:
        pricer = OptionPricer(K, r, sigma, option_type, lsm_poly_degree, seed, use_heston, heston_params, nn_hidden,
                              nn_epochs, nn_lr, verbose, regressor)
        return pricer.compute_curve_for_S0(S0, intervals_per_day, total_points, num_simulations, plot_paths)
    except Exception as e:  # noqa: BLE001 -- om2:455-457
        logging.error(f"Error in worker for S0={S0}: {e}")
        return []


# ---------------------------------------------------------------------------------------------------
# Heston calibration pricer (hc:34-90, hc:197-312)
# ---------------------------------------------------------------------------------------------------
@dataclass
class HestonParams:
    kappa: float
    theta: float
    sigma: float
    rho: float
    v0: float

    # open intervals the calibrator accepts (hc:43-54); the messages are the reference's
    _BOUNDS = (("kappa", 0, 20), ("theta", 0, 2), ("sigma", 0, 3), ("rho", -1, 1), ("v0", 0, 2))

    def __post_init__(self):
        for name, lo, hi in self._BOUNDS:
            value = getattr(self, name)
            if not lo < value < hi:
                raise ValueError(f"{name}={value} must be in ({lo}, {hi})")

    def to_array(self) -> np.ndarray:
        return np.array([getattr(self, name) for name, _, _ in self._BOUNDS])

    @classmethod
    def from_array(cls, x: np.ndarray) -> "HestonParams":
        return cls(*(x[i] for i in range(len(cls._BOUNDS))))

    def feller_condition(self) -> bool:
        return 2 * self.kappa * self.theta >= self.sigma**2


def _calibration_config_fields():
    """Field table of the calibrator's configuration (hc:75-90).  The pricing path reads n_mc_paths, n_time_steps,
    use_antithetic, seed, use_vega_weighting and min_vega_weight; the rest belongs to the optimiser / reporting side
    (out of scope) and is carried so that a reference-side `CalibrationConfig(**kwargs)` keeps working."""
    return [("use_vega_weighting", bool, True), ("min_vega_weight", float, 0.01), ("max_iterations", int, 2000),
            ("tolerance", float, 1e-8), ("n_mc_paths", int, 100000), ("n_time_steps", int, 100), ("use_antithetic", bool, True),
            ("seed", int, 42), ("verbose", bool, True), ("plot_results", bool, True),
            ("optimization_methods", List[str], field(default_factory=lambda: ["L-BFGS-B", "differential_evolution", "dual_annealing"])),
            ("fallback_enabled", bool, True), ("regime_detection", bool, True)]


CalibrationConfig = make_dataclass("CalibrationConfig", _calibration_config_fields())


class HestonPricer:
    """hc:197-312.  simulate_paths draws from the persistent numpy generator exactly as the reference
    (hc:226-227) and returns path-major arrays; the pricing methods use the fused no-store kernel with
    Philox keyed from that generator (one integer draw per call), unless reference_draws=True."""

    def __init__(self, config: CalibrationConfig, reference_draws: bool = False, dtype: str = "f32", device: int = 0):
        self.config = config
        self.rng = np.random.default_rng(config.seed)
        self.reference_draws = reference_draws
        self.dtype = dtype
        self.device = device

    def _model(self, params: HestonParams, S0: float, T: float, r: float) -> E.ModelSpec:
        return E.heston(S0, r, T, params.v0, params.kappa, params.theta, params.sigma, params.rho,
                        scheme=L.SCHEME_HESTON_REF_CALIB)

    def _simulate_device(self, params: HestonParams, S0: float, T: float, r: float):
        """Draw (Z1, Z2_indep) from the persistent generator as hc:226-227 does, step the calibrator
        scheme on the device in fp64.  Returns step-major device slabs S, V [(N+1), M] and M."""
        import torch

        eng = _engine(self.device)
        n_paths, N = self.config.n_mc_paths, self.config.n_time_steps
        anti = bool(self.config.use_antithetic)
        n_sim = n_paths // 2 if anti else n_paths
        Z1 = self.rng.standard_normal((n_sim, N))
        Z2i = self.rng.standard_normal((n_sim, N))
        z1 = torch.from_numpy(Z1).to(eng.tdev).t().contiguous()   # step-major [N][n_sim]
        z2 = torch.from_numpy(Z2i).to(eng.tdev).t().contiguous()
        M = 2 * n_sim if anti else n_sim
        S, V = eng.paths(self._model(params, S0, T, r), M, N, "f64", E.RngSpec(z1=z1, z2=z2, antithetic=anti),
                         return_v=True)
        return S, V, M

    def simulate_paths(self, params: HestonParams, S0: float, T: float, r: float = 0.05) -> Tuple[np.ndarray, np.ndarray]:
        n_paths, N = self.config.n_mc_paths, self.config.n_time_steps
        S, V, M = self._simulate_device(params, S0, T, r)
        Sh = np.zeros((n_paths, N + 1)); Vh = np.zeros((n_paths, N + 1))
        Sh[:M] = S.t().cpu().numpy(); Vh[:M] = V.t().cpu().numpy()   # path-major, as hc:216-217
        if M < n_paths:  # odd n_paths with antithetic: the reference's vstack would fail; keep the row at S0/v0
            Sh[M:, 0] = S0; Vh[M:, 0] = params.v0
        return Sh, Vh

    def price_european_option(self, params: HestonParams, S0: float, K: float, T: float, r: float = 0.05,
                              option_type: str = "call") -> float:
        try:
            ot = option_type.lower()
            if ot not in ("call", "put"):
                raise ValueError(f"Unknown option type: {option_type}")
            if self.reference_draws:
                S, _, M = self._simulate_device(params, S0, T, r)
                mean, _ = _engine(self.device).european_from_slab(S[self.config.n_time_steps].contiguous(), K, r, T, ot)
                return float(mean)
            seed = int(self.rng.integers(0, 2**63 - 1))
            mean, _ = _engine(self.device).price_european_batch(
                self._model(params, S0, T, r), self.config.n_mc_paths // 2 * 2 if self.config.use_antithetic
                else self.config.n_mc_paths, self.config.n_time_steps, [K], [T], [1 if ot == "put" else 0], self.dtype,
                E.RngSpec(seed=seed, antithetic=bool(self.config.use_antithetic)))
            return float(mean[0])
        except Exception as e:  # noqa: BLE001 -- hc:279-281
            print(f"Warning: Pricing failed for K={K}, T={T}: {e}")
            return float("nan")

    def price_options_batch(self, params: HestonParams, S0: float, K_array: np.ndarray, T_array: np.ndarray,
                            r: float = 0.05) -> np.ndarray:
        """hc:283-312: call options only (hc:305); options with the same T share one set of paths."""
        K_array = np.asarray(K_array, dtype=np.float64)
        T_array = np.asarray(T_array, dtype=np.float64)
        prices = np.zeros(len(K_array))
        if len(K_array) == 0:
            return prices
        try:
            uniq, inv = np.unique(T_array, return_inverse=True)
            seed = int(self.rng.integers(0, 2**63 - 1))
            M = self.config.n_mc_paths // 2 * 2 if self.config.use_antithetic else self.config.n_mc_paths
            mean, _ = _engine(self.device).price_european_batch(
                self._model(params, S0, float(T_array[0]), r), M, self.config.n_time_steps, K_array, T_array,
                np.zeros(len(K_array), dtype=np.int32), self.dtype,
                E.RngSpec(seed=seed, antithetic=bool(self.config.use_antithetic)), stream_id=inv.astype(np.int32))
            prices[:] = mean
        except Exception as e:  # noqa: BLE001 -- hc:308-310
            print(f"Warning: Batch pricing failed: {e}")
            prices[:] = np.nan
        return prices


# ---------------------------------------------------------------------------------------------------
# calibration objective (hc:404-472) -- the reference prices the surface row by row (one simulation per
# row, ~170 s per evaluation); here the whole surface is ONE fused no-store launch
# ---------------------------------------------------------------------------------------------------
def bs_price(S: float, K: float, T: float, r: float, sigma: float, option_type: str = "call") -> float:
    """hc:326-345."""
    if T <= 0 or sigma <= 0:
        return max(S - K, 0) if option_type == "call" else max(K - S, 0)
    return float(BlackScholesGreeks.black_scholes_price(S, K, T, r, sigma, option_type.lower()))


def bs_vega(S: float, K: float, T: float, r: float, sigma: float) -> float:
    """hc:314-324."""
    if T <= 0 or sigma <= 0:
        return 1e-8
    d1 = (math.log(S / K) + (r + 0.5 * sigma**2) * T) / (sigma * math.sqrt(T))
    return max(float(S * math.exp(-0.5 * d1 * d1) / math.sqrt(2.0 * math.pi) * math.sqrt(T)), 1e-8)


def objective_from_prices(x: np.ndarray, heston_prices: np.ndarray, S0: float, r: float, K: np.ndarray, T: np.ndarray,
                          sigma_iv: np.ndarray, use_vega_weighting: bool = True, min_vega_weight: float = 0.01) -> float:
    """hc:420-472 given the model prices of the rows: vega-weighted RMSE of ln(P_heston / P_bs(sigma_market)) plus the
    Feller penalty; rows with a NaN / tiny price are skipped; 1e6 for invalid parameters or an empty sum."""
    try:
        params = HestonParams.from_array(np.asarray(x, dtype=np.float64))
    except (ValueError, TypeError):
        return 1e6
    total_error = total_weight = 0.0
    for hp_, k, t, iv in zip(heston_prices, K, T, sigma_iv):
        if np.isnan(hp_) or hp_ <= 1e-8:
            continue
        bs = bs_price(S0, float(k), float(t), r, float(iv), "call")
        if bs <= 1e-8:
            continue
        w = max(bs_vega(S0, float(k), float(t), r, float(iv)) / 100.0, min_vega_weight) if use_vega_weighting else 1.0
        total_error += w * math.log(hp_ / bs) ** 2
        total_weight += w
    if total_weight == 0:
        return 1e6
    penalty = 0.0 if params.feller_condition() else 100.0 * abs(2 * params.kappa * params.theta - params.sigma**2)
    return math.sqrt(total_error / total_weight) + penalty


class HestonObjective:
    """``HestonCalibrator._objective_function`` (hc:404-472) as a callable for any optimiser (scipy.optimize etc.):
    ``f = HestonObjective(pricer, S0, r, K, T, sigma_iv); f(x)`` with x = (kappa, theta, sigma, rho, v0).
    One ``optmc_price_european_batch`` launch prices every row with its own paths, as the reference simulates per row.
    common_random_numbers=True keeps the Philox key fixed across evaluations: a smooth objective for
    gradient-based optimisers instead of the reference's re-seeded noise."""

    def __init__(self, pricer: "HestonPricer", S0: float, r: float, K, T, sigma_iv, common_random_numbers: bool = False,
                 rank: int = 0, world: int = 1, all_gather=None):
        """rank / world / all_gather: option-sharded evaluation over several GPUs (the reference loops over
        the rows, hc:424-433): every rank prices rows rank::world in one launch, ``all_gather(local_prices) ->
        [prices of rank 0, rank 1, ...]`` (host plumbing, e.g. torch.distributed.all_gather_object) is the only exchange,
        and every rank returns the same objective.  All ranks must construct the pricer with the same seed."""
        self.rank, self.world, self.all_gather = int(rank), int(world), all_gather
        if self.world > 1 and all_gather is None:
            raise ValueError("world > 1 needs an all_gather callable")
        self.pricer, self.S0, self.r = pricer, float(S0), float(r)
        self.K = np.asarray(K, dtype=np.float64).ravel()
        self.T = np.asarray(T, dtype=np.float64).ravel()
        self.sigma_iv = np.asarray(sigma_iv, dtype=np.float64).ravel()
        self.crn = common_random_numbers
        self._seed = int(pricer.rng.integers(0, 2**63 - 1))
        self.last_prices: Optional[np.ndarray] = None

    def prices(self, params: "HestonParams") -> np.ndarray:
        cfg = self.pricer.config
        seed = self._seed if self.crn else int(self.pricer.rng.integers(0, 2**63 - 1))
        M = cfg.n_mc_paths // 2 * 2 if cfg.use_antithetic else cfg.n_mc_paths
        n = len(self.K)
        rows = np.arange(self.rank, n, self.world)  # this rank's rows; row i keeps Philox stream i on any rank count
        mean = np.zeros(0)
        if rows.size:
            mean, _ = _engine(self.pricer.device).price_european_batch(
                self.pricer._model(params, self.S0, float(self.T[0]), self.r), M, cfg.n_time_steps, self.K[rows],
                self.T[rows], np.zeros(rows.size, dtype=np.int32), self.pricer.dtype,
                E.RngSpec(seed=seed, antithetic=bool(cfg.use_antithetic)), stream_id=rows.astype(np.int32))
        if self.world == 1:
            return np.asarray(mean)
        parts = self.all_gather(np.asarray(mean))
        out = np.empty(n)
        for rk, part in enumerate(parts):
            out[rk::self.world] = np.asarray(part)
        return out

    def __call__(self, x) -> float:
        try:
            params = HestonParams.from_array(np.asarray(x, dtype=np.float64))
        except (ValueError, TypeError):
            return 1e6
        try:
            self.last_prices = self.prices(params)
        except Exception as e:  # noqa: BLE001 -- the reference skips rows whose pricing failed (hc:456-457)
            print(f"Warning: Batch pricing failed: {e}")
            return 1e6
        cfg = self.pricer.config
        return objective_from_prices(x, self.last_prices, self.S0, self.r, self.K, self.T, self.sigma_iv,
                                     cfg.use_vega_weighting, cfg.min_vega_weight)
