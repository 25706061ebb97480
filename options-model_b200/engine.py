"""Thin Python handle over the C ABI (include/optmc.h).  torch is used only to own device memory and
streams; every computation happens in liboptmc.so's CUDA kernels.  No CPU fallback: constructing an
``Engine`` without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


@dataclass
class ModelSpec:
    """Mirror of optmc_model_params."""
    model: int
    scheme: int
    S0: float
    r: float
    T: float
    sigma: float = 0.0
    v0: float = 0.0
    kappa: float = 0.0
    theta: float = 0.0
    xi: float = 0.0
    rho: float = 0.0

    def c(self) -> L.ModelParams:
        return L.ModelParams(self.model, self.scheme, self.S0, self.r, self.T, self.sigma, self.v0, self.kappa,
                             self.theta, self.xi, self.rho)


def gbm(S0, r, T, sigma, scheme=L.SCHEME_GBM_LOG_EULER) -> ModelSpec:
    return ModelSpec(L.MODEL_GBM, scheme, float(S0), float(r), float(T), sigma=float(sigma))


def heston(S0, r, T, v0, kappa, theta, xi, rho, scheme=L.SCHEME_HESTON_REF_ABSORB) -> ModelSpec:
    return ModelSpec(L.MODEL_HESTON, scheme, float(S0), float(r), float(T), v0=float(v0), kappa=float(kappa),
                     theta=float(theta), xi=float(xi), rho=float(rho))


@dataclass
class RngSpec:
    """Mirror of optmc_rng_params.  z1/z2: optional torch CUDA tensors of external normals [N][M/2]."""
    seed: int = 42
    stream: int = 0
    z1: object = None
    z2: object = None
    antithetic: bool = True
    pair_offset: int = 0

    def c(self) -> L.RngParams:
        zd = L.F32
        p1 = p2 = None
        if self.z1 is not None:
            import torch

            assert self.z1.is_cuda and self.z1.is_contiguous()
            zd = L.F64 if self.z1.dtype == torch.float64 else L.F32
            p1 = self.z1.data_ptr()
            if self.z2 is not None:
                assert self.z2.is_cuda and self.z2.is_contiguous() and self.z2.dtype == self.z1.dtype
                p2 = self.z2.data_ptr()
        return L.RngParams(self.seed & 0xFFFFFFFFFFFFFFFF, self.stream & 0xFFFFFFFFFFFFFFFF, p1, p2, zd,
                           1 if self.antithetic else 0, self.pair_offset)


@dataclass
class SweepResult:
    price: float
    stderr: float
    n_paths: int
    impl_used: int
    n_launches: int
    betas: Optional[np.ndarray] = field(default=None, repr=False)
    boundary: Optional[np.ndarray] = field(default=None, repr=False)
    ex_count: Optional[np.ndarray] = field(default=None, repr=False)
    n_itm: Optional[np.ndarray] = field(default=None, repr=False)


def _dtype_code(dtype) -> int:
    """Accepts "f32"/"f64", numpy / torch dtypes or the C enum value."""
    if isinstance(dtype, int):
        return L.F64 if dtype == L.F64 else L.F32
    return L.F64 if str(dtype).replace("torch.", "") in ("f64", "float64", "double", "<class 'numpy.float64'>") else L.F32


def _torch_dtype(code: int):
    import torch

    return torch.float64 if code == L.F64 else torch.float32


def workspace_estimate(M: int, N: int, dtype="f32", n_options: int = 1) -> int:
    """Upper estimate of the device memory price_american(_batch) allocates; needs no GPU."""
    b = C.c_int64()
    L.check(L.load_library().optmc_workspace_bytes(int(M), int(N), _dtype_code(dtype), int(n_options), C.byref(b)))
    return int(b.value)


class Engine:
    """One context per (device, host thread).  Not thread-safe (SURVEY.md 8(b) threading)."""

    def __init__(self, device: int = 0, follow_torch_stream: bool = True):
        self.lib = L.load_library()
        import torch

        if not torch.cuda.is_available():
            raise L.OptmcError("no CUDA device visible to torch; options_model_b200 has no CPU fallback")
        self.torch = torch
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        h = C.c_void_p()
        L.check(self.lib.optmc_ctx_create(self.device, C.byref(h)))
        self._h = h
        self.follow_torch_stream = follow_torch_stream
        info = (C.c_int64 * 4)()
        L.check(self.lib.optmc_ctx_device_info(self._h, info))
        self.sm_count, self.l2_bytes, self.max_smem, self.cc = (int(x) for x in info)

    def workspace_bytes(self) -> int:
        """Device bytes the context holds right now (the library's own workspaces; they only grow)."""
        b = C.c_int64()
        L.check(self.lib.optmc_ctx_workspace_bytes(self._h, C.byref(b)))
        return int(b.value)

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.optmc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sync_stream(self):
        """Issue work on torch's current stream so torch-side copies/events order with our kernels."""
        if self.follow_torch_stream:
            ptr = self.torch.cuda.current_stream(self.tdev).cuda_stream
            L.check(self.lib.optmc_ctx_set_stream(self._h, C.c_void_p(ptr if ptr else 1)))  # 1 == cudaStreamLegacy

    def synchronize(self):
        L.check(self.lib.optmc_ctx_synchronize(self._h))

    def kernel_times(self):
        """(paths_ms, sweep_ms) of the last fused pricing call (CUDA events inside the library)."""
        a, b = C.c_double(), C.c_double()
        L.check(self.lib.optmc_ctx_kernel_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self) -> int:
        return int(self.lib.optmc_ctx_launch_count(self._h))

    # -- path simulation ---------------------------------------------------------------------------
    def alloc_slab(self, M: int, N: int, dtype):
        """Step-major slab [(N+1)][ld]; ld padded to 64 elements so every row is 256-byte aligned."""
        ld = (M + 63) // 64 * 64
        return self.torch.empty((N + 1, ld), dtype=_torch_dtype(_dtype_code(dtype)), device=self.tdev)

    def paths(self, model: ModelSpec, M: int, N: int, dtype="f32", rng: Optional[RngSpec] = None, out=None,
              return_v: bool = False):
        """Returns the slab view S[(N+1), M] (and V when return_v) -- om3:211-251 / om3:473-480 layout."""
        rng = rng or RngSpec()
        code = _dtype_code(dtype)
        self._sync_stream()
        S = out if out is not None else self.alloc_slab(M, N, dtype)
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1 and S.shape[0] >= N + 1
        ld = S.stride(0)
        mp, rp = model.c(), rng.c()
        V = None
        if model.model == L.MODEL_HESTON:
            if return_v:
                V = self.torch.empty_like(S)
            L.check(self.lib.optmc_paths_heston(self._h, C.byref(mp), C.byref(rp), M, N, code, S.data_ptr(),
                                                V.data_ptr() if V is not None else None, ld))
        else:
            L.check(self.lib.optmc_paths_gbm(self._h, C.byref(mp), C.byref(rp), M, N, code, S.data_ptr(), ld))
        if return_v:
            return S[: N + 1, :M], (V[: N + 1, :M] if V is not None else None)
        return S[: N + 1, :M]

    # -- local volatility: the IV network inside the step (om3:263-333) ---------------------------
    @staticmethod
    def _ivnet_c(net: dict, K: float):
        w = np.ascontiguousarray(net["weights"], dtype=np.float32)
        c = L.IvNet(int(net["hidden"]), int(net["layers"]), int(w.size), float(net.get("epsilon", 1e-4)),
                    w.ctypes.data_as(C.POINTER(C.c_float)), float(net["m_scale"]), float(net["tau_scale"]), float(K))
        return c, w  # keep `w` alive during the call

    def paths_localvol(self, S0, r, T, net: dict, K, M: int, N: int, dtype="f32", rng: Optional[RngSpec] = None, out=None):
        """simulate_local_vol_paths_antithetic (om3:300-333).  ``net`` = dict(hidden, layers, weights (the
        ImprovedIVNetwork state_dict flattened in its own order), m_scale, tau_scale, epsilon)."""
        rng = rng or RngSpec()
        S = out if out is not None else self.alloc_slab(M, N, dtype)
        mp = ModelSpec(L.MODEL_GBM, L.SCHEME_GBM_LOG_EULER, float(S0), float(r), float(T), sigma=0.0).c()
        rp = rng.c()
        c, keep = self._ivnet_c(net, K)
        self._sync_stream()
        L.check(self.lib.optmc_paths_localvol(self._h, C.byref(mp), C.byref(rp), C.byref(c), int(M), int(N), _dtype_code(dtype),
                                              S.data_ptr(), S.stride(0)))
        del keep
        return S[:, :M]

    def ivnet_sigma(self, net: dict, K, S_batch, tau: float):
        """IVModel.get_volatility_batch (om3:277-298) on the device: fp64 spots in, fp64 volatilities out."""
        t = self.torch
        Sd = t.as_tensor(np.ascontiguousarray(S_batch, dtype=np.float64)).to(self.tdev) if not t.is_tensor(S_batch) else S_batch.to(self.tdev, t.float64).contiguous()
        outd = t.empty_like(Sd)
        c, keep = self._ivnet_c(net, K)
        self._sync_stream()
        L.check(self.lib.optmc_ivnet_sigma(self._h, C.byref(c), float(tau), Sd.data_ptr(), int(Sd.numel()), outd.data_ptr()))
        del keep
        return outd

    def philox_normals(self, model_id: int, M: int, N: int, which: int = 0, dtype="f32",
                       rng: Optional[RngSpec] = None):
        rng = rng or RngSpec()
        code = _dtype_code(dtype)
        self._sync_stream()
        cols = M // 2 if rng.antithetic else M
        Z = self.torch.empty((N, cols), dtype=_torch_dtype(code), device=self.tdev)
        rp = rng.c()
        L.check(self.lib.optmc_philox_normals(self._h, C.byref(rp), model_id, M, N, which, code, Z.data_ptr()))
        return Z

    def philox_kat(self, ctr: Sequence[Sequence[int]], key: Sequence[Sequence[int]]) -> np.ndarray:
        n = len(ctr)
        c = np.ascontiguousarray(np.array(ctr, dtype=np.uint32).reshape(n, 4))
        k = np.ascontiguousarray(np.array(key, dtype=np.uint32).reshape(n, 2))
        o = np.zeros((n, 4), dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        self._sync_stream()
        L.check(self.lib.optmc_philox_kat(self._h, n, c.ctypes.data_as(u32p), k.ctypes.data_as(u32p),
                                          o.ctypes.data_as(u32p)))
        return o

    # -- LSM sweep ------------------------------------------------------------------------------------
    @staticmethod
    def _lsm_params(K, r, T, option_type, basis, semantics, impl) -> L.LsmParams:
        b = {"poly2": L.BASIS_POLY2, "poly3": L.BASIS_POLY3, "ref7": L.BASIS_REF7}.get(basis, basis)
        s = {"reference": L.SEM_REFERENCE, "textbook": L.SEM_TEXTBOOK}.get(semantics, semantics)
        i = {"auto": L.SWEEP_AUTO, "resident": L.SWEEP_RESIDENT, "split": L.SWEEP_SPLIT}.get(impl, impl)
        return L.LsmParams(float(K), float(r), float(T), 1 if option_type == "put" else 0, int(b), int(s), int(i))

    def _result_block(self, N: int, p: int, arrays: bool):
        res = L.LsmResult()
        keep = {}
        if arrays:
            keep["betas"] = np.full((N + 1, p), np.nan)
            keep["boundary"] = np.full(N + 1, np.nan)
            keep["ex_count"] = np.zeros(N + 1, dtype=np.int64)
            keep["n_itm"] = np.zeros(N + 1, dtype=np.int64)
            res.betas = keep["betas"].ctypes.data_as(C.POINTER(C.c_double))
            res.boundary = keep["boundary"].ctypes.data_as(C.POINTER(C.c_double))
            res.ex_count = keep["ex_count"].ctypes.data_as(C.POINTER(C.c_int64))
            res.n_itm = keep["n_itm"].ctypes.data_as(C.POINTER(C.c_int64))
        return res, keep

    @staticmethod
    def _to_result(res: L.LsmResult, keep) -> SweepResult:
        return SweepResult(res.price, res.stderr_, int(res.n_paths), int(res.impl_used), int(res.n_launches),
                           keep.get("betas"), keep.get("boundary"), keep.get("ex_count"), keep.get("n_itm"))

    def lsm(self, S, K, r, T, option_type="put", basis="poly2", semantics="reference", impl="auto", arrays=True,
            M: Optional[int] = None, asynchronous: bool = False):
        """Sweep an existing slab S[(N+1), M] (torch CUDA tensor, row stride = ld)."""
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, basis, semantics, impl)
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        self._sync_stream()
        if asynchronous:
            L.check(self.lib.optmc_lsm_poly(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp), None))
            return None
        p = 3 if lp.basis == L.BASIS_POLY2 else 4
        res, keep = self._result_block(N, p, arrays)
        L.check(self.lib.optmc_lsm_poly(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp), C.byref(res)))
        return self._to_result(res, keep)

    # -- path-sharded sweep: in-kernel exchange over NVLink peer memory (include/optmc.h, optmc_comm_*) -------
    def comm_export(self) -> bytes:
        buf = C.create_string_buffer(L.COMM_HANDLE_BYTES)
        L.check(self.lib.optmc_comm_export(self._h, buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, handles) -> None:
        blob = b"".join(handles)
        assert len(blob) == nranks * L.COMM_HANDLE_BYTES
        L.check(self.lib.optmc_comm_init(self._h, int(rank), int(nranks), C.c_char_p(blob)))

    def comm_finalize(self) -> None:
        L.check(self.lib.optmc_comm_finalize(self._h))

    def lsm_sharded(self, S_local, M_total: int, K, r, T, option_type="put", basis="poly2", semantics="reference",
                    arrays=False, M: Optional[int] = None):
        """This rank's part of a path-sharded sweep; every rank must call it (same order).  Returns the GLOBAL
        price / stderr / betas / n_itm; ex_count / boundary cover this rank's paths."""
        assert S_local.is_cuda and S_local.dim() == 2 and S_local.stride(1) == 1
        N = S_local.shape[0] - 1
        M = int(M if M is not None else S_local.shape[1])
        lp = self._lsm_params(K, r, T, option_type, basis, semantics, "resident")
        code = L.F64 if S_local.dtype == self.torch.float64 else L.F32
        p = 3 if lp.basis == L.BASIS_POLY2 else 4
        res, keep = self._result_block(N, p, arrays)
        self._sync_stream()
        L.check(self.lib.optmc_lsm_poly_sharded(self._h, S_local.data_ptr(), S_local.stride(0), M, int(M_total), N, code,
                                                C.byref(lp), C.byref(res)))
        return self._to_result(res, keep)

    def lsm_apply_policy(self, S, betas, K, r, T, option_type="put", basis="poly2", semantics="textbook", arrays=True,
                         M: Optional[int] = None) -> SweepResult:
        """Out-of-sample exercise: price the policy given by per-date ``betas`` (as returned by ``lsm``) on another slab."""
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, basis, semantics, "auto")
        p = 3 if lp.basis == L.BASIS_POLY2 else 4
        b = np.ascontiguousarray(betas, dtype=np.float64)
        assert b.shape == (N + 1, p)
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        res, keep = self._result_block(N, p, arrays)
        self._sync_stream()
        L.check(self.lib.optmc_lsm_apply_policy(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp),
                                                b.ctypes.data_as(C.POINTER(C.c_double)), C.byref(res)))
        return self._to_result(res, keep)

    def lsm_global(self, S, K, r, T, option_type="put", semantics="reference", arrays=True, M: Optional[int] = None):
        """Global-regression LSM (the reference's v3 structure, om3:482-651, with a linear model on the seven
        reference features).  Returns a dict: price, stderr, n_rows, rank, beta[7], boundary, ex_count."""
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, L.BASIS_REF7, semantics, "auto")
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        res = L.GlobalResult()
        bnd = exc = None
        if arrays:
            bnd = np.full(N + 1, np.nan)
            exc = np.zeros(N + 1, dtype=np.int64)
            res.boundary = bnd.ctypes.data_as(C.POINTER(C.c_double))
            res.ex_count = exc.ctypes.data_as(C.POINTER(C.c_int64))
        self._sync_stream()
        L.check(self.lib.optmc_lsm_global(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp), C.byref(res)))
        return dict(price=res.price, stderr=res.stderr_, n_paths=int(res.n_paths), n_rows=int(res.n_rows),
                    rank=int(res.rank), n_launches=int(res.n_launches), beta=np.array(list(res.beta)), boundary=bnd,
                    ex_count=exc)

    def lsm_mlp(self, S, K, r, T, option_type="put", semantics="reference", hidden=32, epochs=10, lr=1e-3, seed=42,
                arrays=True, M: Optional[int] = None) -> SweepResult:
        """Per-date neural-network LSM (om2:277-310): a fresh ContNet per exercise date, full-batch Adam."""
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, "poly2", semantics, "split")
        npar = L.MlpParams(int(hidden), int(epochs), float(lr), int(seed) & 0xFFFFFFFFFFFFFFFF)
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        res, keep = self._result_block(N, 3, arrays)
        self._sync_stream()
        L.check(self.lib.optmc_lsm_mlp(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp), C.byref(npar),
                                       C.byref(res)))
        return self._to_result(res, keep)

    def mlp_grad_debug(self, hidden: int, xs: np.ndarray, ys: np.ndarray, params: np.ndarray):
        """One full-batch gradient + forward of the ContNet on host rows (test aid).  -> (grads[P], cont[n])."""
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        params = np.ascontiguousarray(params, dtype=np.float32)
        P = 3 * hidden + hidden * hidden + hidden + 1
        assert params.size == P and xs.size == ys.size
        grads = np.zeros(P, dtype=np.float32)
        cont = np.zeros(xs.size, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        self._sync_stream()
        L.check(self.lib.optmc_mlp_grad_debug(self._h, int(hidden), xs.size, xs.ctypes.data_as(fp), ys.ctypes.data_as(fp),
                                              params.ctypes.data_as(fp), grads.ctypes.data_as(fp), cont.ctypes.data_as(fp)))
        return grads, cont

    GNET_PARAMS = 7 * 128 + 128 + 2 * (128 * 128 + 128) + 128 + 1  # SingleLSMNet(7, 128, 3): 34177

    def lsm_gnet(self, S, K, r, T, option_type="put", semantics="reference", variant="cpu", epochs=25, batch=None, lr=1e-3,
                 weight_decay=None, dropout=0.1, seed=42, inference_dropout=-1, arrays=True, M: Optional[int] = None,
                 init_params=None, return_params=False, M_total: Optional[int] = None, **over):
        """The reference's v3 algorithm with its own regressor (om3:482-651): one SingleLSMNet(7,128,3) trained on the
        rows of all dates (tcgen05), then the decision pass.  ``variant`` picks the defaults of the CPU file
        (om3:565-613: batch 256, Adam + L2 1e-5, ReduceLROnPlateau, patience 8, population std) or of the torch-GPU
        file (om3gpu:740-798: batch 8192, AdamW 1e-4, no scheduler, patience 3, sample std).  Returns a dict.
        ``M_total``: this rank's part of a PATH-SHARDED fit (optmc_lsm_gnet_sharded; every rank of the comm_init group
        calls with its block of paths): price / stderr / best_loss / params are the global ones, identical on every
        rank; ex_count / boundary cover this rank's paths (sharded.gnet_sharded combines them)."""
        assert S.is_cuda and S.dim() == 2 and S.stride(1) == 1
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, L.BASIS_REF7, semantics, "auto")
        cpu = variant == "cpu"
        gp = L.GnetParams(128, 3, int(epochs), int(batch if batch is not None else (256 if cpu else 8192)), float(lr),
                          float(weight_decay if weight_decay is not None else (1e-5 if cpu else 1e-4)), 0 if cpu else 1,
                          5 if cpu else 0, 0.5, 1e-6, 8 if cpu else 3, 0 if cpu else 1, 1e-6, float(dropout),
                          int(inference_dropout), 0, int(seed) & 0xFFFFFFFFFFFFFFFF, None, None)
        for k, v in over.items():
            setattr(gp, k, v)
        fp = C.POINTER(C.c_float)
        p_in = p_out = None
        if init_params is not None:  # warm start: the torch-GPU file keeps one network across calls (om3gpu:741-748)
            p_in = np.ascontiguousarray(init_params, dtype=np.float32)
            assert p_in.size == self.GNET_PARAMS
            gp.init_params = p_in.ctypes.data_as(fp)
        if return_params:
            p_out = np.zeros(self.GNET_PARAMS, dtype=np.float32)
            gp.final_params = p_out.ctypes.data_as(fp)
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        res = L.GnetResult()
        bnd = exc = None
        if arrays:
            bnd = np.full(N + 1, np.nan)
            exc = np.zeros(N + 1, dtype=np.int64)
            res.boundary = bnd.ctypes.data_as(C.POINTER(C.c_double))
            res.ex_count = exc.ctypes.data_as(C.POINTER(C.c_int64))
        self._sync_stream()
        if M_total is not None:
            L.check(self.lib.optmc_lsm_gnet_sharded(self._h, S.data_ptr(), S.stride(0), M, int(M_total), N, code, C.byref(lp),
                                                    C.byref(gp), C.byref(res)))
        else:
            L.check(self.lib.optmc_lsm_gnet(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp), C.byref(gp), C.byref(res)))
        return dict(price=res.price, stderr=res.stderr_, n_paths=int(res.n_paths), n_rows=int(res.n_rows),
                    epochs_run=int(res.epochs_run), n_launches=int(res.n_launches), best_loss=res.best_loss,
                    final_lr=res.final_lr, boundary=bnd, ex_count=exc, params=p_out)

    def gnet_shard_plan(self, n_rows, batch: int, b: int, rank: int):
        """(lo, hi, global_rows): rank's slice of its own shuffled rows in optimiser step b of a path-sharded fit, and
        the rows of all ranks in that step.  Host arithmetic only (optmc_gnet_shard_plan)."""
        return gnet_shard_plan(self.lib, n_rows, batch, b, rank)

    def gnet_grad_debug(self, feat: np.ndarray, ys: np.ndarray, params: np.ndarray):
        """MSE loss and gradient of SingleLSMNet(7,128,3) on host rows (normalised features [n,7]) -- test aid."""
        feat = np.ascontiguousarray(feat, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        params = np.ascontiguousarray(params, dtype=np.float32)
        assert feat.ndim == 2 and feat.shape[1] == 7 and ys.size == feat.shape[0] and params.size == self.GNET_PARAMS
        grads = np.zeros(self.GNET_PARAMS, dtype=np.float32)
        loss = C.c_float()
        fp = C.POINTER(C.c_float)
        self._sync_stream()
        L.check(self.lib.optmc_gnet_grad_debug(self._h, feat.shape[0], feat.ctypes.data_as(fp), ys.ctypes.data_as(fp),
                                               params.ctypes.data_as(fp), grads.ctypes.data_as(fp), C.byref(loss)))
        return grads, float(loss.value)

    def gnet_streams_debug(self, seed: int, epoch: int = 0, step: int = 1, dropout: float = 0.1, n_rows: int = 0, row_ids=None):
        """The shuffle of `epoch` over n_rows rows and the keep masks [len(row_ids), 3, 128] of optimiser step `step`
        (0 = the decision pass) as the device code evaluates them -- test aid for the paired training check."""
        perm = np.zeros(int(n_rows), dtype=np.int64)
        ids = np.ascontiguousarray(row_ids if row_ids is not None else np.zeros(0), dtype=np.uint32)
        words = np.zeros((ids.size, 3, 4), dtype=np.uint32)
        self._sync_stream()
        L.check(self.lib.optmc_gnet_streams_debug(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch), int(step), float(dropout),
                                                  perm.size, perm.ctypes.data_as(C.POINTER(C.c_int64)),
                                                  ids.ctypes.data_as(C.POINTER(C.c_uint32)), ids.size,
                                                  words.ctypes.data_as(C.POINTER(C.c_uint32))))
        keep = ((words[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(ids.size, 3, 128)
        return perm, keep

    def mlp_init_params(self, seed: int, date: int, hidden: int = 32) -> np.ndarray:
        n = 3 * hidden + hidden * hidden + hidden + 1
        out = np.zeros(n, dtype=np.float32)
        got = self.lib.optmc_mlp_init_params(hidden, int(seed) & 0xFFFFFFFFFFFFFFFF, int(date),
                                             out.ctypes.data_as(C.POINTER(C.c_float)))
        if got != n:
            L.check(got if got < 0 else L.EINVAL)
        return out

    def lsm_zero_cashflows(self) -> int:
        """Paths whose final cash-flow is exactly zero, after a split / network sweep (om1:168)."""
        n = C.c_int64()
        L.check(self.lib.optmc_lsm_zero_cashflows(self._h, C.byref(n)))
        return int(n.value)

    def lsm_fetch(self, N: int, basis="poly2", arrays=True) -> SweepResult:
        p = 3 if basis in ("poly2", L.BASIS_POLY2) else 4
        res, keep = self._result_block(N, p, arrays)
        L.check(self.lib.optmc_lsm_fetch(self._h, C.byref(res)))
        return self._to_result(res, keep)

    # per-date building blocks (path-sharded multi-GPU sweep; see sharded.py)
    def lsm_begin(self, S, K, r, T, option_type="put", basis="poly2", semantics="reference", M=None):
        N = S.shape[0] - 1
        M = int(M if M is not None else S.shape[1])
        lp = self._lsm_params(K, r, T, option_type, basis, semantics, "split")
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        self._sync_stream()
        L.check(self.lib.optmc_lsm_begin(self._h, S.data_ptr(), S.stride(0), M, N, code, C.byref(lp)))

    def lsm_gram_date(self, t: int, gram):
        self._sync_stream()
        L.check(self.lib.optmc_lsm_gram_date(self._h, t, gram.data_ptr()))

    def lsm_update_date(self, t: int, gram):
        self._sync_stream()
        L.check(self.lib.optmc_lsm_update_date(self._h, t, gram.data_ptr()))

    def lsm_finish(self, sums):
        self._sync_stream()
        L.check(self.lib.optmc_lsm_finish(self._h, sums.data_ptr()))

    def gram_len(self, basis="poly2") -> int:
        b = {"poly2": L.BASIS_POLY2, "poly3": L.BASIS_POLY3, "ref7": L.BASIS_REF7}.get(basis, basis)
        return int(self.lib.optmc_lsm_gram_len(b))

    # -- fused host-facing calls ---------------------------------------------------------------------
    def price_american(self, model: ModelSpec, M: int, N: int, K, option_type="put", dtype="f32",
                       rng: Optional[RngSpec] = None, basis="poly2", semantics="reference", impl="auto",
                       arrays=False, asynchronous=False):
        """price_american_enhanced_lsm (om3:439-651): host scalars in, host results out."""
        rng = rng or RngSpec()
        lp = self._lsm_params(K, model.r, model.T, option_type, basis, semantics, impl)
        mp, rp = model.c(), rng.c()
        self._sync_stream()
        if asynchronous:
            L.check(self.lib.optmc_price_american(self._h, C.byref(mp), C.byref(rp), M, N, _dtype_code(dtype),
                                                  C.byref(lp), None))
            return None
        p = 3 if lp.basis == L.BASIS_POLY2 else 4
        res, keep = self._result_block(N, p, arrays)
        L.check(self.lib.optmc_price_american(self._h, C.byref(mp), C.byref(rp), M, N, _dtype_code(dtype),
                                              C.byref(lp), C.byref(res)))
        return self._to_result(res, keep)

    def price_american_batch(self, model: ModelSpec, M: int, S0, K, T, N, is_put, dtype="f32",
                             rng: Optional[RngSpec] = None, basis="poly2", semantics="reference", streams=None,
                             details=False, european=False, M_total: int = 0):
        """Grid of American options in a few grouped launches (om3:697-713 / om3gpu:934-956 curve drivers,
        BASELINE config 4).  S0, K, T, N, is_put: scalars or arrays of length n.  -> (price[n], stderr[n]).

        details / european / M_total select optmc_price_american_batch_ex and return (price, stderr, extras) with
        extras = dict(ex_count, boundary, betas, n_itm [n, max N + 1 (, 4)], european [n, 2], shape [4]):
        per-date outputs of every option, the European leg on each option's own paths (control variate,
        om3:653-677), the launched sweep shape.  M_total > 0: path-sharded batch -- this rank holds M paths of EVERY
        option starting at pair rng.pair_offset, totals are exchanged inside the sweep kernel (comm_init first)."""
        rng = rng or RngSpec()
        arrs = np.broadcast_arrays(np.asarray(S0, dtype=np.float64), np.asarray(K, dtype=np.float64),
                                   np.asarray(T, dtype=np.float64), np.asarray(N, dtype=np.int64),
                                   np.asarray(is_put, dtype=np.int64))
        S0a, Ka, Ta, Na, Pa = (np.atleast_1d(a).ravel() for a in arrs)
        n = S0a.size
        sid = np.arange(n, dtype=np.uint64) if streams is None else np.asarray(streams, dtype=np.uint64).ravel()
        assert sid.size == n
        opts = (L.AmericanOption * n)()
        for i in range(n):
            opts[i] = L.AmericanOption(float(S0a[i]), float(Ka[i]), float(Ta[i]), int(Na[i]), int(Pa[i]), int(sid[i]))
        out = (L.PriceResult * n)()
        lp = self._lsm_params(1.0, model.r, 1.0, "put", basis, semantics, "auto")
        mp, rp = model.c(), rng.c()
        self._sync_stream()
        if not (details or european or M_total):
            L.check(self.lib.optmc_price_american_batch(self._h, C.byref(mp), C.byref(rp), int(M), _dtype_code(dtype),
                                                        int(lp.basis), int(lp.semantics), n, opts, out))
            return np.array([o.price for o in out]), np.array([o.stderr_ for o in out])
        ex = L.BatchExtras()
        n1 = int(Na.max()) + 1
        extras = {}
        if details:
            extras["ex_count"] = np.zeros((n, n1), dtype=np.int64)
            extras["n_itm"] = np.zeros((n, n1), dtype=np.int64)
            extras["boundary"] = np.full((n, n1), np.nan)
            extras["betas"] = np.full((n, n1, 4), np.nan)
            ex.ex_count = extras["ex_count"].ctypes.data_as(C.POINTER(C.c_int64))
            ex.n_itm = extras["n_itm"].ctypes.data_as(C.POINTER(C.c_int64))
            ex.boundary = extras["boundary"].ctypes.data_as(C.POINTER(C.c_double))
            ex.betas = extras["betas"].ctypes.data_as(C.POINTER(C.c_double))
            ex.ld_dates = n1
        if european:
            extras["european"] = np.zeros((n, 2))
            ex.european = extras["european"].ctypes.data_as(C.POINTER(C.c_double))
        ex.M_total = int(M_total)
        L.check(self.lib.optmc_price_american_batch_ex(self._h, C.byref(mp), C.byref(rp), int(M), _dtype_code(dtype),
                                                       int(lp.basis), int(lp.semantics), n, opts, out, C.byref(ex)))
        extras["shape"] = [int(v) for v in ex.shape]
        return np.array([o.price for o in out]), np.array([o.stderr_ for o in out]), extras

    def qmc_normals(self, M: int, N: int, factors: int = 1, bridge: bool = True, dtype="f64", shift_seed: Optional[int] = None,
                    pair_offset: int = 0):
        """Sobol' (Joe-Kuo) + Brownian-bridge normals as step-major device tensors [N, M/2] for
        RngSpec(z1=..., z2=...).  shift_seed: random digital shift (randomised QMC; None = the plain sequence)."""
        t = self.torch
        td = t.float64 if dtype == "f64" else t.float32
        Z1 = t.empty((N, M // 2), dtype=td, device=self.tdev)
        Z2 = t.empty((N, M // 2), dtype=td, device=self.tdev) if factors == 2 else None
        shift = None
        if shift_seed is not None:
            sh = np.random.default_rng(shift_seed).integers(0, 2**32, size=factors * N, dtype=np.uint64).astype(np.uint32)
            shift = sh.ctypes.data_as(C.POINTER(C.c_uint32))
        self._sync_stream()
        L.check(self.lib.optmc_qmc_normals(self._h, int(M), int(N), int(factors), 1 if bridge else 0, int(pair_offset), shift,
                                           _dtype_code(dtype), Z1.data_ptr(), Z2.data_ptr() if Z2 is not None else None))
        return (Z1, Z2) if factors == 2 else Z1

    def qmc_bridge_schedule(self, N: int):
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        idx, left, right = (np.zeros(N, dtype=np.int32) for _ in range(3))
        wl, wr, sd = (np.zeros(N) for _ in range(3))
        L.check(self.lib.optmc_qmc_bridge_schedule(int(N), idx.ctypes.data_as(ip), left.ctypes.data_as(ip), right.ctypes.data_as(ip),
                                                   wl.ctypes.data_as(dp), wr.ctypes.data_as(dp), sd.ctypes.data_as(dp)))
        return idx, left, right, wl, wr, sd

    def price_european_grid(self, model: ModelSpec, M: int, S0, K, T, N, is_put, dtype="f32",
                            rng: Optional[RngSpec] = None, stream_id=None):
        """Fused no-store European pricing with per-option spot / strike / maturity / step count (the independent
        control-variate leg of a whole curve, om3:653-713, as ONE launch).  -> (mean[n], stderr[n])."""
        rng = rng or RngSpec()
        arrs = np.broadcast_arrays(np.asarray(S0, dtype=np.float64), np.asarray(K, dtype=np.float64),
                                   np.asarray(T, dtype=np.float64), np.asarray(N, dtype=np.int32),
                                   np.asarray(is_put, dtype=np.int32))
        S0a, Ka, Ta, Na, Pa = (np.ascontiguousarray(np.atleast_1d(a).ravel()) for a in arrs)
        n = Ka.size
        out = (L.EuropeanResult * n)()
        mp, rp = model.c(), rng.c()
        self._sync_stream()
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        sid = None
        if stream_id is not None:
            sid_arr = np.ascontiguousarray(np.asarray(stream_id, dtype=np.int32))
            assert sid_arr.size == n
            sid = sid_arr.ctypes.data_as(ip)
        L.check(self.lib.optmc_price_european_grid(self._h, C.byref(mp), C.byref(rp), int(M), _dtype_code(dtype), n,
                                                   S0a.ctypes.data_as(dp), Ka.ctypes.data_as(dp), Ta.ctypes.data_as(dp),
                                                   Na.ctypes.data_as(ip), Pa.ctypes.data_as(ip), sid, out))
        return (np.array([o.mean for o in out]), np.array([o.stderr_ for o in out]))

    def price_european_batch(self, model: ModelSpec, M: int, N: int, K, T, is_put, dtype="f32",
                             rng: Optional[RngSpec] = None, stream_id=None):
        """Fused no-store European pricing of n options (om3:382-437, hc:259-281).  -> (mean[n], stderr[n])."""
        rng = rng or RngSpec()
        K = np.ascontiguousarray(np.atleast_1d(np.asarray(K, dtype=np.float64)))
        T = np.ascontiguousarray(np.atleast_1d(np.asarray(T, dtype=np.float64)))
        P = np.ascontiguousarray(np.atleast_1d(np.asarray(is_put, dtype=np.int32)))
        n = K.size
        assert T.size == n and P.size == n
        out = (L.EuropeanResult * n)()
        mp, rp = model.c(), rng.c()
        self._sync_stream()
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        sid = None
        if stream_id is not None:
            sid_arr = np.ascontiguousarray(np.asarray(stream_id, dtype=np.int32))
            assert sid_arr.size == n
            sid = sid_arr.ctypes.data_as(ip)
        L.check(self.lib.optmc_price_european_batch(self._h, C.byref(mp), C.byref(rp), M, N, _dtype_code(dtype), n,
                                                    K.ctypes.data_as(dp), T.ctypes.data_as(dp), P.ctypes.data_as(ip),
                                                    sid, out))
        return (np.array([o.mean for o in out]), np.array([o.stderr_ for o in out]))

    def european_from_slab(self, S_T, K, r, T, option_type="put"):
        assert S_T.is_cuda and S_T.is_contiguous()
        out = L.EuropeanResult()
        code = L.F64 if S_T.dtype == self.torch.float64 else L.F32
        self._sync_stream()
        L.check(self.lib.optmc_european_from_slab(self._h, S_T.data_ptr(), S_T.numel(), code, float(K), float(r),
                                                  float(T), 1 if option_type == "put" else 0, C.byref(out)))
        return out.mean, out.stderr_

    def features_ref7(self, S, K, r, T, t_current):
        """create_regression_features (om3:105-121) on the device -> [n, 7]."""
        assert S.is_cuda and S.is_contiguous()
        F = self.torch.empty((S.numel(), 7), dtype=S.dtype, device=S.device)
        code = L.F64 if S.dtype == self.torch.float64 else L.F32
        self._sync_stream()
        L.check(self.lib.optmc_features_ref7(self._h, S.data_ptr(), S.numel(), code, float(K), float(r), float(T),
                                             float(t_current), F.data_ptr()))
        return F


_DEFAULT = {}


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the compat layer's pricer objects share it)."""
    if device not in _DEFAULT:
        _DEFAULT[device] = Engine(device)
    return _DEFAULT[device]


def gnet_shard_plan(lib, n_rows, batch: int, b: int, rank: int):
    """optmc_gnet_shard_plan without a context (no device needed)."""
    arr = (C.c_int64 * len(n_rows))(*[int(x) for x in n_rows])
    lo, hi, g = C.c_int64(), C.c_int64(), C.c_int64()
    L.check(lib.optmc_gnet_shard_plan(arr, len(n_rows), int(batch), int(b), int(rank), C.byref(lo), C.byref(hi), C.byref(g)))
    return int(lo.value), int(hi.value), int(g.value)
