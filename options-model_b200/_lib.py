"""ctypes binding of liboptmc.so -- one Python declaration per prototype in include/optmc.h."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# enums (include/optmc.h)
OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
F32, F64 = 0, 1
MODEL_GBM, MODEL_HESTON = 0, 1
COMM_HANDLE_BYTES = 64
SCHEME_GBM_LOG_EULER, SCHEME_GBM_LOGSPACE, SCHEME_HESTON_REF_ABSORB, SCHEME_HESTON_FULL_TRUNC, SCHEME_HESTON_REF_CALIB, \
    SCHEME_HESTON_QE = range(6)
BASIS_POLY2, BASIS_POLY3, BASIS_REF7 = 2, 3, 7
SEM_STICKY_MASK, SEM_REF_DISCOUNT = 1, 2
SEM_REFERENCE, SEM_TEXTBOOK = 3, 0
SWEEP_AUTO, SWEEP_RESIDENT, SWEEP_SPLIT = 0, 1, 2


class OptmcError(RuntimeError):
    pass


class ModelParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("scheme", C.c_int32), ("S0", C.c_double), ("r", C.c_double), ("T", C.c_double),
                ("sigma", C.c_double), ("v0", C.c_double), ("kappa", C.c_double), ("theta", C.c_double),
                ("xi", C.c_double), ("rho", C.c_double)]


class RngParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("stream", C.c_uint64), ("z1_dev", C.c_void_p), ("z2_dev", C.c_void_p),
                ("z_dtype", C.c_int32), ("antithetic", C.c_int32), ("pair_offset", C.c_int64)]


class LsmParams(C.Structure):
    _fields_ = [("K", C.c_double), ("r", C.c_double), ("T", C.c_double), ("is_put", C.c_int32), ("basis", C.c_int32),
                ("semantics", C.c_uint32), ("impl", C.c_int32)]


class LsmResult(C.Structure):
    _fields_ = [("price", C.c_double), ("stderr_", C.c_double), ("n_paths", C.c_int64), ("impl_used", C.c_int32),
                ("n_launches", C.c_int32), ("betas", C.POINTER(C.c_double)), ("boundary", C.POINTER(C.c_double)),
                ("ex_count", C.POINTER(C.c_int64)), ("n_itm", C.POINTER(C.c_int64))]


class AmericanOption(C.Structure):
    _fields_ = [("S0", C.c_double), ("K", C.c_double), ("T", C.c_double), ("N", C.c_int32), ("is_put", C.c_int32),
                ("stream", C.c_uint64)]


class PriceResult(C.Structure):
    _fields_ = [("price", C.c_double), ("stderr_", C.c_double)]


class BatchExtras(C.Structure):  # optmc_batch_extras
    _fields_ = [("ex_count", C.POINTER(C.c_int64)), ("boundary", C.POINTER(C.c_double)), ("betas", C.POINTER(C.c_double)),
                ("n_itm", C.POINTER(C.c_int64)), ("ld_dates", C.c_int32), ("reserved", C.c_int32),
                ("european", C.POINTER(C.c_double)), ("shape", C.c_int32 * 4), ("M_total", C.c_int64)]


class MlpParams(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("epochs", C.c_int32), ("lr", C.c_double), ("seed", C.c_uint64)]


class IvNet(C.Structure):  # optmc_ivnet
    _fields_ = [("hidden", C.c_int32), ("layers", C.c_int32), ("n_weights", C.c_int32), ("epsilon", C.c_float),
                ("weights", C.POINTER(C.c_float)), ("m_scale", C.c_double), ("tau_scale", C.c_double), ("K", C.c_double)]


class GnetParams(C.Structure):  # optmc_gnet_params
    _fields_ = [("hidden", C.c_int32), ("layers", C.c_int32), ("epochs", C.c_int32), ("batch", C.c_int32),
                ("lr", C.c_double), ("weight_decay", C.c_double), ("decoupled_wd", C.c_int32), ("sched_patience", C.c_int32),
                ("sched_factor", C.c_double), ("min_lr", C.c_double), ("stop_patience", C.c_int32), ("target_ddof", C.c_int32),
                ("min_delta", C.c_double), ("dropout", C.c_double), ("inference_dropout", C.c_int32), ("per_date", C.c_int32),
                ("seed", C.c_uint64), ("init_params", C.POINTER(C.c_float)), ("final_params", C.POINTER(C.c_float))]


class GnetResult(C.Structure):  # optmc_gnet_result
    _fields_ = [("price", C.c_double), ("stderr_", C.c_double), ("n_paths", C.c_int64), ("n_rows", C.c_int64),
                ("epochs_run", C.c_int32), ("n_launches", C.c_int32), ("best_loss", C.c_double), ("final_lr", C.c_double),
                ("boundary", C.POINTER(C.c_double)), ("ex_count", C.POINTER(C.c_int64))]


class GlobalResult(C.Structure):
    _fields_ = [("price", C.c_double), ("stderr_", C.c_double), ("n_paths", C.c_int64), ("n_rows", C.c_int64),
                ("n_launches", C.c_int32), ("rank", C.c_int32), ("beta", C.c_double * 7),
                ("boundary", C.POINTER(C.c_double)), ("ex_count", C.POINTER(C.c_int64))]


class EuropeanResult(C.Structure):
    _fields_ = [("mean", C.c_double), ("stderr_", C.c_double), ("n_paths", C.c_int64)]


# name -> (restype, argtypes); must list every symbol include/optmc.h declares (tests check this)
_P = C.POINTER
PROTOTYPES = {
    "optmc_abi_version": (C.c_int, []),
    "optmc_last_error": (C.c_char_p, []),
    "optmc_ctx_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "optmc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "optmc_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "optmc_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "optmc_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "optmc_ctx_kernel_times": (C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_double)]),
    "optmc_ctx_device_info": (C.c_int, [C.c_void_p, _P(C.c_int64)]),
    "optmc_workspace_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P(C.c_int64)]),
    "optmc_ctx_workspace_bytes": (C.c_int, [C.c_void_p, _P(C.c_int64)]),
    "optmc_paths_gbm": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_int64]),
    "optmc_paths_heston": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_int64]),
    "optmc_paths_localvol": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), _P(IvNet), C.c_int64, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int64]),
    "optmc_ivnet_sigma": (C.c_int, [C.c_void_p, _P(IvNet), C.c_double, C.c_void_p, C.c_int64, C.c_void_p]),
    "optmc_philox_normals": (C.c_int, [C.c_void_p, _P(RngParams), C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "optmc_philox_kat": (C.c_int, [C.c_void_p, C.c_int32, _P(C.c_uint32), _P(C.c_uint32), _P(C.c_uint32)]),
    "optmc_lsm_poly": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams),
                                 _P(LsmResult)]),
    "optmc_lsm_fetch": (C.c_int, [C.c_void_p, _P(LsmResult)]),
    "optmc_lsm_zero_cashflows": (C.c_int, [C.c_void_p, _P(C.c_int64)]),
    "optmc_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "optmc_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "optmc_comm_finalize": (C.c_int, [C.c_void_p]),
    "optmc_lsm_poly_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                         _P(LsmParams), _P(LsmResult)]),
    "optmc_lsm_global": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams),
                                   _P(GlobalResult)]),
    "optmc_lsm_mlp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams),
                                _P(MlpParams), _P(LsmResult)]),
    "optmc_mlp_init_params": (C.c_int, [C.c_int32, C.c_uint64, C.c_int32, _P(C.c_float)]),
    "optmc_mlp_grad_debug": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, _P(C.c_float), _P(C.c_float), _P(C.c_float),
                                       _P(C.c_float), _P(C.c_float)]),
    "optmc_lsm_apply_policy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams),
                                         _P(C.c_double), _P(LsmResult)]),
    "optmc_lsm_gnet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams),
                                 _P(GnetParams), _P(GnetResult)]),
    "optmc_lsm_gnet_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                         _P(LsmParams), _P(GnetParams), _P(GnetResult)]),
    "optmc_gnet_shard_plan": (C.c_int, [_P(C.c_int64), C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P(C.c_int64), _P(C.c_int64),
                                        _P(C.c_int64)]),
    "optmc_gnet_grad_debug": (C.c_int, [C.c_void_p, C.c_int64, _P(C.c_float), _P(C.c_float), _P(C.c_float), _P(C.c_float),
                                        _P(C.c_float)]),
    "optmc_gnet_streams_debug": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int64, _P(C.c_int64),
                                           _P(C.c_uint32), C.c_int64, _P(C.c_uint32)]),
    "optmc_lsm_gram_len": (C.c_int, [C.c_int32]),
    "optmc_lsm_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P(LsmParams)]),
    "optmc_lsm_gram_date": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "optmc_lsm_update_date": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "optmc_lsm_finish": (C.c_int, [C.c_void_p, C.c_void_p]),
    "optmc_price_american": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32, C.c_int32,
                                       _P(LsmParams), _P(LsmResult)]),
    "optmc_price_american_batch": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32, C.c_int32,
                                             C.c_uint32, C.c_int32, _P(AmericanOption), _P(PriceResult)]),
    "optmc_price_american_batch_ex": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32,
                                                C.c_int32, C.c_uint32, C.c_int32, _P(AmericanOption), _P(PriceResult),
                                                _P(BatchExtras)]),
    "optmc_price_european_batch": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32,
                                             C.c_int32, C.c_int32, _P(C.c_double), _P(C.c_double), _P(C.c_int32),
                                             _P(C.c_int32), _P(EuropeanResult)]),
    "optmc_qmc_normals": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P(C.c_uint32), C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "optmc_qmc_bridge_schedule": (C.c_int, [C.c_int32, _P(C.c_int32), _P(C.c_int32), _P(C.c_int32), _P(C.c_double),
                                            _P(C.c_double), _P(C.c_double)]),
    "optmc_price_european_grid": (C.c_int, [C.c_void_p, _P(ModelParams), _P(RngParams), C.c_int64, C.c_int32, C.c_int32,
                                            _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_int32), _P(C.c_int32),
                                            _P(C.c_int32), _P(EuropeanResult)]),
    "optmc_european_from_slab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                           C.c_double, C.c_int32, _P(EuropeanResult)]),
    "optmc_features_ref7": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_void_p]),
}


def library_path() -> str:
    return os.environ.get("OPTMC_LIB", os.path.join(HERE, "liboptmc.so"))


def load_library():
    """Load liboptmc.so (built in-tree by __graft_entry__.build() / make).  No fallback of any kind."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise OptmcError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         f"(or `make -C options-model_b200`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.optmc_abi_version() != 1:
        raise OptmcError("liboptmc.so ABI version mismatch; rebuild")
    _LIB = lib
    return lib


def check(rc: int):
    """Map a C status to the reference's exception types (om3:447-452 raise ValueError)."""
    if rc == OK:
        return
    msg = load_library().optmc_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise OptmcError(msg)
