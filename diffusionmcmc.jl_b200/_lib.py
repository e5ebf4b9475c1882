"""ctypes binding of libdmcmc_b200.so (include/dmcmc.h).  Fails loudly when the CUDA library
is missing: there is no CPU fallback for the hot path."""
import ctypes as C
import os

from . import build as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_lp = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("model", C.c_int32), ("d", C.c_int32),
                ("n_chains", C.c_int32), ("chain_offset", C.c_int32), ("interval_offset", C.c_int32),
                ("device", C.c_int32),
                ("seed", C.c_uint64), ("pin_eps", C.c_double), ("store_noise", C.c_int32)]


ABI_VERSION = 5


class Adapt(C.Structure):
    _fields_ = [("target_accpt_rate", C.c_double), ("scale", C.c_double), ("min", C.c_double), ("max", C.c_double),
                ("offset", C.c_double), ("adapt_every_k_steps", C.c_int32)]


# name -> (restype, argtypes); every symbol include/dmcmc.h declares
SYMBOLS = {
    "dmcmc_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "dmcmc_destroy": (None, [C.c_void_p]),
    "dmcmc_last_error": (C.c_char_p, [C.c_void_p]),
    "dmcmc_set_grid": (C.c_int, [C.c_void_p, C.c_int32, _ip, _ip, _dp]),
    "dmcmc_set_obs": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp]),
    "dmcmc_set_start": (C.c_int, [C.c_void_p, _dp]),
    "dmcmc_set_theta": (C.c_int, [C.c_void_p, _dp, C.c_int32]),
    "dmcmc_set_aux": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp]),
    "dmcmc_set_layout": (C.c_int, [C.c_void_p, C.c_int32, _ip, _ip, _ip]),
    "dmcmc_chequerboard_layout": (C.c_int, [C.c_int32, _ip, C.c_int32, C.c_int32, _ip, _ip, _ip]),
    "dmcmc_set_rho": (C.c_int, [C.c_void_p, _dp]),
    "dmcmc_backward_filter": (C.c_int, [C.c_void_p, C.c_int32]),
    "dmcmc_upload_tables": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]),
    "dmcmc_download_tables": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]),
    "dmcmc_init_paths": (C.c_int, [C.c_void_p, _dp]),
    "dmcmc_init_ll": (C.c_int, [C.c_void_p]),
    "dmcmc_impute": (C.c_int, [C.c_void_p, C.c_uint32, _dp, _dp, _dp, _dp, _bp]),
    "dmcmc_run_sweeps": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, _dp,
                                   C.c_int32, _dp, _dp, _bp]),
    "dmcmc_set_adaptation": (C.c_int, [C.c_void_p, C.POINTER(Adapt), C.c_int32, _dp, _dp]),
    "dmcmc_get_rho": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "dmcmc_host_alloc": (C.c_void_p, [C.c_size_t]),
    "dmcmc_host_free": (None, [C.c_void_p]),
    "dmcmc_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "dmcmc_host_unregister": (C.c_int, [C.c_void_p]),
    "dmcmc_save_path_async": (C.c_int, [C.c_void_p, C.c_int32, _ip, C.c_int32, _ip]),
    "dmcmc_save_path_wait": (C.c_int, [C.c_void_p, C.c_int32, _dp]),
    "dmcmc_relayout": (C.c_int, [C.c_void_p, C.c_int32, _ip, _ip, _ip]),
    "dmcmc_param_loglik": (C.c_int, [C.c_void_p, _dp, C.c_int32, C.c_int32, _dp, _bp]),
    "dmcmc_param_commit": (C.c_int, [C.c_void_p, C.c_int32]),
    "dmcmc_download_path": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "dmcmc_download_state": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp]),
    "dmcmc_set_block_mask": (C.c_int, [C.c_void_p, _ip]),
    "dmcmc_download_segment": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "dmcmc_upload_segment": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp]),
    "dmcmc_segment_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32]),
    "dmcmc_set_start_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "dmcmc_stream": (C.c_void_p, [C.c_void_p]),
    "dmcmc_impute_async": (C.c_int, [C.c_void_p, C.c_uint32]),
    "dmcmc_comm_unique_id": (C.c_int, [C.c_void_p]),
    "dmcmc_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "dmcmc_comm_destroy": (C.c_int, [C.c_void_p]),
    "dmcmc_halo_exchange": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dmcmc_comm_enable_p2p": (C.c_int, [C.c_void_p, C.c_int64]),
    "dmcmc_impute_exchange_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dmcmc_comm_p2p_enabled": (C.c_int, [C.c_void_p]),
    "dmcmc_param_loglik_sharded": (C.c_int, [C.c_void_p, _dp, C.c_int32, C.c_int32, C.c_int32, _ip, _dp, _dp, _bp]),
    "dmcmc_conjugate_dim": (C.c_int32, [C.c_void_p, _ip]),
    "dmcmc_conjugate_sums": (C.c_int, [C.c_void_p, _dp, _dp]),
    "dmcmc_n_blocks": (C.c_int32, [C.c_void_p]),
    "dmcmc_get_ll": (C.c_int, [C.c_void_p, _dp]),
    "dmcmc_set_ll": (C.c_int, [C.c_void_p, _dp]),
    "dmcmc_accept_counts": (C.c_int, [C.c_void_p, _lp, _lp]),
    "dmcmc_kernel_times": (C.c_int, [C.c_void_p, _dp, _lp]),
    "dmcmc_device_bytes": (C.c_int64, [C.c_void_p]),
    "dmcmc_fp64_peak": (C.c_double, [C.c_int32]),
    "dmcmc_fp64_tensor_peak": (C.c_double, [C.c_int32]),
    "dmcmc_set_threads": (C.c_int, [C.c_void_p, C.c_int32]),
    "dmcmc_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "dmcmc_launch_count": (C.c_int64, [C.c_void_p]),
    "dmcmc_timer_start": (C.c_int, [C.c_void_p]),
    "dmcmc_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "dmcmc_set_fast_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "dmcmc_set_cooperative": (C.c_int, [C.c_void_p, C.c_int32]),
}

_lib = None


def load():
    """Load the in-tree CUDA library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: run `python __graft_entry__.py` (build()) first. "
            "The imputation path is CUDA-only; there is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError when the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
