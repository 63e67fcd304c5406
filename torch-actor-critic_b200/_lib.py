"""ctypes binding of libsacb200.so (C ABI: include/sacb200.h).  No torch types cross this boundary: tensors are
handed over as `data_ptr()` integers and the current CUDA stream as a uintptr_t."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SACB200_LIB") or os.path.join(_HERE, "libsacb200.so")   # override: investigation builds

SACB_OK, SACB_ERR_INVALID, SACB_ERR_CUDA, SACB_ERR_STATE, SACB_ERR_SAMPLE, SACB_ERR_NCCL, SACB_ERR_PEER = 0, -1, -2, -3, -4, -5, -6
FLAG_CANONICAL_LOGPROB, FLAG_PI_FROM_STATES, FLAG_NO_ACTOR_GRAD_SYNC = 1, 2, 4


class SacbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32),
        ("hidden1", C.c_int32), ("hidden2", C.c_int32), ("batch_size", C.c_int32), ("max_burst", C.c_int32),
        ("log_min_std", C.c_float), ("log_max_std", C.c_float), ("act_limit", C.c_float), ("reserved0", C.c_int32),
        ("alpha", C.c_double), ("gamma", C.c_double), ("polyak", C.c_double), ("reward_scale", C.c_double),
        ("lr_pi", C.c_double), ("lr_q", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
        ("adam_eps", C.c_double), ("flags", C.c_uint32), ("seed", C.c_uint64), ("rank", C.c_int32), ("world", C.c_int32),
    ]


class SacbLayout(C.Structure):
    _fields_ = [
        ("actor_floats", C.c_int64), ("actor_w0", C.c_int64), ("actor_b0", C.c_int64), ("actor_w1", C.c_int64),
        ("actor_b1", C.c_int64), ("actor_w_mu", C.c_int64), ("actor_w_ls", C.c_int64), ("actor_b_mu", C.c_int64),
        ("actor_b_ls", C.c_int64), ("critic_floats", C.c_int64), ("critic_w0", C.c_int64 * 2), ("critic_b0", C.c_int64 * 2),
        ("critic_w1", C.c_int64 * 2), ("critic_b1", C.c_int64 * 2), ("critic_w2", C.c_int64 * 2), ("critic_b2", C.c_int64 * 2),
    ]


class SacvConfig(C.Structure):
    """sacv_config of include/sacb200.h (visual SAC engine)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32),
        ("hidden1", C.c_int32), ("hidden2", C.c_int32), ("batch_size", C.c_int32),
        ("vis_c", C.c_int32), ("vis_h", C.c_int32), ("vis_w", C.c_int32),
        ("filters", C.c_int32 * 3), ("kernels", C.c_int32 * 3), ("strides", C.c_int32 * 3), ("dense", C.c_int32),
        ("log_min_std", C.c_float), ("log_max_std", C.c_float), ("act_limit", C.c_float), ("precision", C.c_int32),
        ("alpha", C.c_double), ("gamma", C.c_double), ("polyak", C.c_double), ("reward_scale", C.c_double),
        ("lr_pi", C.c_double), ("lr_q", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double),
        ("flags", C.c_uint32), ("reserved0", C.c_uint32), ("seed", C.c_uint64),
    ]


class SacvLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("actor_floats", "a_l0w", "a_l0b", "a_l1w", "a_l1b", "a_cnn", "a_muw", "a_lsw", "a_mub", "a_lsb",
                                         "critic_floats", "critic_net_floats", "c_l0w", "c_l0b", "c_l1w", "c_l1b", "c_l2w", "c_l2b", "c_fw",
                                         "c_fb", "c_cnn", "cnn_floats")] + [
        ("cnn_w", C.c_int64 * 3), ("cnn_b", C.c_int64 * 3), ("cnn_lw", C.c_int64), ("cnn_lb", C.c_int64), ("cnn_fw", C.c_int64),
        ("cnn_fb", C.c_int64), ("conv_h", C.c_int32 * 3), ("conv_w", C.c_int32 * 3), ("flat", C.c_int32), ("reserved0", C.c_int32)]


class SacbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsacb200 error {code}: {msg}")
        self.code = code


_lib = None
_P, _I32, _I64, _U = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t       # _U: uintptr_t stream

_PROTOS = {
    "sacb_abi_version": (C.c_int, []),
    "sacb_last_error": (C.c_char_p, []),
    "sacb_layout": (C.c_int, [C.POINTER(SacbConfig), C.POINTER(SacbLayout)]),
    "sacb_create": (C.c_int, [C.POINTER(SacbConfig), C.POINTER(_P)]),
    "sacb_destroy": (C.c_int, [_P]),
    "sacb_bind_params": (C.c_int, [_P, _P, _P, _P]),
    "sacb_bind_adam": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I64]),
    "sacb_get_adam_steps": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "sacb_set_scalar": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "sacb_bind_buffer": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64]),
    "sacb_bind_buffer_packed": (C.c_int, [_P, _P, _I32, _I64]),
    "sacb_store": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _U]),
    "sacb_buffer_state": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "sacb_set_buffer_state": (C.c_int, [_P, _I64, _I64]),
    "sacb_sample": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _P, _U]),
    "sacb_update": (C.c_int, [_P, _I32, _P, _P, _P, _P, _U]),
    "sacb_load_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _U]),
    "sacb_update_critic": (C.c_int, [_P, _P, _P, _U]),
    "sacb_update_policy": (C.c_int, [_P, _P, _P, _U]),
    "sacb_update_targets": (C.c_int, [_P, _U]),
    "sacb_actor_forward": (C.c_int, [_P, _P, _I32, _P, _I32, _P, _P, _U]),
    "sacb_critic_forward": (C.c_int, [_P, _I32, _P, _P, _I32, _P, _P, _U]),
    "sacb_comm_unique_id": (C.c_int, [C.c_char_p, _P]),
    "sacb_comm_init": (C.c_int, [_P, C.c_char_p, _P, _I32, _I32]),
    "sacb_sync_params": (C.c_int, [_P, _U]),
    "sacb_peer_export": (C.c_int, [_P, _P, C.POINTER(_I64)]),
    "sacb_peer_bind": (C.c_int, [_P, _P, _I32, _I32]),
    "sacb_peer_status": (C.c_int, [_P, _U]),
    "sacb_debug_read": (C.c_int, [_P, C.c_char_p, _P, _I64, C.POINTER(_I64)]),
    "sacb_debug_profile": (C.c_int, [_P, _I32, _P, _I64, C.POINTER(_I32), C.POINTER(_I32), _U]),
    "sacb_launch_count": (_I64, [_P]),
    "sacb_set_mode": (C.c_int, [_P, _I32]),
    "sacb_set_precision": (C.c_int, [_P, _I32]),
    # ---- visual SAC engine (networks/convolutional.py + buffer/visual_replay_buffer.py path) ----
    "sacv_last_error": (C.c_char_p, []),
    "sacv_layout": (C.c_int, [C.POINTER(SacvConfig), C.POINTER(SacvLayout)]),
    "sacv_create": (C.c_int, [C.POINTER(SacvConfig), C.POINTER(_P)]),
    "sacv_destroy": (C.c_int, [_P]),
    "sacv_bind_params": (C.c_int, [_P, _P, _P, _P]),
    "sacv_bind_adam": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I64]),
    "sacv_get_adam_steps": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "sacv_set_scalar": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "sacv_bind_buffer": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64]),
    "sacv_buffer_state": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "sacv_set_buffer_state": (C.c_int, [_P, _I64, _I64]),
    "sacv_store": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _U]),
    "sacv_sample": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U]),
    "sacv_update": (C.c_int, [_P, _I32, _P, _P, _P, _P, _U]),
    "sacv_load_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U]),
    "sacv_update_critic": (C.c_int, [_P, _P, _P, _U]),
    "sacv_update_policy": (C.c_int, [_P, _P, _P, _U]),
    "sacv_update_targets": (C.c_int, [_P, _U]),
    "sacv_actor_forward": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _P, _P, _U]),
    "sacv_critic_forward": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P, _U]),
    "sacv_debug_read": (C.c_int, [_P, C.c_char_p, _P, _I64, C.POINTER(_I64)]),
    "sacv_launch_count": (_I64, [_P]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)


def load():
    """Load libsacb200.so.  Fails loudly (ImportError) when it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.sacb_abi_version() != 1:
        raise ImportError("libsacb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise SacbError(code, load().sacb_last_error().decode())
    return code


def checkv(code):
    """Status check for the sacv_* (visual engine) entry points."""
    if code != 0:
        raise SacbError(code, load().sacv_last_error().decode())
    return code


def nccl_library_path():
    """The libnccl.so.2 torch itself loads (so the engine shares it), or None for the default search path."""
    try:
        import nvidia.nccl as _n
        for base in list(getattr(_n, "__path__", [])):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return None
