"""ctypes binding of ``libdqn_b200.so`` (``include/dqn_b200.h``).

There is deliberately no fallback: if the CUDA library has not been built (``__graft_entry__.build()``
or ``make -C deep-q-learning_b200/csrc``) importing a compute entry point raises, and if no sm_100a
device is present ``dqn_create`` fails with ``DQN_E_ARCH``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DQN_B200_LIB") or os.path.join(_HERE, "csrc", "libdqn_b200.so")   # env override: A/B kernel variants only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dqn_b200.h")

DQN_OPT_ADAM, DQN_OPT_ADAMW = 0, 1
DQN_PARAMS_ONLINE, DQN_PARAMS_TARGET = 0, 1
LOSS_KINDS = {"huber": 0, "l2": 1, "mse": 1}     # "huber" = the reference (q_learning_functions.py:36); "l2"/"mse" = 0.5 e^2, an extension
DQN_STEP_AUTO, DQN_STEP_CTA, DQN_STEP_CLUSTER, DQN_STEP_CTA_TC = 0, 1, 2, 3
STEP_KERNELS = {"auto": DQN_STEP_AUTO, "cta": DQN_STEP_CTA, "cluster": DQN_STEP_CLUSTER, "cta_tc": DQN_STEP_CTA_TC}
DQN_MAX_BATCH, DQN_MAX_OBS_DIM, DQN_MAX_ACTIONS = 1024, 16, 7
LOSS_RING = 4096


class DqnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdqn_b200 error {code}: {msg}")
        self.code = code


class DqnConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("n_agents", C.c_int32),
        ("obs_dim", C.c_int32), ("num_actions", C.c_int32), ("hidden1", C.c_int32),
        ("hidden2", C.c_int32), ("batch_size", C.c_int32), ("buffer_size", C.c_int64),
        ("gamma", C.c_float), ("opt_kind", C.c_int32),
        ("lr", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float),
        ("eps_root", C.c_float), ("weight_decay", C.c_float),
        ("seed", C.c_uint64), ("agent_id_base", C.c_int32), ("step_kernel", C.c_int32), ("stream", C.c_void_p), ("arena", C.c_void_p), ("arena_bytes", C.c_uint64),
    ]


class DqnHparams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("batch_size", C.c_int32), ("lr", C.c_float), ("b1", C.c_float),
                ("b2", C.c_float), ("eps", C.c_float), ("eps_root", C.c_float), ("weight_decay", C.c_float)]


class DqnEpisodeConfig(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("epsilon_decay_rate", C.c_double), ("min_epsilon", C.c_double),
                ("reward_to_reach", C.c_double), ("max_episodes", C.c_int32), ("max_steps", C.c_int32),
                ("training_start", C.c_int32), ("train_frequency", C.c_int32), ("replace_frequency", C.c_int32),
                ("reserved", C.c_int32)]


class DqnEpisodeState(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("average_reward", C.c_double), ("last_episode_reward", C.c_double),
                ("episode_reward", C.c_double), ("step_count", C.c_int64), ("policy_calls", C.c_int64),
                ("episode", C.c_int32), ("step_in_episode", C.c_int32), ("window_len", C.c_int32), ("finished", C.c_int32)]


class DqnCounters(C.Structure):
    _fields_ = [("ring_counter", C.c_int64), ("train_steps", C.c_int64), ("adam_count", C.c_int32), ("reserved", C.c_int32),
                ("pb1", C.c_double), ("pb2", C.c_double)]


class DqnDebugTaps(C.Structure):
    _fields_ = [("indices", C.c_void_p), ("q", C.c_void_p), ("next_q", C.c_void_p), ("next_q_tm", C.c_void_p),
                ("max_actions", C.c_void_p), ("targets", C.c_void_p), ("loss", C.c_void_p), ("grads", C.c_void_p)]


class DqnLbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("obs_dim", C.c_int32), ("num_actions", C.c_int32),
        ("hidden1", C.c_int32), ("hidden2", C.c_int32), ("batch_local", C.c_int32), ("gemm_mode", C.c_int32),
        ("buffer_size", C.c_int64), ("gamma", C.c_float), ("opt_kind", C.c_int32),
        ("lr", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float), ("eps_root", C.c_float),
        ("weight_decay", C.c_float), ("seed", C.c_uint64), ("rank", C.c_int32), ("world", C.c_int32),
        ("stream", C.c_void_p), ("arena", C.c_void_p), ("arena_bytes", C.c_uint64),
    ]


_H = C.c_void_p          # dqn_handle*
_P = C.c_void_p          # any data pointer
_i32, _i64 = C.c_int32, C.c_int64

# name -> (restype, argtypes); one entry per function declared in include/dqn_b200.h
PROTOTYPES = {
    "dqn_abi_version": (C.c_int, []),
    "dqn_last_error": (C.c_char_p, []),
    "dqn_arena_bytes": (C.c_int, [C.POINTER(DqnConfig), C.POINTER(C.c_uint64)]),
    "dqn_create": (C.c_int, [C.POINTER(DqnConfig), C.POINTER(_H)]),
    "dqn_destroy": (C.c_int, [_H]),
    "dqn_param_count": (C.c_int, [_H, C.POINTER(_i32)]),
    "dqn_synchronize": (C.c_int, [_H]),
    "dqn_set_params": (C.c_int, [_H, _i32, _i32, _P, _i32]),
    "dqn_get_params": (C.c_int, [_H, _i32, _i32, _P, _i32]),
    "dqn_set_opt_state": (C.c_int, [_H, _i32, _i32, _P, _P, _i32]),
    "dqn_get_opt_state": (C.c_int, [_H, _i32, C.POINTER(_i32), _P, _P, _i32]),
    "dqn_set_hparams": (C.c_int, [_H, _i32, C.POINTER(DqnHparams)]),
    "dqn_get_hparams": (C.c_int, [_H, _i32, C.POINTER(DqnHparams)]),
    "dqn_set_step_kernel": (C.c_int, [_H, _i32]),
    "dqn_set_session": (C.c_int, [_H, _i32]),
    "dqn_get_counters": (C.c_int, [_H, _i32, C.POINTER(DqnCounters)]),
    "dqn_set_counters": (C.c_int, [_H, _i32, C.POINTER(DqnCounters)]),
    "dqn_store": (C.c_int, [_H, _i32, _i64, _P, _P, _P, _P, _P]),
    "dqn_store_device": (C.c_int, [_H, _i32, _i64, _P, _P, _P, _P, _P]),
    "dqn_buffer_state": (C.c_int, [_H, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "dqn_buffer_export": (C.c_int, [_H, _i32, _P, _P, _P, _P, _P]),
    "dqn_sample_indices": (C.c_int, [_H, _i32, _i64, _i32, _P]),
    "dqn_sample_batch": (C.c_int, [_H, _i32, _P, _i64, _i32, _P, _P, _P, _P, _P]),
    "dqn_sample_batch_device": (C.c_int, [_H, _i32, _P, _i64, _i32, _P, _P, _P, _P, _P]),
    "dqn_train_step": (C.c_int, [_H, _i32, _i32, _i32, _P, C.POINTER(DqnDebugTaps)]),
    "dqn_store_train_step": (C.c_int, [_H, _i32, _i64, _P, _P, _P, _P, _P, _i32, _P]),
    "dqn_train_step_device_idx": (C.c_int, [_H, _i32, _i32, _i32, _P]),
    "dqn_get_losses": (C.c_int, [_H, _i32, _i32, _P, C.POINTER(_i64)]),
    "dqn_get_loss_lagged": (C.c_int, [_H, _i32, _i32, _P]),
    "dqn_sync_target": (C.c_int, [_H, _i32, _i32]),
    "dqn_polyak_target": (C.c_int, [_H, _i32, _i32, C.c_float]),
    "dqn_set_loss_kind": (C.c_int, [_H, _i32, _i32, _i32]),
    "dqn_act": (C.c_int, [_H, _i32, _P, C.POINTER(_i32)]),
    "dqn_act_batch": (C.c_int, [_H, _i32, _i32, _P, _P]),
    # episode-loop control on the device
    "dqn_episode_configure": (C.c_int, [_H, _i32, _i32, C.POINTER(DqnEpisodeConfig), _i32]),
    "dqn_policy_batch": (C.c_int, [_H, _i32, _i32, _P, _P]),
    "dqn_observe_batch": (C.c_int, [_H, _i32, _i32, _P, _P, _P, _P, _P, _P]),
    "dqn_train_flagged": (C.c_int, [_H, _i32, _i32]),
    "dqn_episode_get_state": (C.c_int, [_H, _i32, C.POINTER(DqnEpisodeState)]),
    # large-batch data-parallel mode
    "dqn_lb_arena_bytes": (C.c_int, [C.POINTER(DqnLbConfig), C.POINTER(C.c_uint64)]),
    "dqn_lb_create": (C.c_int, [C.POINTER(DqnLbConfig), C.POINTER(_H)]),
    "dqn_lb_destroy": (C.c_int, [_H]),
    "dqn_lb_param_count": (C.c_int, [_H, C.POINTER(_i32)]),
    "dqn_lb_set_params": (C.c_int, [_H, _i32, _P, _i32]),
    "dqn_lb_get_params": (C.c_int, [_H, _i32, _P, _i32]),
    "dqn_lb_set_opt_state": (C.c_int, [_H, _i32, _P, _P, _i32]),
    "dqn_lb_get_opt_state": (C.c_int, [_H, C.POINTER(_i32), _P, _P, _i32]),
    "dqn_lb_store": (C.c_int, [_H, _i64, _P, _P, _P, _P, _P]),
    "dqn_lb_store_device": (C.c_int, [_H, _i64, _P, _P, _P, _P, _P]),
    "dqn_lb_buffer_state": (C.c_int, [_H, C.POINTER(_i64), C.POINTER(_i64)]),
    "dqn_lb_forward_backward": (C.c_int, [_H, _P, _i32]),
    "dqn_lb_grads": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(_i64)]),
    "dqn_lb_comm_init": (C.c_int, [_H, _P, C.POINTER(C.c_void_p)]),
    "dqn_lb_comm_connect": (C.c_int, [_H, _P, C.POINTER(C.c_void_p)]),
    "dqn_lb_allreduce": (C.c_int, [_H]),
    "dqn_lb_apply": (C.c_int, [_H]),
    "dqn_lb_train_step": (C.c_int, [_H, _P, _i32]),
    "dqn_lb_sync_target": (C.c_int, [_H]),
    "dqn_lb_polyak_target": (C.c_int, [_H, C.c_float]),
    "dqn_lb_set_loss_kind": (C.c_int, [_H, _i32]),
    "dqn_lb_get_loss": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "dqn_lb_debug_read": (C.c_int, [_H, _i32, _P, C.c_uint64]),
    "dqn_lb_synchronize": (C.c_int, [_H]),
    # prioritized replay
    "dqn_per_arena_bytes": (C.c_int, [_i64, C.POINTER(C.c_uint64)]),
    "dqn_per_create": (C.c_int, [_i32, _i64, C.c_float, C.c_float, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(_H)]),
    "dqn_per_destroy": (C.c_int, [_H]),
    "dqn_per_update": (C.c_int, [_H, _P, _P, _i32, _i32]),
    "dqn_per_update_host": (C.c_int, [_H, _P, _P, _i32, _i32]),
    "dqn_per_fill": (C.c_int, [_H, _P, _i64]),
    "dqn_per_sample": (C.c_int, [_H, _i64, _i32, _P, _P]),
    "dqn_per_sample_host": (C.c_int, [_H, _i64, _i32, _P, _P]),
    "dqn_per_total": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "dqn_per_read_nodes": (C.c_int, [_H, _i64, _i64, _P]),
    "dqn_per_leaf_base": (C.c_int, [_H, C.POINTER(_i64)]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
            "g.build()').  There is no CPU/PyTorch fallback for the DQN hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)            # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.dqn_abi_version() != 1:
        raise ImportError("libdqn_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = _lib.dqn_last_error() if _lib is not None else b""
        raise DqnError(rc, (msg or b"").decode("utf-8", "replace"))


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    return None if a is None else a.ctypes.data
