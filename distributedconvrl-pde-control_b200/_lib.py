"""ctypes binding of libpdeb200.so (include/pdeb200.h).

There is NO CPU fallback: if the shared library is missing or no B200 is
visible, construction fails loudly.
"""
import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libpdeb200.so"

OK = 0
ABI_VERSION = 2
COMM_NONE, COMM_NCCL, COMM_PEER = 0, 1, 2
UNIQUE_ID_BYTES = 128
F32, F64 = 0, 1
KS, KSEG1D, NS2D, KSEG2D = 0, 1, 2, 3
CHECK_NONE, CHECK_Y, CHECK_REWARD = 0, 1, 2
ACT_IDENTITY, ACT_RELU, ACT_TANH = 0, 1, 2
NET_BEHAVIOR_ACTOR, NET_BEHAVIOR_CRITIC, NET_TARGET_ACTOR, NET_TARGET_CRITIC = 0, 1, 2, 3
(ARR_Y, ARR_P, ARR_STATE, ARR_ACTION, ARR_DELTA_ACTION, ARR_REWARD, ARR_DONE, ARR_TIME, ARR_STEPS, ARR_Y0,
 ARR_GRADS, ARR_LOSSES, ARR_SENSORS, ARR_ACTION_IN, ARR_STATS, ARR_NSUB) = range(16)


class Config(C.Structure):
    """struct pdeb200_config (include/pdeb200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "problem", "dtype", "nx", "ny", "n_envs", "n_sensors", "n_actuators", "window_size",
        "temporal_steps", "memory_size", "oversampling", "check_max_value", "mono", "sensors_per_axis", "ifpad")] + \
        [(n, C.c_double) for n in (
            "Lx", "Ly", "dt", "te", "t0", "mu", "nu", "agent_power", "max_value", "obs_scale", "reward_gain",
            "reward_pow", "reward_div", "reward_offset", "action_punish", "delta_action_punish", "rtol", "atol")] + \
        [("adaptive", C.c_int32), ("reserved0", C.c_int32)]


class PdeB200Error(RuntimeError):
    pass


_lib = None

_PROTOS = {
    # name: (restype, argtypes)
    "pdeb200_abi_version": (C.c_int32, []),
    "pdeb200_default_config": (C.c_int32, [C.c_int32, C.POINTER(Config)]),
    "pdeb200_create": (C.c_int32, [C.POINTER(Config), C.c_int32, C.POINTER(C.c_void_p)]),
    "pdeb200_destroy": (C.c_int32, [C.c_void_p]),
    "pdeb200_last_error": (C.c_char_p, [C.c_void_p]),
    "pdeb200_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_synchronize": (C.c_int32, [C.c_void_p]),
    "pdeb200_set_bases": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "pdeb200_set_y0": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pdeb200_reset": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_step": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_step_device": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_step_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_act_step_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "pdeb200_noise_prefetch": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_result_layout": (C.c_int32, [C.c_void_p] + [C.POINTER(C.c_size_t)] * 4),
    "pdeb200_result_select": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdeb200_get": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "pdeb200_get_env": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]),
    "pdeb200_reset_diverged": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pdeb200_set": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "pdeb200_device_ptr": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "pdeb200_obs_rows": (C.c_int32, [C.c_void_p]),
    "pdeb200_obs_cols": (C.c_int32, [C.c_void_p]),
    "pdeb200_net_set": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_net_get": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "pdeb200_net_num_params": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdeb200_net_forward": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pdeb200_net_forward_device": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                               C.c_int32, C.POINTER(C.c_int32)]),
    "pdeb200_policy_act": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "pdeb200_policy_act_rng": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double]),
    "pdeb200_rollout": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
    "pdeb200_traj_create": (C.c_int32, [C.c_void_p, C.c_int64]),
    "pdeb200_traj_length": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    "pdeb200_traj_push_pre": (C.c_int32, [C.c_void_p]),
    "pdeb200_traj_push_post": (C.c_int32, [C.c_void_p]),
    "pdeb200_traj_episode_end": (C.c_int32, [C.c_void_p]),
    "pdeb200_traj_pop_tail": (C.c_int32, [C.c_void_p]),
    "pdeb200_sample": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64]),
    "pdeb200_set_batch": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_ddpg_set_path": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdeb200_ddpg_critic_grads": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_int64]),
    "pdeb200_ddpg_critic_apply": (C.c_int32, [C.c_void_p, C.c_double]),
    "pdeb200_ddpg_actor_grads": (C.c_int32, [C.c_void_p, C.c_int64]),
    "pdeb200_ddpg_actor_apply": (C.c_int32, [C.c_void_p, C.c_double, C.c_double]),
    "pdeb200_ddpg_update": (C.c_int32, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "pdeb200_train_updates": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_int32, C.c_uint64]),
    "pdeb200_net_set_params": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "pdeb200_opt_get": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pdeb200_opt_set": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pdeb200_traj_info": (C.c_int32, [C.c_void_p] + [C.POINTER(C.c_int64)] * 5),
    "pdeb200_traj_get": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_traj_set": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_rng_get": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "pdeb200_rng_set": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "pdeb200_get_batch": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdeb200_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "pdeb200_comm_init": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "pdeb200_comm_destroy": (C.c_int32, [C.c_void_p]),
    "pdeb200_comm_info": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pdeb200_comm_allreduce_f64": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pdeb200_launch_count": (C.c_int64, [C.c_void_p]),
    "pdeb200_last_step_ms": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float)]),
    "pdeb200_last_phase_ms": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float)]),
    "pdeb200_last_core_ms": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float)]),
    "pdeb200_last_core_kernel": (C.c_char_p, [C.c_void_p]),
    "pdeb200_enable_step_timing": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdeb200_measure_fma_peak": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "pdeb200_debug_timeline": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pdeb200_step_cost": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def exported_symbols():
    """Names include/pdeb200.h declares (used by the CPU test-suite)."""
    return sorted(_PROTOS)


def load():
    """dlopen libpdeb200.so and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PdeB200Error(
            "%s is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pdeb200_abi_version() != ABI_VERSION:
        raise PdeB200Error("libpdeb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != OK:
        msg = load().pdeb200_last_error(ctx)
        raise PdeB200Error("pdeb200 error %d: %s" % (rc, (msg or b"").decode(errors="replace")))
