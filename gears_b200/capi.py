"""ctypes binding of include/gears_physics_cuda.h (the same symbols the Rust -sys crate binds).

There is no CPU path: if the shared library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgears_physics_cuda.so")

GP_OK = 0
GP_WARN_MARGIN_RETRIED = 1
GP_ERR_BAD_ARG, GP_ERR_CUDA, GP_ERR_NO_DEVICE, GP_ERR_CAPACITY = -1, -2, -3, -4
GP_ERR_PAIR_OVERFLOW, GP_ERR_MARGIN, GP_ERR_NCCL, GP_ERR_STATE = -5, -6, -7, -8
GP_ERR_PEER_TIMEOUT = -9
GP_FLAG_RECORD_PAIRS = 0x1
GP_FLAG_NO_RETRY = 0x2
GP_FLAG_PROFILE = 0x4
GP_FLAG_BIG_STATIC_ONLY = 0x8
GP_NUM_KERNEL_GROUPS = 8
KERNEL_GROUPS = ["keys", "exchange", "cell_scan", "scatter", "find_pairs", "chains", "sweep", "finish"]
GP_UNIQUE_ID_BYTES = 128

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class GpConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_bodies", C.c_uint32),
                ("max_pairs", C.c_uint64), ("margin_init", C.c_float), ("margin_max", C.c_float),
                ("cell_size", C.c_float), ("max_big", C.c_uint32), ("flags", C.c_uint32),
                ("grid_min", C.c_float * 3), ("grid_max", C.c_float * 3)]


class GpStepStats(C.Structure):
    _fields_ = [("n_bodies", C.c_uint32), ("n_big", C.c_uint32), ("n_candidates", C.c_uint64),
                ("n_snapshot", C.c_uint64), ("n_contacts", C.c_uint64), ("n_rounds", C.c_uint32),
                ("n_retries", C.c_uint32), ("margin", C.c_float), ("max_displacement", C.c_float),
                ("cell_size", C.c_float), ("grid_dim", C.c_uint32 * 3), ("n_heavy", C.c_uint32),
                ("n_migrated_out", C.c_uint32), ("n_migrated_in", C.c_uint32), ("n_halo_out", C.c_uint32),
                ("n_halo_in", C.c_uint32), ("steps_total", C.c_uint64), ("steps_inexact", C.c_uint64),
                ("n_unverified", C.c_uint32), ("unverified_total", C.c_uint64), ("n_watch", C.c_uint32),
                ("n_activated", C.c_uint32)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class GpSlabConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_int32), ("nranks", C.c_int32), ("x_lo", C.c_float),
                ("x_hi", C.c_float), ("halo", C.c_float), ("max_exchange", C.c_uint32), ("flags", C.c_uint32)]


GP_SLAB_NO_VERIFY = 0x1


class GpKernelTimes(C.Structure):
    _fields_ = [("steps", C.c_uint32), ("ms", C.c_float * 8), ("launches", C.c_uint32 * 8)]


# every symbol declared in include/gears_physics_cuda.h
SYMBOLS = [
    "gp_abi_version", "gp_create", "gp_destroy", "gp_last_error", "gp_upload", "gp_set_state", "gp_update_dirty",
    "gp_update_shape", "gp_step", "gp_step_async", "gp_sync", "gp_download", "gp_pairs", "gp_damping_factors",
    "gp_host_alloc", "gp_host_free", "gp_last_step_ms", "gp_kernel_launches", "gp_kernel_times_read", "gp_comm_unique_id", "gp_comm_init",
    "gp_comm_init_local", "gp_step_group", "gp_owned_count", "gp_download_owned", "gp_set_state_owned", "gp_slab_rebalance",
    "gp_slab_rebalance_local", "gp_instances", "gp_raycast",
    "gp_obstacle_grid", "gp_cap_velocity", "gp_frame_submit", "gp_frame_wait", "gp_test_radix_sort", "gp_test_radix_sort64",
    "gp_test_exclusive_scan",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(SO, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    L.gp_abi_version.restype = C.c_int32
    L.gp_create.argtypes = [C.POINTER(GpConfig), C.POINTER(vp)]
    L.gp_create.restype = C.c_int32
    L.gp_destroy.argtypes = [vp]
    L.gp_destroy.restype = None
    L.gp_last_error.argtypes = [vp]
    L.gp_last_error.restype = C.c_char_p
    L.gp_upload.argtypes = [vp, C.c_uint32] + [C.c_void_p] * 12
    L.gp_upload.restype = C.c_int32
    L.gp_set_state.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_set_state.restype = C.c_int32
    L.gp_update_dirty.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_update_dirty.restype = C.c_int32
    L.gp_update_shape.argtypes = [vp, C.c_uint32] + [C.c_void_p] * 8
    L.gp_update_shape.restype = C.c_int32
    L.gp_step.argtypes = [vp, C.c_float, C.c_void_p, C.POINTER(GpStepStats)]
    L.gp_step.restype = C.c_int32
    L.gp_step_async.argtypes = [vp, C.c_float, C.c_void_p, C.c_uint32]
    L.gp_step_async.restype = C.c_int32
    L.gp_sync.argtypes = [vp, C.POINTER(GpStepStats)]
    L.gp_sync.restype = C.c_int32
    L.gp_download.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.gp_download.restype = C.c_int32
    L.gp_pairs.argtypes = [vp, C.c_int32, C.c_void_p, C.c_uint64, u64p]
    L.gp_pairs.restype = C.c_int32
    L.gp_damping_factors.argtypes = [C.c_float, f32p]
    L.gp_damping_factors.restype = None
    L.gp_host_alloc.argtypes = [C.c_size_t]
    L.gp_host_alloc.restype = C.c_void_p
    L.gp_host_free.argtypes = [C.c_void_p]
    L.gp_host_free.restype = None
    L.gp_last_step_ms.argtypes = [vp, f32p]
    L.gp_last_step_ms.restype = C.c_int32
    L.gp_kernel_launches.argtypes = [vp]
    L.gp_kernel_launches.restype = C.c_uint64
    L.gp_kernel_times_read.argtypes = [vp, C.POINTER(GpKernelTimes)]
    L.gp_kernel_times_read.restype = C.c_int32
    L.gp_comm_unique_id.argtypes = [C.c_void_p]
    L.gp_comm_unique_id.restype = C.c_int32
    L.gp_comm_init.argtypes = [vp, C.c_void_p, C.POINTER(GpSlabConfig)]
    L.gp_comm_init.restype = C.c_int32
    L.gp_comm_init_local.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(GpSlabConfig)]
    L.gp_comm_init_local.restype = C.c_int32
    L.gp_step_group.argtypes = [C.POINTER(vp), C.c_int32, C.c_float, C.c_void_p, C.c_uint32]
    L.gp_step_group.restype = C.c_int32
    L.gp_owned_count.argtypes = [vp, u32p]
    L.gp_owned_count.restype = C.c_int32
    L.gp_download_owned.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_download_owned.restype = C.c_int32
    L.gp_set_state_owned.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_set_state_owned.restype = C.c_int32
    L.gp_slab_rebalance.argtypes = [vp, C.c_float, u32p]
    L.gp_slab_rebalance.restype = C.c_int32
    L.gp_slab_rebalance_local.argtypes = [C.POINTER(vp), C.c_int32, C.c_float, u32p]
    L.gp_slab_rebalance_local.restype = C.c_int32
    L.gp_instances.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_instances.restype = C.c_int32
    L.gp_raycast.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.gp_raycast.restype = C.c_int32
    L.gp_obstacle_grid.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gp_obstacle_grid.restype = C.c_int32
    L.gp_cap_velocity.argtypes = [vp, C.c_uint32, C.c_void_p]
    L.gp_cap_velocity.restype = C.c_int32
    L.gp_frame_submit.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    L.gp_frame_submit.restype = C.c_int32
    L.gp_frame_wait.argtypes = [vp, C.c_uint32, C.POINTER(GpStepStats)]
    L.gp_frame_wait.restype = C.c_int32
    L.gp_test_radix_sort.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    L.gp_test_radix_sort.restype = C.c_int32
    L.gp_test_radix_sort64.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    L.gp_test_radix_sort64.restype = C.c_int32
    L.gp_test_exclusive_scan.argtypes = [C.c_int32, C.c_uint32, C.c_void_p, u32p]
    L.gp_test_exclusive_scan.restype = C.c_int32
    _lib = L
    return L
