"""ctypes binding of include/gadcu.h. No torch types cross this boundary: plain pointers and sizes.

The library is loaded from gad_b200/lib/libgadcu.so (built in-tree by gad_b200.build). There is no
CPU fallback: if the library is missing this module raises, and every compute entry point returns
GADCU_ERR_CUDA on a host without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgadcu.so")


class Dim4(C.Structure):
    _fields_ = [("d", C.c_uint64 * 4)]

    @staticmethod
    def of(dims) -> "Dim4":
        d = list(int(x) for x in dims)
        if len(d) > 4:
            raise ValueError("at most 4 dimensions")
        d += [1] * (4 - len(d))
        return Dim4((C.c_uint64 * 4)(*d))

    def tuple(self):
        return tuple(int(x) for x in self.d)


class MatProp(C.Structure):
    _fields_ = [("transposed", C.c_uint8), ("conjugated", C.c_uint8)]


class ValueC(C.Structure):
    _fields_ = [("data", C.c_void_p), ("arena", C.c_uint32), ("index", C.c_uint32)]


VJP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ValueC))
DROP_FN = C.CFUNCTYPE(None, C.c_void_p)

P = C.c_void_p
PP = C.POINTER(C.c_void_p)
VP = C.POINTER(ValueC)
DP = C.POINTER(Dim4)

_UNARY = ["zeros", "ones", "neg", "exp", "log", "log1p", "sin", "cos", "tanh", "sigmoid", "reciprocal", "sqrt",
          "abs", "sign", "relu", "flat", "as_scalar", "norm2"]
_BINARY = ["add", "sub", "mul", "div", "pow", "min", "max", "scale", "dot", "matmul_nn"]
_CONST = ["setc", "addc", "mulc", "powc"]
_DIMS = ["moddims", "tile_as", "sum_as", "constant_as", "max_as", "argmax_as", "softmax_as"]

# name -> (restype, argtypes). Every symbol declared in include/gadcu.h appears here; the CPU test
# suite checks the two lists against each other.
SIGNATURES = {
    "gadcu_last_error": (C.c_char_p, []),
    "gadcu_version": (C.c_char_p, []),
    "gadcu_debug_gemm_schedule": (C.c_size_t, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gadcu_ctx_create": (C.c_int, [C.c_int, PP]),
    "gadcu_ctx_destroy": (C.c_int, [P]),
    "gadcu_ctx_synchronize": (C.c_int, [P]),
    "gadcu_ctx_stream": (P, [P]),
    "gadcu_ctx_set_fusion": (C.c_int, [P, C.c_int]),
    "gadcu_ctx_set_strict_reference": (C.c_int, [P, C.c_int]),
    "gadcu_ctx_set_gemm_epilogue": (C.c_int, [P, C.c_int]),
    "gadcu_ctx_set_auto_replay": (C.c_int, [P, C.c_int]),
    "gadcu_ctx_flush": (C.c_int, [P]),
    "gadcu_ctx_replay_stats": (C.c_int, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gadcu_ctx_memory_stats": (C.c_int, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gadcu_ctx_profile_begin": (C.c_int, [P]),
    "gadcu_ctx_profile_end": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gadcu_array_from_host": (C.c_int, [P, C.c_int, DP, P, PP]),
    "gadcu_array_to_host": (C.c_int, [P, P]),
    "gadcu_array_alloc": (C.c_int, [P, C.c_int, DP, PP]),
    "gadcu_ctx_set_lazy": (C.c_int, [P, C.c_int, C.c_uint64]),
    "gadcu_comm_allreduce_sum_async": (C.c_int, [P, PP, C.c_size_t]),
    "gadcu_comm_join": (C.c_int, [P]),
    "gadcu_evaluate_gradients_dp": (C.c_int, [P, C.c_uint32, C.c_uint32, P, C.c_int, P, PP]),
    "gadcu_evaluate_gradients_dp_with": (C.c_int, [P, C.c_uint32, C.c_uint32, P, C.c_int, P, PP, C.c_size_t, PP]),
    "gadcu_batched_rerun": (C.c_int, [P, PP, C.c_size_t, PP]),
    "gadcu_array_upload": (C.c_int, [P, P, P]),
    "gadcu_array_upload_async": (C.c_int, [P, P]),
    "gadcu_ctx_wait_uploads": (C.c_int, [P]),
    "gadcu_array_download_async": (C.c_int, [P, P, C.POINTER(C.c_uint64)]),
    "gadcu_ctx_wait_download": (C.c_int, [P, C.c_uint64]),
    "gadcu_array_clone": (C.c_int, [P, PP]),
    "gadcu_array_retain": (C.c_int, [P]),
    "gadcu_array_release": (C.c_int, [P]),
    "gadcu_array_dims": (C.c_int, [P, DP]),
    "gadcu_array_dtype": (C.c_int, [P]),
    "gadcu_array_is_scalar": (C.c_int, [P]),
    "gadcu_array_device_ptr": (P, [P]),
    "gadcu_scalar_from_host": (C.c_int, [P, C.c_int, C.c_double, PP]),
    "gadcu_scalar_to_host": (C.c_int, [P, C.POINTER(C.c_double)]),
    "gadcu_value_release": (C.c_int, [VP]),
    "gadcu_array_random": (C.c_int, [P, C.c_int, DP, C.c_uint64, C.c_int, C.c_double, C.c_double, PP]),
    "gadcu_array_add_assign": (C.c_int, [P, P]),
    "gadcu_array_scale": (C.c_int, [P, C.c_double, PP]),
    "gadcu_check_equal": (C.c_int, [C.POINTER(DP), C.c_size_t, DP]),
    "gadcu_check_select_argmax": (C.c_int, [DP, DP, DP, DP, DP]),
    "gadcu_check_flat": (C.c_int, [DP, DP]),
    "gadcu_check_moddims": (C.c_int, [DP, DP, DP]),
    "gadcu_check_tile_as": (C.c_int, [DP, DP, DP]),
    "gadcu_check_sum_as": (C.c_int, [DP, DP, DP]),
    "gadcu_check_as_scalar": (C.c_int, [DP]),
    "gadcu_check_dot": (C.c_int, [DP, DP]),
    "gadcu_check_reduced": (C.c_int, [DP, DP, C.c_int, DP]),
    "gadcu_check_transpose": (C.c_int, [DP, DP]),
    "gadcu_check_matmul": (C.c_int, [DP, DP, MatProp, MatProp, DP]),
    "gadcu_algebra_create": (C.c_int, [P, C.c_int, PP]),
    "gadcu_algebra_clone": (C.c_int, [P, PP]),
    "gadcu_algebra_destroy": (C.c_int, [P]),
    "gadcu_algebra_get_kind": (C.c_int, [P]),
    "gadcu_algebra_num_nodes": (C.c_size_t, [P]),
    "gadcu_variable": (C.c_int, [P, P, VP]),
    "gadcu_constant": (C.c_int, [P, P, VP]),
    "gadcu_add_all": (C.c_int, [P, C.POINTER(VP), C.c_size_t, VP]),
    "gadcu_select_argmax": (C.c_int, [P, VP, VP, VP, VP, VP]),
    "gadcu_matmul": (C.c_int, [P, VP, VP, MatProp, MatProp, VP]),
    "gadcu_transpose": (C.c_int, [P, VP, C.c_int, VP]),
    "gadcu_make_node": (C.c_int, [P, P, C.POINTER(VP), C.c_size_t, VJP_FN, P, P, VP]),
    "gadcu_store_add_gradient": (C.c_int, [P, P, C.c_uint32, C.c_uint32, VP]),
    "gadcu_evaluate_gradients": (C.c_int, [P, C.c_uint32, C.c_uint32, P, C.c_int, PP]),
    "gadcu_compute_gradients": (C.c_int, [P, C.c_uint32, C.c_uint32, VP, PP]),
    "gadcu_compute_gradients_dp": (C.c_int, [P, C.c_uint32, C.c_uint32, VP, P, PP]),
    "gadcu_store_get": (C.c_int, [P, C.c_uint32, C.c_uint32, VP, C.POINTER(C.c_int)]),
    "gadcu_store_ids": (C.c_size_t, [P, C.POINTER(C.c_uint32), C.c_size_t]),
    "gadcu_store_destroy": (C.c_int, [P]),
    "gadcu_capture_begin": (C.c_int, [P]),
    "gadcu_capture_end": (C.c_int, [P, PP]),
    "gadcu_capture_launch": (C.c_int, [P]),
    "gadcu_capture_destroy": (C.c_int, [P]),
    "gadcu_batched_create": (C.c_int, [P, C.c_int, C.c_uint64, PP]),
    "gadcu_batched_force": (C.c_int, [P, VP, PP]),
    "gadcu_batched_program_stats": (C.c_int, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gadcu_batched_sweep_memo_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gadcu_comm_unique_id": (C.c_int, [P]),
    "gadcu_comm_init": (C.c_int, [P, C.c_int, C.c_int, P, PP]),
    "gadcu_comm_allreduce_sum": (C.c_int, [P, PP, C.c_size_t]),
    "gadcu_comm_destroy": (C.c_int, [P]),
}
for _n in _UNARY:
    SIGNATURES["gadcu_" + _n] = (C.c_int, [P, VP, VP])
for _n in _BINARY:
    SIGNATURES["gadcu_" + _n] = (C.c_int, [P, VP, VP, VP])
for _n in _CONST:
    SIGNATURES["gadcu_" + _n] = (C.c_int, [P, VP, C.c_double, C.c_int, VP])
for _n in _DIMS:
    SIGNATURES["gadcu_" + _n] = (C.c_int, [P, VP, DP, VP])

_lib = None


def load() -> C.CDLL:
    """Load libgadcu.so (once). Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m gad_b200.build` (nvcc, sm_100a). "
            "gad_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
