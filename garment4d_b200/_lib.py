"""ctypes binding of libgarment4d_b200.so (the C ABI declared in include/garment4d_b200.h).

There is no fallback: if the shared library is missing or a launch fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libgarment4d_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t


class SaMlpDesc(ctypes.Structure):
    """struct g4d_sa_mlp_desc"""
    _fields_ = [("c_in", _i), ("c1", _i), ("c2", _i), ("c3", _i), ("nsample", _i), ("k0", _i)]


class Mlp2Desc(ctypes.Structure):
    """struct g4d_mlp2_desc"""
    _fields_ = [("c_in", _i), ("c1", _i), ("c2", _i)]


class FpDesc(ctypes.Structure):
    """struct g4d_fp_desc"""
    _fields_ = [("c_in", _i), ("c1", _i), ("c2", _i), ("h1", _i), ("h2", _i)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "g4d_last_error": (ctypes.c_char_p, []),
    "g4d_abi_version": (_i, []),
    "g4d_sm_count": (_i, []),
    "g4d_launch_count": (ctypes.c_ulonglong, []),
    "g4d_furthest_point_sampling": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_gather_points": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_gather_points_grad": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_ball_query": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "g4d_group_points": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_group_points_grad": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_three_nn": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_three_interpolate": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_three_interpolate_grad": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fps_gather": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_ball_query2": (_i, [_i, _i, _i, _f, _i, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "g4d_group_fused": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_group_fused_pm": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_grid_bytes": (_sz, [_i, _i]),
    "g4d_grid_build": (_i, [_i, _i, _vp, _f, _vp, _vp]),
    "g4d_fps_gather_grid": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_fps_concurrency_hint": (None, [_i]),
    "g4d_fps_workspace_bytes": (_sz, [_i, _i]),
    "g4d_fps_gather_ws": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_ball_query2_grid": (_i, [_i, _i, _i, _f, _i, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "g4d_ball_query2_grid_ordered": (_i, [_i, _i, _i, _f, _i, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_three_nn_grid": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_sa_mlp_k0": (_i, [_i]),
    "g4d_sa_mlp_param_bytes": (_sz, [ctypes.POINTER(SaMlpDesc)]),
    "g4d_sa_mlp_pack_params": (_i, [ctypes.POINTER(SaMlpDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_sa_mlp_max": (_i, [ctypes.POINTER(SaMlpDesc), _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "g4d_debug_timeline": (None, [_vp]),
    "g4d_debug_fps_phases": (_i, [_vp]),
    "g4d_debug_mlp2_counters": (_i, [_vp]),
    "g4d_debug_fp_counters": (None, [_vp]),
    "g4d_fp_param_bytes": (_sz, [ctypes.POINTER(FpDesc)]),
    "g4d_fp_pack_params": (_i, [ctypes.POINTER(FpDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fp_interp_mlp": (_i, [ctypes.POINTER(FpDesc), _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fp_interp_mlp_labels": (_i, [ctypes.POINTER(FpDesc), _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_mlp2_param_bytes": (_sz, [ctypes.POINTER(Mlp2Desc)]),
    "g4d_mlp2_pack_params": (_i, [ctypes.POINTER(Mlp2Desc), _vp, _vp, _vp, _vp, _vp]),
    "g4d_mlp2_rows": (_i, [ctypes.POINTER(Mlp2Desc), _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_bias_relu_inplace": (_i, [_i, _i, ctypes.c_longlong, _vp, _vp, _i, _vp]),
    "g4d_fp_interp_concat": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fp_interp_concat_cbn_h": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fp_interp_concat_pm_cbn_h": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_fp_interp_concat_rows_h": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_bias_relu_rows_h": (_i, [ctypes.c_longlong, _i, _vp, _vp, _i, _vp]),
    "g4d_bias_relu_rows_unpack": (_i, [_i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "g4d_bias_relu_h": (_i, [_i, ctypes.c_longlong, _vp, _vp, _i, _vp]),
    "g4d_bias_relu_unpack": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "g4d_bias_relu_pm": (_i, [_i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "g4d_scatter_det_workspace_bytes": (_sz, [_i, _i, _i]),
    "g4d_scatter_det_build": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "g4d_scatter_det_apply": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_select_points": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_knn_points": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_knn_inverse_weights": (_i, [ctypes.c_longlong, _i, _i, _vp, _vp, _vp]),
    "g4d_knn_blend_weights": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_smooth_weights": (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_pe_mlp_max": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_batch_rodrigues": (_i, [_i, _vp, _vp, _vp]),
    "g4d_blend_shapes": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_vertices2joints": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "g4d_batch_rigid_transform": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4d_lbs_skin": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4d_lbs_workspace_bytes": (_sz, [_i, _i, _i]),
    "g4d_lbs": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_LIB = None


class G4DError(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise G4DError(
                f"{SO_PATH} is missing: the CUDA extension is not built. Run garment4d_b200/csrc/build.sh "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().g4d_last_error()
        raise G4DError(f"{what} failed (cudaError {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
