"""ctypes binding of the C ABI in ``include/dcpt_ops.h`` (mirrors how the reference binds its
native ops: ``basicsr/ops/layernorm/layernorm.py:7-29`` loads a compiled extension or builds it
when ``BASICSR_JIT=True``)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcpt_sm100.so")
LIB_PATH_FP16 = os.path.join(HERE, "libdcpt_sm100_fp16.so")

_lib = None


class DcptError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_int), ("a_mn", C.c_int),
                ("B", C.c_void_p), ("ldb", C.c_int), ("b_mn", C.c_int),
                ("splits", C.c_int), ("epilogue", C.c_int),
                ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("ldo", C.c_int),
                ("bias", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int),
                ("out2_bf16", C.c_void_p), ("ldo2", C.c_int),
                ("aux_bf16", C.c_void_p), ("ldaux", C.c_int),
                ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cseg", C.c_int),
                ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p), ("ln_out", C.c_void_p), ("ld_ln", C.c_int),
                ("ln_stats", C.c_void_p), ("ln_eps", C.c_float),
                ("lnb_x", C.c_void_p), ("ld_lnb", C.c_int), ("lnb_stats", C.c_void_p), ("lnb_weight", C.c_void_p),
                ("lnb_dres", C.c_void_p), ("lnb_dweight", C.c_void_p), ("lnb_dbias", C.c_void_p), ("lnb_colsum", C.c_void_p),
                ("ln_nocenter", C.c_int)]


_VP, _I, _F, _SZ, _LL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes).  Every symbol include/dcpt_ops.h declares is listed here;
# tests/test_abi_cpu.py checks the two stay in sync.
PROTOTYPES = {
    "dcpt_abi_version": (_I, []),
    "dcpt_operand_dtype": (_I, []),
    "dcpt_last_error": (C.c_char_p, []),
    "dcpt_launch_count": (_LL, []),
    "dcpt_prof_enable": (_I, [_I]),
    "dcpt_prof_dump": (_LL, [C.c_char_p, _LL]),
    "dcpt_layernorm2d_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _F, _VP]),
    "dcpt_layernorm2d_bwd": (_I, [_VP] * 10 + [_I, _I, _VP]),
    "dcpt_gemm_bf16": (_I, [_VP, _I, _I, _VP, _I, _I, _I, _I, _I, _VP, _VP, _I, _VP, _VP, _I, _I, _I, _VP]),
    "dcpt_gemm_ex": (_I, [C.POINTER(GemmDesc), _I, _VP]),
    "dcpt_dwconv3x3_gate_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_nafblock_packed_bytes": (_SZ, [_I]),
    "dcpt_nafblock_saved_bytes": (_SZ, [_I, _I, _I, _I]),
    "dcpt_nafblock_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "dcpt_nafblock_pack": (_I, [_PP, _VP, _I, _VP]),
    "dcpt_nafblock_fwd": (_I, [_PP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_nafblock_bwd": (_I, [_PP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _PP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_nafnet_create": (_VP, [_I, _I, _I, C.POINTER(_I), _I, C.POINTER(_I), _I]),
    "dcpt_nafnet_destroy": (None, [_VP]),
    "dcpt_nafnet_set_tlc": (_I, [_VP, C.POINTER(_I), C.POINTER(_I), _I]),
    "dcpt_nafnet_set_hook_blocks": (_I, [_VP, C.POINTER(_I), _I]),
    "dcpt_nafnet_set_bwd_split_event": (_I, [_VP, _VP, _I]),
    "dcpt_nafnet_num_params": (_I, [_VP]),
    "dcpt_nafnet_param_shape": (_LL, [_VP, _I, C.POINTER(_I)]),
    "dcpt_nafnet_packed_bytes": (_SZ, [_VP]),
    "dcpt_nafnet_saved_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_nafnet_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_nafnet_pack": (_I, [_VP, _PP, _VP, _VP]),
    "dcpt_nafnet_fwd": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _PP, _I, _I, _I, _I, _VP]),
    "dcpt_nafnet_set_keep_activations": (_I, [_VP, _I]),
    "dcpt_nafnet_bwd": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _PP, _PP, _VP, _I, _I, _I, _VP]),
    "dcpt_pack_matrix": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "dcpt_conv3x3_packed_elems": (_SZ, [_I, _I, _I]),
    "dcpt_conv3x3_pack": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "dcpt_conv3x3_fwd": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "dcpt_conv3x3_wgrad": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "dcpt_ln_act_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _F, _VP]),
    "dcpt_ln_act_bwd": (_I, [_VP] * 9 + [_I, _I, _I, _VP]),
    "dcpt_mix_fwd": (_I, [_VP, _VP, _VP, _VP, _LL, _VP]),
    "dcpt_mix_bwd": (_I, [_VP, _VP, _VP, _VP, _VP, _LL, _VP]),
    "dcpt_maxpool2_relu_fwd": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_maxpool2_relu_bwd": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_meanpool_fc_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_meanpool_fc_bwd": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_im2col7x7s2": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "dcpt_add_bf16": (_I, [_VP, _VP, _VP, _LL, _VP]),
    "dcpt_layernorm_rows_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _F, _I, _VP]),
    "dcpt_dwconv3x3_fwd": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "dcpt_dwconv3x3_gelu_gate_fwd": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "dcpt_restormer_create": (_VP, [_I, _I, _I, C.POINTER(_I), _I, C.POINTER(_I), C.c_double, _I, _I]),
    "dcpt_restormer_destroy": (None, [_VP]),
    "dcpt_restormer_set_attention": (_I, [_VP, _I]),
    "dcpt_restormer_num_params": (_I, [_VP]),
    "dcpt_restormer_param_shape": (_LL, [_VP, _I, C.POINTER(_I)]),
    "dcpt_restormer_packed_bytes": (_SZ, [_VP]),
    "dcpt_restormer_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_restormer_pack": (_I, [_VP, _PP, _VP, _VP]),
    "dcpt_restormer_block_fwd": (_I, [_VP, _I, _I, _PP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "dcpt_restormer_block_saved_bytes": (_SZ, [_VP, _I, _I, _I, _I, _I]),
    "dcpt_restormer_block_workspace_bytes": (_SZ, [_VP, _I, _I, _I, _I, _I]),
    "dcpt_restormer_block_fwd_train": (_I, [_VP, _I, _I, _PP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "dcpt_restormer_block_bwd": (_I, [_VP, _I, _I, _PP, _VP, _VP, _VP, _VP, _VP, _PP, _VP, _I, _I, _I, _VP]),
    "dcpt_restormer_saved_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_restormer_bwd_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_restormer_fwd_train": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _VP, _PP, _I, _I, _I, _I, _VP]),
    "dcpt_restormer_bwd": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _PP, _PP, _VP, _I, _I, _I, _VP]),
    "dcpt_optim_create": (_VP, [C.POINTER(_LL), _I]),
    "dcpt_optim_destroy": (None, [_VP]),
    "dcpt_optim_workspace_bytes": (_SZ, [_VP]),
    "dcpt_optim_num_chunks": (_LL, [_VP]),
    "dcpt_optim_bind": (_I, [_VP, _VP, _PP, _PP, _PP, _PP, _PP, _VP]),
    "dcpt_optim_grad_norm": (_I, [_VP, _VP, _VP, _VP]),
    "dcpt_optim_set_norm": (_I, [_VP, _VP, _VP, _VP]),
    "dcpt_optim_param_hash": (_I, [_VP, _VP, _VP, _VP]),
    "dcpt_optim_step": (_I, [_VP, _VP, _I, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _LL, C.c_double, C.c_double, _VP]),
    "dcpt_restormer_fwd": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _PP, _I, _I, _I, _I, _VP]),
    "dcpt_promptir_create": (_VP, [_I, _I, _I, C.POINTER(_I), _I, C.POINTER(_I), C.c_double, _I, _I]),
    "dcpt_promptir_packed_bytes": (_SZ, [_VP]),
    "dcpt_promptir_workspace_bytes": (_SZ, [_VP, _I, _I, _I]),
    "dcpt_promptir_pack": (_I, [_VP, _PP, _VP, _VP]),
    "dcpt_promptir_fwd": (_I, [_VP, _PP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
}


def load_library(path=None):
    """Load ``libdcpt_sm100.so`` (built in-tree by ``python -m dcpt_b200.build`` /
    ``__graft_entry__.build()``; ``BASICSR_JIT=True`` builds it on first use like the reference's
    ops do).  Raises DcptError if it is missing — there is no fallback path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # DCPT_LIB: debug builds (libdcpt_sm100_trace.so); DCPT_OPERAND=fp16: the parity build with IEEE-half operands
    path = path or os.getenv("DCPT_LIB") or (LIB_PATH_FP16 if os.getenv("DCPT_OPERAND", "bf16").lower() == "fp16" else LIB_PATH)
    if not os.path.exists(path) and os.getenv("BASICSR_JIT") == "True":
        from .build import build
        build()
    if not os.path.exists(path):
        raise DcptError(f"{path} not found: build it with `python -m dcpt_b200.build` (needs nvcc) "
                        "or set BASICSR_JIT=True. dcpt_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale -> loud
        fn.restype = res
        fn.argtypes = args
    if lib.dcpt_abi_version() != 1:
        raise DcptError(f"ABI version mismatch: library {lib.dcpt_abi_version()} != binding 1")
    _lib = lib
    return lib


def operand_dtype():
    """torch dtype of the 16-bit operand tensors of the loaded library (bf16 by default, fp16 with DCPT_OPERAND=fp16)."""
    import torch
    return torch.float16 if load_library().dcpt_operand_dtype() == 1 else torch.bfloat16


def check(rc, what=""):
    if rc != 0:
        msg = _lib.dcpt_last_error().decode(errors="replace") if _lib is not None else ""
        raise DcptError(f"{what} failed (rc={rc}): {msg}")


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) if p else None for p in ptrs])
    return arr
