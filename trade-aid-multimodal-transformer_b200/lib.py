"""ctypes binding of libtradeaid_b200.so (C ABI: include/tradeaid_b200.h).

Fails loudly: if the shared library is absent or does not export the expected ABI version, import raises — there is
no CPU or PyTorch fallback for any op declared here."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtradeaid_b200.so")
ABI_VERSION = 1


class TaOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("major", C.c_int32), ("_pad", C.c_int32), ("ld", C.c_int64),
                ("gs0", C.c_int64), ("gs1", C.c_int64)]


class TaEpilogue(C.Structure):
    _fields_ = [
        ("alpha", C.c_float), ("act", C.c_int32),
        ("bias", C.c_void_p), ("bias_gs0", C.c_int64), ("bias_gs1", C.c_int64),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("aux_gs0", C.c_int64), ("aux_gs1", C.c_int64),
        ("aux_mode", C.c_int32), ("_pad", C.c_int32),
        ("residual", C.c_void_p), ("ld_res", C.c_int64), ("res_gs0", C.c_int64), ("res_gs1", C.c_int64),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int64), ("f32_gs0", C.c_int64), ("f32_gs1", C.c_int64),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int64), ("bf16_gs0", C.c_int64), ("bf16_gs1", C.c_int64),
        ("dropout_p", C.c_float), ("_pad2", C.c_int32), ("dropout_seed", C.c_uint64),
    ]


ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2
AUX_NONE, AUX_TANH_BWD, AUX_RELU_BWD = 0, 1, 2


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). tradeaid_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.ta_last_error.restype = C.c_char_p
    lib.ta_abi_version.restype = C.c_int32
    lib.ta_launch_count.restype = C.c_int64
    lib.ta_reset_launch_count.restype = None
    if lib.ta_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.ta_abi_version()} != expected {ABI_VERSION}; rebuild")
    return lib


_lib = _load()


class TaError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise TaError(_lib.ta_last_error().decode())


def launch_count():
    return int(_lib.ta_launch_count())


def reset_launch_count():
    _lib.ta_reset_launch_count()


def raw():
    """The ctypes CDLL handle (symbol table checks in tests)."""
    return _lib


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ---------------------------------------------------------------------------------------------------- GEMM
_lib.ta_gemm_bf16.restype = C.c_int32
_lib.ta_gemm_bf16.argtypes = [C.POINTER(TaOperand), C.POINTER(TaOperand), C.c_int64, C.c_int64, C.c_int64,
                              C.c_int32, C.c_int32, C.POINTER(TaEpilogue), C.c_void_p]


def _grp(t, nd_tail):
    """(gs0, gs1, G0, G1) of a tensor whose last nd_tail dims are the matrix/vector and up to two leading dims are
    groups (innermost group first = g0)."""
    lead = t.dim() - nd_tail
    if lead == 0:
        return 0, 0, 1, 1
    if lead == 1:
        return t.stride(0), 0, t.shape[0], 1
    if lead == 2:
        return t.stride(1), t.stride(0), t.shape[1], t.shape[0]
    raise ValueError("at most two group dims")


def _operand(t, major):
    import torch
    assert t.dtype == torch.bfloat16 and t.stride(-1) == 1, "gemm operands are bf16 with unit inner stride"
    gs0, gs1, _, _ = _grp(t, 2)
    return TaOperand(t.data_ptr(), major, 0, t.stride(-2), gs0, gs1)


def gemm(A, B, *, a_major=0, b_major=0, out_f32=None, out_bf16=None, bias=None, act=ACT_NONE, aux=None,
         aux_mode=AUX_NONE, residual=None, alpha=1.0):
    """D[g] = epi(alpha * A[g] @ B[g]^T). Tensors are [..groups.., rows, cols] views (up to two group dims, inner
    stride 1). For a_major/b_major = 1 the stored matrix is [K, M] / [K, N]."""
    K = A.shape[-1] if a_major == 0 else A.shape[-2]
    M = A.shape[-2] if a_major == 0 else A.shape[-1]
    N = B.shape[-2] if b_major == 0 else B.shape[-1]
    Kb = B.shape[-1] if b_major == 0 else B.shape[-2]
    assert K == Kb, f"reduction dims differ: {K} vs {Kb}"
    _, _, G0, G1 = _grp(A, 2)
    assert _grp(B, 2)[2:] == (G0, G1), "A and B group dims differ"
    ep = TaEpilogue()
    ep.alpha = alpha
    ep.act = act
    if bias is not None:
        assert bias.dtype.is_floating_point and bias.element_size() == 4 and bias.shape[-1] == N
        s0, s1, g0, g1 = _grp(bias, 1)
        assert (g0, g1) == (G0, G1)
        ep.bias, ep.bias_gs0, ep.bias_gs1 = bias.data_ptr(), s0, s1
    if aux is not None:
        assert aux.shape[-2:] == (M, N) and aux.stride(-1) == 1 and aux.element_size() == 2
        s0, s1, g0, g1 = _grp(aux, 2)
        assert (g0, g1) == (G0, G1)
        ep.aux, ep.ld_aux, ep.aux_gs0, ep.aux_gs1, ep.aux_mode = aux.data_ptr(), aux.stride(-2), s0, s1, aux_mode
    if residual is not None:
        assert residual.shape[-2:] == (M, N) and residual.stride(-1) == 1 and residual.element_size() == 4
        s0, s1, g0, g1 = _grp(residual, 2)
        assert (g0, g1) == (G0, G1)
        ep.residual, ep.ld_res, ep.res_gs0, ep.res_gs1 = residual.data_ptr(), residual.stride(-2), s0, s1
    if out_f32 is not None:
        assert out_f32.shape[-2:] == (M, N) and out_f32.stride(-1) == 1 and out_f32.element_size() == 4
        s0, s1, g0, g1 = _grp(out_f32, 2)
        assert (g0, g1) == (G0, G1)
        ep.out_f32, ep.ld_f32, ep.f32_gs0, ep.f32_gs1 = out_f32.data_ptr(), out_f32.stride(-2), s0, s1
    if out_bf16 is not None:
        assert out_bf16.shape[-2:] == (M, N) and out_bf16.stride(-1) == 1 and out_bf16.element_size() == 2
        s0, s1, g0, g1 = _grp(out_bf16, 2)
        assert (g0, g1) == (G0, G1)
        ep.out_bf16, ep.ld_bf16, ep.bf16_gs0, ep.bf16_gs1 = out_bf16.data_ptr(), out_bf16.stride(-2), s0, s1
    a, b = _operand(A, a_major), _operand(B, b_major)
    _check(_lib.ta_gemm_bf16(C.byref(a), C.byref(b), M, N, K, G0, G1, C.byref(ep), _stream()))
