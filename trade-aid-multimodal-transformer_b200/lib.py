"""ctypes binding of libtradeaid_b200.so (C ABI: include/tradeaid_b200.h).

Fails loudly: if the shared library is absent or does not export the expected ABI version, import raises — there is
no CPU or PyTorch fallback for any op declared here."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRADEAID_B200_LIB") or os.path.join(_HERE, "libtradeaid_b200.so")   # env override: dev A/B builds
ABI_VERSION = 9


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


_lib.ta_set_sm_margin.restype = C.c_int32
_lib.ta_set_sm_margin.argtypes = [C.c_int32]
_lib.ta_get_sm_margin.restype = C.c_int32


def set_sm_margin(sms):
    """SMs the persistent kernels of the following launches leave free (for a concurrent NCCL kernel); 0 = none."""
    _check(_lib.ta_set_sm_margin(int(sms)))


_lib.ta_set_pdl.restype = C.c_int32
_lib.ta_set_pdl.argtypes = [C.c_int32]


def set_pdl(on):
    """Programmatic dependent launch for the following launches of this thread: True / False, or None = the TA_PDL variable."""
    _check(_lib.ta_set_pdl(-1 if on is None else int(bool(on))))


def get_sm_margin():
    return int(_lib.ta_get_sm_margin())


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
         aux_mode=AUX_NONE, residual=None, alpha=1.0, dropout_p=0.0, dropout_seed=0, dropout_step=None):
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
    ep.dropout_p = float(dropout_p)
    ep.dropout_seed = int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
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
    with _step_counter(dropout_step if dropout_p > 0 else None):
        _check(_lib.ta_gemm_bf16(C.byref(a), C.byref(b), M, N, K, G0, G1, C.byref(ep), _stream()))


# ---------------------------------------------------------------------------------------------------- attention
class TaAttnTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("head_stride", C.c_int64), ("gs", C.c_int64)]


_lib.ta_attention_fwd.restype = C.c_int32
_lib.ta_attention_fwd.argtypes = [C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.c_int32,
                                  C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_void_p]


_lib.ta_attention_decode.restype = C.c_int32
_lib.ta_attention_decode.argtypes = [C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor),
                                     C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]


def _attn_tensor(t, head_stride):
    """t: bf16 view [G, B*T, >=H*hs] (or [B*T, ...] for G = 1) whose column 0 is head 0 / dim 0."""
    assert t.element_size() == 2 and t.stride(-1) == 1
    gs = t.stride(0) if t.dim() == 3 else 0
    return TaAttnTensor(t.data_ptr(), t.stride(-2), head_stride, gs)


def attention_fwd(q, ks, vs, o, lse, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None, o_src=None, dropout_p=0.0,
                  dropout_seed=0, dropout_step=None):
    """q: [G, B*T, *]; ks/vs: lists (one per source) of [G, B*T, *] views; o: [G, B*T, H*hs] bf16;
    lse: fp32 [n_src, G, B, H, T]; o_src: optional [n_src, G, B*T, H*hs] bf16."""
    n_src = len(ks)
    G = q.shape[0] if q.dim() == 3 else 1
    hq = head_stride_q if head_stride_q is not None else hs
    hkv = head_stride_kv if head_stride_kv is not None else hs
    qa = (TaAttnTensor * 1)(_attn_tensor(q, hq))
    ka = (TaAttnTensor * n_src)(*[_attn_tensor(k, hkv) for k in ks])
    va = (TaAttnTensor * n_src)(*[_attn_tensor(v, hkv) for v in vs])
    assert o.stride(-1) == 1 and lse.is_contiguous() and lse.numel() == n_src * G * B * H * T
    o_gs = o.stride(0) if o.dim() == 3 else 0
    if o_src is not None:
        assert o_src.stride(-2) == o.stride(-2) and (o.dim() < 3 or o_src.stride(1) == o.stride(0))
    with _step_counter(dropout_step if dropout_p > 0 else None):
        _check(_lib.ta_attention_fwd(qa, ka, va, n_src, _ptr(o), o.stride(-2), o_gs, _ptr(o_src),
                                     o_src.stride(0) if o_src is not None else 0, _ptr(lse), G, B, T, H, hs, float(dropout_p),
                                     int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _stream()))


_lib.ta_attention_fwd_f32.restype = C.c_int32
_lib.ta_attention_fwd_f32.argtypes = [C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.c_int32,
                                      C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]


def attention_fwd_f32(q, ks, vs, o, lse, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None, o_src=None):
    """attention_fwd on fp32 tensors (plain fp32 arithmetic, the bf16x3 precision mode); same views and strides."""
    import torch
    n_src = len(ks)
    G = q.shape[0] if q.dim() == 3 else 1
    hq = head_stride_q if head_stride_q is not None else hs
    hkv = head_stride_kv if head_stride_kv is not None else hs
    for t in [q, o, *ks, *vs] + ([o_src] if o_src is not None else []):
        assert t.dtype == torch.float32 and t.stride(-1) == 1

    def at(t, hstride):
        return TaAttnTensor(t.data_ptr(), t.stride(-2), hstride, t.stride(0) if t.dim() == 3 else 0)

    qa = (TaAttnTensor * 1)(at(q, hq))
    ka = (TaAttnTensor * n_src)(*[at(k, hkv) for k in ks])
    va = (TaAttnTensor * n_src)(*[at(v, hkv) for v in vs])
    assert lse.is_contiguous() and lse.numel() == n_src * G * B * H * T
    o_gs = o.stride(0) if o.dim() == 3 else 0
    if o_src is not None:
        assert o_src.stride(-2) == o.stride(-2) and (o.dim() < 3 or o_src.stride(1) == o.stride(0))
    _check(_lib.ta_attention_fwd_f32(qa, ka, va, n_src, _ptr(o), o.stride(-2), o_gs, _ptr(o_src),
                                     o_src.stride(0) if o_src is not None else 0, _ptr(lse), G, B, T, H, hs, _stream()))


def attention_decode(q, ks, vs, o, *, length, H, hs, head_stride_q=None, head_stride_kv=None, len_dev=None):
    """q: bf16 [G, B, *] (one new row per batch); ks/vs: lists (one per source) of cache views [G, B, capacity, *] whose
    column 0 is head 0 / dim 0; o: bf16 [G, B, H*hs]. Attends to cached positions [0, length) - with len_dev (int32 [1] on the
    device: CUDA-graph replay) to [0, *len_dev + length)."""
    n_src = len(ks)
    G, B = q.shape[0], q.shape[1]
    hq = head_stride_q if head_stride_q is not None else hs
    hkv = head_stride_kv if head_stride_kv is not None else hs
    assert q.dim() == 3 and o.dim() == 3 and o.stride(-1) == 1 and q.stride(-1) == 1
    qa = (TaAttnTensor * 1)(TaAttnTensor(q.data_ptr(), q.stride(1), hq, q.stride(0)))
    for t in list(ks) + list(vs):
        assert t.dim() == 4 and t.shape[0] == G and t.shape[1] == B and t.shape[2] >= length and t.stride(-1) == 1
    if len_dev is not None:
        import torch
        assert len_dev.dtype == torch.int32 and len_dev.numel() == 1
    ka = (TaAttnTensor * n_src)(*[TaAttnTensor(k.data_ptr(), k.stride(2), hkv, k.stride(0)) for k in ks])
    va = (TaAttnTensor * n_src)(*[TaAttnTensor(v.data_ptr(), v.stride(2), hkv, v.stride(0)) for v in vs])
    bs = (C.c_int64 * n_src)(*[k.stride(1) for k in ks])
    _check(_lib.ta_attention_decode(qa, ka, va, bs, n_src, _ptr(o), o.stride(1), o.stride(0), G, B, int(length), _ptr(len_dev),
                                    H, hs, _stream()))


# ---------------------------------------------------------------------------------------------------- row ops
def _sig(name, *argtypes, restype=C.c_int32):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_P, _I32, _I64, _F32, _F64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_uint64
_gather_batch = _sig("ta_gather_batch", _P, _I64, _I32, _P, _I32, _I32, _P, _P, _P)
_embed_fwd = _sig("ta_embed_fwd", _P, _P, _P, _P, _P, _I32, _I64, _I32, _I32, _P)
_embed_fwd_at = _sig("ta_embed_fwd_at", _P, _P, _P, _P, _P, _I32, _I64, _I32, _P, _P)
_kv_append = _sig("ta_kv_append", _P, _I64, _I64, _P, _I64, _I64, _I64, _I32, _I32, _I32, _P, _P)
_embed_bwd = _sig("ta_embed_bwd", _P, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _I32, _P)
_ln_fwd = _sig("ta_layernorm_fwd", _P, _P, _P, _P, _I64, _P, _P, _I64, _I64, _I32, _F32, _P)
_ln_fwd_f32 = _sig("ta_layernorm_fwd_f32", _P, _P, _P, _P, _I64, _P, _P, _I64, _I64, _I32, _F32, _P)
_split = _sig("ta_split_bf16", _P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P)
_convert = _sig("ta_convert", _P, _I32, _I64, _I64, _I64, _P, _I32, _I64, _I64, _I64, _I32, _I32, _I32, _I32, _P)
_ln_bwd_ws = _sig("ta_layernorm_bwd_workspace_floats", _I64, _I64, _I32, restype=_I64)
_ln_bwd = _sig("ta_layernorm_bwd", _P, _I64, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _P)
_ce = _sig("ta_cross_entropy", _P, _I64, _P, _P, _P, _P, _I64, _I64, _I32, _F32, _P)
_colsum_ws = _sig("ta_colsum_workspace_floats", _I64, _I32, _I32, restype=_I64)
_colsum = _sig("ta_colsum_bf16", _P, _I64, _I64, _P, _I64, _P, _I64, _I32, _I32, _P)
_cast_pad = _sig("ta_cast_pad_bf16", _P, _I64, _I64, _P, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _P)
_unpad = _sig("ta_unpad_f32", _P, _I64, _I64, _P, _I64, _I64, _I32, _I32, _I32, _P)
_cast = _sig("ta_cast_bf16", _P, _P, _I64, _P)
_adamw_tick = _sig("ta_adamw_tick", _P, _P, _F64, _F64, _F64, _F64, _P)
_adamw_step = _sig("ta_adamw_step", _P, _P, _P, _P, _P, _I64, _P, _F32, _F32, _F32, _F32, _P)
_adamw_step_bf16g = _sig("ta_adamw_step_bf16g", _P, _P, _P, _P, _P, _I64, _P, _F32, _F32, _F32, _F32, _P)


_drop_set = _sig("ta_dropout_set_step_counter", _P)
_drop_bf16 = _sig("ta_dropout_bf16", _P, _I64, _I64, _P, _I64, _I64, _I32, _I64, _I32, _F32, _U64, _P)
_drop_attn_mask = _sig("ta_dropout_attn_mask", _P, _I32, _I32, _I32, _I32, _I32, _F32, _U64, _P)


class _step_counter:
    """Scope in which launches pick up `counter` (int64 device tensor with one element, or None) as the dropout step
    counter: its value is mixed into every dropout seed when the kernels RUN, so CUDA-graph replays draw fresh masks as
    long as the graph increments it. The library reads the registration at launch time only, on the calling thread."""

    def __init__(self, counter):
        self.counter = counter

    def __enter__(self):
        if self.counter is not None:
            _check(_drop_set(_ptr(self.counter)))

    def __exit__(self, *exc):
        if self.counter is not None:
            _check(_drop_set(C.c_void_p(0)))
        return False


def dropout_bf16(src, dst, p, seed, dropout_step=None):
    """dst = keep ? src / (1 - p) : 0 with the GEMM-epilogue mask of `seed`; src/dst bf16 [G, rows, cols] views."""
    assert src.dim() == 3 and dst.shape == src.shape and src.stride(-1) == 1 and dst.stride(-1) == 1
    G, rows, cols = src.shape
    with _step_counter(dropout_step):
        _check(_drop_bf16(_ptr(src), src.stride(1), src.stride(0), _ptr(dst), dst.stride(1), dst.stride(0), G, rows, cols,
                          float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()))


def dropout_attn_mask(mask, n_src, G, B, H, T, p, seed, dropout_step=None):
    """mask: uint8 [n_src, G, B, H, T, T] contiguous; the keep mask the attention kernels use for `seed` (tests)."""
    assert mask.is_contiguous() and mask.numel() == n_src * G * B * H * T * T
    with _step_counter(dropout_step):
        _check(_drop_attn_mask(_ptr(mask), n_src, G, B, H, T, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()))


def gather_batch(store, ix, T, xb, yb):
    """store int32 [M, N]; ix int64 [B] (device); xb, yb int64 [M, B, T] (training_utils.py:378-379)."""
    M, B = store.shape[0], ix.shape[0]
    assert store.stride(1) == 1 and xb.is_contiguous() and yb.is_contiguous() and xb.shape == (M, B, T)
    _check(_gather_batch(_ptr(store), store.stride(0), M, _ptr(ix), B, T, _ptr(xb), _ptr(yb), _stream()))


def embed_fwd(idx, tok, voff, pos, x, T):
    """idx int64 [M, rows]; tok fp32 [sumV, C]; voff int64 [M+1]; pos fp32 [block, C]; x fp32 [M, rows, C]."""
    M, rows = idx.shape
    Cc = tok.shape[1]
    assert idx.is_contiguous() and tok.is_contiguous() and pos.is_contiguous() and x.is_contiguous()
    _check(_embed_fwd(_ptr(idx), _ptr(tok), _ptr(voff), _ptr(pos), _ptr(x), M, rows, T, Cc, _stream()))


def embed_fwd_at(idx, tok, voff, pos, x, pos_dev):
    """One new position per sequence at index *pos_dev (int32 [1] on the device): idx int64 [M, rows]; x fp32 [M, rows, C]."""
    import torch
    M, rows = idx.shape
    assert idx.is_contiguous() and tok.is_contiguous() and pos.is_contiguous() and x.is_contiguous()
    assert pos_dev.dtype == torch.int32 and pos_dev.numel() == 1
    _check(_embed_fwd_at(_ptr(idx), _ptr(tok), _ptr(voff), _ptr(pos), _ptr(x), M, rows, tok.shape[1], _ptr(pos_dev), _stream()))


def kv_append(src, cache, pos_dev):
    """cache[g, b, *pos_dev, :W] = src[g, b, :W]; src [G, B, W] view (unit inner stride), cache [G, B, capacity, >= W] view of
    the same dtype (bf16 or fp32; W * element size a multiple of 4)."""
    import torch
    G, Bn, W = src.shape
    es = src.element_size()
    assert cache.dim() == 4 and cache.shape[0] == G and cache.shape[1] == Bn and cache.dtype == src.dtype
    assert src.stride(-1) == 1 and cache.stride(-1) == 1 and (W * es) % 4 == 0 and pos_dev.dtype == torch.int32
    k = 4 // es
    for st in (src.stride(0), src.stride(1), cache.stride(0), cache.stride(1), cache.stride(2)):
        assert st % k == 0
    _check(_kv_append(_ptr(src), src.stride(0) // k, src.stride(1) // k, _ptr(cache), cache.stride(0) // k, cache.stride(1) // k,
                      cache.stride(2) // k, G, Bn, W // k, _ptr(pos_dev), _stream()))


def embed_bwd(idx, dx, voff, max_vocab, dtok, dpos, B, T):
    M = idx.shape[0]
    Cc = dx.shape[-1]
    assert idx.is_contiguous() and dx.is_contiguous() and dtok.is_contiguous() and dpos.is_contiguous()
    _check(_embed_bwd(_ptr(idx), _ptr(dx), _ptr(voff), max_vocab, _ptr(dtok), _ptr(dpos), M, B, T, Cc, _stream()))


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps=1e-5):
    """x fp32 [G, rows, C] contiguous; gamma/beta fp32 [G, C]; y bf16 [G, rows, ldy] (ldy = stride of rows)."""
    G, rows, Cc = x.shape
    assert x.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous() and y.stride(-1) == 1
    assert y.stride(0) == rows * y.stride(1)
    _check(_ln_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), y.stride(1), _ptr(mean), _ptr(rstd), G * rows, rows, Cc,
                   eps, _stream()))


def layernorm_fwd_f32(x, gamma, beta, y, mean, rstd, eps=1e-5):
    """layernorm_fwd with an fp32 output y [G, rows, ldy] (bf16x3 precision mode)."""
    import torch
    G, rows, Cc = x.shape
    assert x.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous() and y.stride(-1) == 1
    assert y.dtype == torch.float32 and y.stride(0) == rows * y.stride(1)
    _check(_ln_fwd_f32(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), y.stride(1), _ptr(mean), _ptr(rstd), G * rows, rows, Cc,
                       eps, _stream()))


def _view4(t):
    """(g1 stride, g0 stride, row stride, G1, G0, R, C) of a [..<=2 group dims.., R, C] view with unit inner stride."""
    assert t.dim() >= 2 and t.dim() <= 4 and t.stride(-1) == 1
    gs0, gs1, G0, G1 = _grp(t, 2)
    return gs1, gs0, t.stride(-2), G1, G0, t.shape[-2], t.shape[-1]


SPLIT_A, SPLIT_B = 0b010, 0b100      # lo slot of the three K-concatenated terms: [hi|lo|hi] x [hi|hi|lo]


def split_bf16(src, dst, lo_mask, terms=3, concat_rows=False):
    """src fp32 [.., R, C] -> dst bf16 [.., R, terms*C] (or [.., terms*R, C] with concat_rows): slot k holds lo = bf16(x - hi)
    where bit k of lo_mask is set, hi = bf16(x) elsewhere."""
    import torch
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.dim() == dst.dim()
    s1, s0, sr, G1, G0, R, Cc = _view4(src)
    d1, d0, dr, H1, H0, Rd, Cd = _view4(dst)
    assert (G1, G0) == (H1, H0) and (Rd, Cd) == ((terms * R, Cc) if concat_rows else (R, terms * Cc))
    _check(_split(_ptr(src), s1, s0, sr, _ptr(dst), d1, d0, dr, G1, G0, R, Cc, terms, lo_mask, 1 if concat_rows else 0,
                  _stream()))


def convert(src, dst):
    """dst[...] = src[...] between fp32 and bf16 (either direction, or a same-type strided copy); [..<=2 group dims.., R, C]
    views of equal shape with unit inner stride."""
    import torch
    assert src.shape == dst.shape and src.dtype in (torch.float32, torch.bfloat16) and dst.dtype in (torch.float32, torch.bfloat16)
    s1, s0, sr, G1, G0, R, Cc = _view4(src)
    d1, d0, dr, _, _, _, _ = _view4(dst)
    _check(_convert(_ptr(src), 1 if src.dtype == torch.bfloat16 else 0, s1, s0, sr, _ptr(dst),
                    1 if dst.dtype == torch.bfloat16 else 0, d1, d0, dr, G1, G0, R, Cc, _stream()))


def layernorm_bwd_workspace(G, rows, Cc):
    return int(_ln_bwd_ws(G * rows, rows, Cc))


def layernorm_bwd(dy, x, gamma, mean, rstd, dstream_in, dstream_out, dstream_bf16, dgamma, dbeta, workspace):
    G, rows, Cc = x.shape
    assert x.is_contiguous() and dy.stride(-1) == 1 and dy.stride(0) == rows * dy.stride(1)
    assert dstream_out.is_contiguous() and (dstream_in is None or dstream_in.is_contiguous())
    assert dgamma.is_contiguous() and dbeta.is_contiguous()
    ldb = 0
    if dstream_bf16 is not None:
        assert dstream_bf16.stride(-1) == 1 and dstream_bf16.stride(0) == rows * dstream_bf16.stride(1)
        ldb = dstream_bf16.stride(1)
    assert workspace.numel() >= layernorm_bwd_workspace(G, rows, Cc)
    _check(_ln_bwd(_ptr(dy), dy.stride(1), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dstream_in),
                   _ptr(dstream_out), _ptr(dstream_bf16), ldb, _ptr(dgamma), _ptr(dbeta), _ptr(workspace), G * rows,
                   rows, Cc, _stream()))


def cross_entropy(logits, targets, row_loss, loss_mean, dlogits=None, grad_scale=1.0):
    """logits fp32 [rows, V] (ld = stride 0); targets int64 [rows]; loss_mean fp32 [1]; dlogits bf16 [rows, ldd]."""
    rows, V = logits.shape
    assert logits.stride(1) == 1 and targets.is_contiguous()
    _check(_ce(_ptr(logits), logits.stride(0), _ptr(targets), _ptr(row_loss), _ptr(loss_mean), _ptr(dlogits),
               dlogits.stride(0) if dlogits is not None else 0, rows, V, grad_scale, _stream()))


def colsum_workspace(rows, N, G):
    return int(_colsum_ws(rows, N, G))


def colsum(dy, out, workspace):
    """dy bf16 [G, rows, N] (row stride ld); out fp32 [G, N] (group stride out.stride(0))."""
    G, rows, N = dy.shape
    assert dy.stride(-1) == 1 and out.stride(-1) == 1 and workspace.numel() >= colsum_workspace(rows, N, G)
    _check(_colsum(_ptr(dy), dy.stride(1), dy.stride(0), _ptr(out), out.stride(0), _ptr(workspace), rows, N, G,
                   _stream()))


_aug_ws = _sig("ta_augment_workspace_bytes", _I64, _I64, restype=_I64)
_aug = _sig("ta_augment_tokens", _P, _I64, _P, _I64, _I32, _I32, _P, _P, _P)


def augment_workspace_bytes(N, W):
    return int(_aug_ws(N, W))


def augment_tokens(tokens, words, rand_size, vocab_size, workspace, stats):
    """tokens int32 [N] (contiguous row of the token store, modified in place); words int32/uint32 bits [W]: the next W
    outputs of Python's Mersenne Twister; workspace uint8 [>= augment_workspace_bytes(N, W)]; stats int64 [3]."""
    import torch
    assert tokens.dtype == torch.int32 and tokens.is_contiguous() and words.is_contiguous() and words.element_size() == 4
    assert stats.dtype == torch.int64 and stats.numel() >= 3
    assert workspace.numel() * workspace.element_size() >= augment_workspace_bytes(tokens.numel(), words.numel())
    _check(_aug(_ptr(tokens), tokens.numel(), _ptr(words), words.numel(), int(rand_size), int(vocab_size), _ptr(workspace),
                _ptr(stats), _stream()))


_dir_metrics = _sig("ta_directional_metrics", _P, _I64, _P, _P, _I64, _P, _I64, _I32, _I32, _I32, _P, _P, _P)


def directional_metrics(last_logits, vocab_values, prev_tokens, actual_tokens, is_percentage, counts, certainty):
    """last_logits fp32 [B, V] view (last position; any row stride); vocab_values fp64 [V]; prev_tokens / actual_tokens
    int64 [B] views (any stride; prev may be None for percentage data); counts int64 [2] and certainty fp64 [1] are
    device accumulators (+=)."""
    import torch
    Bn, V = last_logits.shape
    assert last_logits.dtype == torch.float32 and last_logits.stride(1) == 1 and vocab_values.dtype == torch.float64
    assert vocab_values.is_contiguous() and vocab_values.numel() == V
    assert actual_tokens.dtype == torch.int64 and counts.dtype == torch.int64 and certainty.dtype == torch.float64
    _check(_dir_metrics(_ptr(last_logits), last_logits.stride(0), _ptr(vocab_values), _ptr(prev_tokens),
                        prev_tokens.stride(0) if prev_tokens is not None else 0, _ptr(actual_tokens), actual_tokens.stride(0),
                        Bn, V, 1 if is_percentage else 0, _ptr(counts), _ptr(certainty), _stream()))


def cast_pad(src, dst):
    """src fp32 [G, R, Cs] (inner stride 1), dst bf16 [G, Rd, Cd] with Rd >= R, Cd >= Cs: zero padded cast."""
    G, R, Cs = src.shape
    _, Rd, Cd = dst.shape
    assert src.stride(2) == 1 and dst.stride(2) == 1 and dst.shape[0] == G and Rd >= R and Cd >= Cs
    _check(_cast_pad(_ptr(src), src.stride(0), src.stride(1), _ptr(dst), dst.stride(0), dst.stride(1), G, R, Cs, Rd, Cd,
                     _stream()))


def unpad(src, dst):
    """dst fp32 [G, R, Cn] = src fp32 [G, >=R, >=Cn][:, :R, :Cn]."""
    G, R, Cn = dst.shape
    assert src.stride(2) == 1 and dst.stride(2) == 1
    _check(_unpad(_ptr(src), src.stride(0), src.stride(1), _ptr(dst), dst.stride(0), dst.stride(1), G, R, Cn, _stream()))


def cast_bf16(src, dst):
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    _check(_cast(_ptr(src), _ptr(dst), src.numel(), _stream()))


def adamw_tick(step_dev, scalars_dev, lr, beta1=0.9, beta2=0.999, weight_decay=0.01):
    _check(_adamw_tick(_ptr(step_dev), _ptr(scalars_dev), lr, beta1, beta2, weight_decay, _stream()))


def adamw_step(p, g, m, v, shadow, scalars_dev, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    import torch
    n = p.numel()
    assert g.numel() == n and m.numel() == n and v.numel() == n and (shadow is None or shadow.numel() == n)
    if g.dtype == torch.bfloat16:        # the exchanged bf16 copy of the gradient (data-parallel step)
        _check(_adamw_step_bf16g(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(shadow), n, _ptr(scalars_dev), beta1, beta2, eps,
                                 grad_scale, _stream()))
        return
    _check(_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(shadow), n, _ptr(scalars_dev), beta1, beta2, eps,
                       grad_scale, _stream()))


_lib.ta_attention_bwd.restype = C.c_int32
_lib.ta_attention_bwd.argtypes = [C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor), C.POINTER(TaAttnTensor),
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_void_p]
_attn_bwd_ws_bytes = _sig("ta_attention_bwd_workspace_bytes", _I32, _I32, _I32, _I32, _I32, restype=_I64)
_attn_bwd_ws = {}      # (device, bytes) -> fp32 tensor; never freed: captured CUDA graphs keep replaying into these addresses


def attention_bwd_workspace(n_src, G, B, T, H, device):
    """The D = rowsum(dO * O_s) table of ta_attention_bwd (fp32 [n_src, G, B, H, T]); one per size and device, shared by
    every call on the compute stream (the table is consumed by the launch that follows the one that writes it)."""
    import torch
    nbytes = int(_attn_bwd_ws_bytes(n_src, G, B, T, H))
    key = (str(device), nbytes)
    ws = _attn_bwd_ws.get(key)
    if ws is None:
        ws = _attn_bwd_ws[key] = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
    return ws


def attention_bwd(q, ks, vs, o, d_o, lse, dq, dks, dvs, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None,
                  o_src=None, dropout_p=0.0, dropout_seed=0, dropout_step=None, workspace=None):
    """Gradients of attention_fwd. d_o has the layout of o; dq of q; dks[s]/dvs[s] of ks[s]/vs[s] (bf16)."""
    n_src = len(ks)
    G = q.shape[0] if q.dim() == 3 else 1
    if workspace is None:
        workspace = attention_bwd_workspace(n_src, G, B, T, H, q.device)
    assert workspace.dtype.itemsize * workspace.numel() >= int(_attn_bwd_ws_bytes(n_src, G, B, T, H)) and workspace.is_contiguous()
    hq = head_stride_q if head_stride_q is not None else hs
    hkv = head_stride_kv if head_stride_kv is not None else hs
    qa = (TaAttnTensor * 1)(_attn_tensor(q, hq))
    ka = (TaAttnTensor * n_src)(*[_attn_tensor(k, hkv) for k in ks])
    va = (TaAttnTensor * n_src)(*[_attn_tensor(v, hkv) for v in vs])
    dqa = (TaAttnTensor * 1)(_attn_tensor(dq, hq))
    dka = (TaAttnTensor * n_src)(*[_attn_tensor(k, hkv) for k in dks])
    dva = (TaAttnTensor * n_src)(*[_attn_tensor(v, hkv) for v in dvs])
    assert o.stride(-1) == 1 and d_o.stride(-1) == 1 and d_o.stride(-2) == o.stride(-2)
    o_gs = o.stride(0) if o.dim() == 3 else 0
    assert o.dim() < 3 or d_o.stride(0) == o.stride(0)
    with _step_counter(dropout_step if dropout_p > 0 else None):
        _check(_lib.ta_attention_bwd(qa, ka, va, n_src, _ptr(o), _ptr(o_src), o_src.stride(0) if o_src is not None else 0,
                                     _ptr(d_o), o.stride(-2), o_gs, _ptr(lse), _ptr(workspace), dqa, dka, dva, G, B, T, H, hs, float(dropout_p),
                                     int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _stream()))


# ---------------------------------------------------------------------------------------------------- K9: data preparation
PREP_STATUS_WORDS = 4
PREP_NONFINITE, PREP_ZERO_PREV, PREP_OVERFLOW, PREP_TABLE_FULL = 0, 1, 2, 3
PREP_NO_ROW = 2 ** 63 - 1
PREP_MAX_RANKS = 4
PREP_MAX_GROUPS = 8192
PREP_DECADES = 633            # decades -324 .. 308 of the threshold / factor tables
PREP_DECADE_MIN = -324
PREP_MAX_DECIMALS = 22

_prep_reset = _sig("ta_prep_status_reset", _P, _P)
_prep_pct = _sig("ta_prep_percent_changes", _P, _P, _I64, _P, _I32, _I32, _P, _P)
_prep_range = _sig("ta_prep_range", _P, _P, _I64, _I32, _I32, _F64, _F64, _F64, _P, _P, _P, _P)
_prep_range_auto_pb = _sig("ta_prep_range_auto_params_bytes", restype=_I64)
_prep_range_auto_p = _sig("ta_prep_range_auto_params", _P, _I32, _F64, _F64, _P)
_prep_range_auto = _sig("ta_prep_range_auto", _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P)
_prep_sel_ws = _sig("ta_prep_select_workspace_bytes", restype=_I64)
_prep_sel = _sig("ta_prep_order_statistics", _P, _I64, _P, _I32, _P, _P, _P, _P)
_prep_mm_ws = _sig("ta_prep_minmax_workspace_bytes", _I64, restype=_I64)
_prep_mm = _sig("ta_prep_filtered_minmax", _P, _I64, _F64, _F64, _P, _P, _P)
_prep_bin = _sig("ta_prep_bin_assign", _P, _P, _I64, _P, _P, _I32, _I32, _P)
_prep_vins = _sig("ta_prep_vocab_insert", _P, _I64, _P, _I64, _P, _P, _P)
_prep_vbuild = _sig("ta_prep_vocab_build", _P, _I64, _I64, _P, _P, _P, _P, _P, _P)
_prep_vlook = _sig("ta_prep_vocab_lookup", _P, _I64, _P, _I64, _P, _P, _P)


def _f64(t):
    import torch
    assert t.dtype == torch.float64 and t.is_contiguous() and t.dim() == 1, "data columns are contiguous float64 [N]"
    return t


def prep_status_reset(status):
    import torch
    assert status.dtype == torch.int64 and status.numel() >= PREP_STATUS_WORDS
    _check(_prep_reset(_ptr(status), _stream()))


def prep_percent_changes(x, out, file_starts, decimals, status):
    """x, out float64 [N] (distinct); file_starts int64 [F] device (first row of every file, [0] = 0)."""
    import torch
    assert file_starts.dtype == torch.int64 and file_starts.is_contiguous() and out.numel() == x.numel()
    _check(_prep_pct(_ptr(_f64(x)), _ptr(_f64(out)), x.numel(), _ptr(file_starts), file_starts.numel(), int(decimals),
                     _ptr(status), _stream()))


def prep_range(x, out, has_whole_digits, decimals, lower, upper, upper_clamp, thresholds, factors, status):
    """x, out float64 [N] (may alias); thresholds / factors float64 [633] device tables (None without whole digits)."""
    assert out.numel() == x.numel()
    if has_whole_digits:
        assert thresholds.numel() == PREP_DECADES and factors.numel() == PREP_DECADES
        _f64(thresholds), _f64(factors)
    _check(_prep_range(_ptr(_f64(x)), _ptr(_f64(out)), x.numel(), 1 if has_whole_digits else 0, int(decimals), float(lower),
                       float(upper), float(upper_clamp), _ptr(thresholds), _ptr(factors), _ptr(status), _stream()))


def prep_range_auto(x, is_int, out, has_whole_digits, lower, upper, upper_clamps, thresholds, factors, status, n_flagged):
    """range_numeric_data with decimal_places=None. is_int: uint8 [N] device mask or None; upper_clamps: 23 python floats;
    n_flagged: int64 [1] device counter (+=) of the elements left as NaN for the host (exponent-notation magnitudes)."""
    import torch
    assert out.numel() == x.numel() and len(upper_clamps) == PREP_MAX_DECIMALS + 1 and n_flagged.dtype == torch.int64
    host = (C.c_uint8 * int(_prep_range_auto_pb()))()
    clamps = (C.c_double * (PREP_MAX_DECIMALS + 1))(*[float(v) for v in upper_clamps])
    _check(_prep_range_auto_p(host, 1 if has_whole_digits else 0, float(lower), float(upper), clamps))
    params = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(x.device)
    if has_whole_digits:
        _f64(thresholds), _f64(factors)
    if is_int is not None:
        assert is_int.dtype == torch.uint8 and is_int.is_contiguous() and is_int.numel() == x.numel()
    _check(_prep_range_auto(_ptr(_f64(x)), _ptr(is_int), _ptr(_f64(out)), x.numel(), _ptr(params), _ptr(thresholds),
                            _ptr(factors), _ptr(status), _ptr(n_flagged), _stream()))
    return params       # keep alive until the launch has run (the caller synchronises when it reads the status)


def prep_select_workspace_bytes():
    return int(_prep_sel_ws())


def prep_order_statistics(x, ranks, out, workspace, status):
    """out[r] = the ranks[r]-th smallest (0-based) element of x; ranks int64 [R <= 4] and out float64 [R] on the device."""
    import torch
    assert ranks.dtype == torch.int64 and out.dtype == torch.float64 and 1 <= ranks.numel() <= PREP_MAX_RANKS
    assert out.numel() >= ranks.numel() and workspace.numel() * workspace.element_size() >= prep_select_workspace_bytes()
    _check(_prep_sel(_ptr(_f64(x)), x.numel(), _ptr(ranks), ranks.numel(), _ptr(out), _ptr(workspace), _ptr(status), _stream()))


def prep_minmax_workspace_bytes(N):
    return int(_prep_mm_ws(int(N)))


def prep_filtered_minmax(x, lo, hi, out3, workspace):
    """out3 float64 [3] = {min, max, count} over the rows with lo <= x <= hi."""
    import torch
    assert out3.dtype == torch.float64 and out3.numel() >= 3
    assert workspace.numel() * workspace.element_size() >= prep_minmax_workspace_bytes(x.numel())
    _check(_prep_mm(_ptr(_f64(x)), x.numel(), float(lo), float(hi), _ptr(out3), _ptr(workspace), _stream()))


def prep_bin_assign(x, out, pos_bounds, neg_bounds, bounds_sorted=False):
    """out[i] = group of x[i] as a double; pos_bounds / neg_bounds float64 [G+1] on the device; bounds_sorted: the caller
    checked that both are non-decreasing (bisection instead of the reference's linear walk, same groups)."""
    G = pos_bounds.numel() - 1
    assert neg_bounds.numel() == G + 1 and out.numel() == x.numel()
    _check(_prep_bin(_ptr(_f64(x)), _ptr(_f64(out)), x.numel(), _ptr(_f64(pos_bounds)), _ptr(_f64(neg_bounds)), G,
                     1 if bounds_sorted else 0, _stream()))


def prep_vocab_insert(x, table, counters, status):
    """table int64/uint64 bits [capacity = 2^k] (cleared here); counters int64 [2] (cleared here; [0] = distinct values)."""
    import torch
    assert table.element_size() == 8 and table.is_contiguous() and counters.dtype == torch.int64 and counters.numel() >= 2
    _check(_prep_vins(_ptr(_f64(x)), x.numel(), _ptr(table), table.numel(), _ptr(counters), _ptr(status), _stream()))


def prep_vocab_build(table, n_distinct, keys_ws, slots_ws, vocabulary, token_of_slot, counters):
    import torch
    assert keys_ws.element_size() == 8 and keys_ws.numel() >= n_distinct and slots_ws.dtype == torch.int32
    assert slots_ws.numel() >= n_distinct and vocabulary.dtype == torch.float64 and vocabulary.numel() >= n_distinct
    assert token_of_slot.dtype == torch.int32 and token_of_slot.numel() >= table.numel()
    _check(_prep_vbuild(_ptr(table), table.numel(), int(n_distinct), _ptr(keys_ws), _ptr(slots_ws), _ptr(vocabulary),
                        _ptr(token_of_slot), _ptr(counters), _stream()))


def prep_vocab_lookup(x, table, token_of_slot, tokens):
    import torch
    assert tokens.dtype == torch.int32 and tokens.is_contiguous() and tokens.numel() == x.numel()
    _check(_prep_vlook(_ptr(_f64(x)), x.numel(), _ptr(table), table.numel(), _ptr(token_of_slot), _ptr(tokens), _stream()))
