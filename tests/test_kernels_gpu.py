"""Per-kernel parity on the GPU: every C-ABI op against a plain PyTorch fp32 statement of the same op (bit-exact for
integer work; fp tolerances written next to each check)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from tradeaid_b200 import lib as L
    return L


def _r8(n):
    return (n + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # M, N, K, groups, a_major, b_major, epilogue
    (128, 64, 64, (), 0, 0, "none"),
    (100, 72, 40, (), 0, 0, "none"),
    (16, 28, 32, (2,), 0, 0, "none"),
    (16, 3, 1, (2,), 0, 0, "bias"),          # demo head: V = 3, V // 2 = 1
    (256, 512, 512, (), 0, 1, "none"),
    (256, 512, 512, (), 1, 1, "none"),
    (200, 136, 72, (2, 3), 0, 1, "none"),
    (200, 136, 72, (2, 3), 1, 1, "none"),
    (300, 200, 136, (3, 2), 0, 0, "bias_tanh"),
    (300, 200, 136, (2,), 0, 0, "bias_relu"),
    (300, 200, 136, (), 0, 0, "res"),
    (300, 200, 136, (), 0, 1, "tanh_bwd"),
    (300, 200, 136, (), 0, 1, "relu_bwd"),
    (4096, 1024, 1024, (), 0, 0, "bias_relu"),
    (1024, 512, 4096, (), 1, 1, "none"),
]


@pytest.mark.parametrize("M,N,K,g,am,bm,epi", GEMM_CASES)
def test_gemm(lib, M, N, K, g, am, bm, epi):
    torch.manual_seed(0)
    dev = "cuda"

    def mk(rows, cols):
        return torch.randn(*g, rows, _r8(cols) + 8, device=dev).to(torch.bfloat16)[..., :cols]

    A = mk(M, K) if am == 0 else mk(K, M)
    B = mk(N, K) if bm == 0 else mk(K, N)
    Af = A.float() if am == 0 else A.float().transpose(-1, -2)
    Bf = B.float() if bm == 0 else B.float().transpose(-1, -2)
    ref = Af @ Bf.transpose(-1, -2)
    kw = {}
    if epi in ("bias", "bias_tanh", "bias_relu"):
        bias = torch.randn(*g, N, device=dev)
        act = {"bias": lib.ACT_NONE, "bias_tanh": lib.ACT_TANH, "bias_relu": lib.ACT_RELU}[epi]
        kw.update(bias=bias, act=act, alpha=0.05)
        ref = ref * 0.05 + bias[..., None, :]
        ref = torch.tanh(ref) if epi == "bias_tanh" else (torch.relu(ref) if epi == "bias_relu" else ref)
    elif epi == "res":
        res, bias = torch.randn(*g, M, N, device=dev), torch.randn(*g, N, device=dev)
        kw.update(residual=res, bias=bias)
        ref = ref + bias[..., None, :] + res
    elif epi in ("tanh_bwd", "relu_bwd"):
        aux = torch.randn(*g, M, _r8(N), device=dev).to(torch.bfloat16)[..., :N]
        kw.update(aux=aux, aux_mode=lib.AUX_TANH_BWD if epi == "tanh_bwd" else lib.AUX_RELU_BWD)
        ref = ref * (1 - aux.float() ** 2) if epi == "tanh_bwd" else torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    of = torch.full((*g, M, N + 3), float("nan"), device=dev)[..., :N]
    ob = torch.zeros(*g, M, _r8(N), device=dev, dtype=torch.bfloat16)[..., :N]
    lib.gemm(A, B, a_major=am, b_major=bm, out_f32=of, out_bf16=ob, **kw)
    torch.cuda.synchronize()
    scale = ref.abs().max().item() + 1e-6
    # operands are exactly representable bf16, accumulation fp32: only summation-order error remains
    assert (of - ref).abs().max().item() <= 1e-4 * scale + 1e-5
    assert (ob.float() - ref).abs().max().item() <= 8e-3 * scale + 1e-3   # bf16 output rounding (2^-8 relative)


GEMM_TMA_CASES = [
    # M, N, K, groups, a_major, b_major, epilogue, outputs   (16-byte pitches everywhere: TMA-staged epilogue)
    (128, 64, 64, (), 0, 0, "none", "bf16"),
    (128, 64, 64, (), 0, 0, "none", "f32"),
    (100, 72, 40, (), 0, 0, "bias_relu", "bf16"),
    (100, 72, 40, (2,), 0, 0, "bias_tanh", "f32"),
    (16, 24, 32, (2,), 0, 0, "none", "bf16"),           # N < one 32-column box
    (300, 200, 136, (3, 2), 0, 0, "bias_tanh", "bf16"),
    (300, 200, 136, (2,), 0, 0, "res", "f32"),
    (300, 200, 136, (2,), 0, 0, "res_inplace", "f32"),
    (300, 200, 136, (), 0, 0, "res", "both"),
    (300, 200, 136, (), 0, 0, "bias", "both"),
    (300, 200, 136, (2,), 0, 1, "tanh_bwd", "bf16"),
    (300, 200, 136, (), 0, 1, "relu_bwd", "bf16"),
    (1000, 520, 264, (), 0, 0, "res", "f32"),            # several tiles per CTA along N and M
    (4096, 1024, 1024, (2,), 0, 0, "bias_relu", "bf16"),
    (4096, 1024, 512, (2,), 0, 0, "res", "f32"),          # more tiles than SMs: buffer rotation across tiles
    (4096, 1024, 1024, (), 0, 1, "relu_bwd", "bf16"),
    (1024, 512, 4096, (3,), 1, 1, "none", "f32"),
    (2048, 64, 32, (2, 24), 0, 0, "none", "bf16"),        # block-diagonal q/k/v stage 2 shape
]


@pytest.mark.parametrize("M,N,K,g,am,bm,epi,outs", GEMM_TMA_CASES)
def test_gemm_tma_epilogue(lib, M, N, K, g, am, bm, epi, outs):
    torch.manual_seed(1)
    dev = "cuda"

    def mk(rows, cols):
        return torch.randn(*g, rows, _r8(cols) + 8, device=dev).to(torch.bfloat16)[..., :cols]

    A = mk(M, K) if am == 0 else mk(K, M)
    B = mk(N, K) if bm == 0 else mk(K, N)
    Af = A.float() if am == 0 else A.float().transpose(-1, -2)
    Bf = B.float() if bm == 0 else B.float().transpose(-1, -2)
    ref = Af @ Bf.transpose(-1, -2)
    kw = {}
    ldf = _r8(N) + 4
    of = torch.full((*g, M, ldf), float("nan"), device=dev)[..., :N] if outs in ("f32", "both") else None
    ob = torch.zeros(*g, M, _r8(N) + 8, device=dev, dtype=torch.bfloat16)[..., :N] if outs in ("bf16", "both") else None
    if epi in ("bias", "bias_tanh", "bias_relu"):
        bias = torch.randn(*g, N, device=dev)
        act = {"bias": lib.ACT_NONE, "bias_tanh": lib.ACT_TANH, "bias_relu": lib.ACT_RELU}[epi]
        kw.update(bias=bias, act=act, alpha=0.05)
        ref = ref * 0.05 + bias[..., None, :]
        ref = torch.tanh(ref) if epi == "bias_tanh" else (torch.relu(ref) if epi == "bias_relu" else ref)
    elif epi == "res":
        res, bias = torch.randn(*g, M, N + 4, device=dev)[..., :N], torch.randn(*g, N, device=dev)
        kw.update(residual=res, bias=bias)
        ref = ref + bias[..., None, :] + res
    elif epi == "res_inplace":      # engine.py: dx = dx + dkv @ Wkv (residual and output are the same buffer)
        of.copy_(torch.randn(*g, M, N, device=dev))
        kw.update(residual=of)
        ref = ref + of
    elif epi in ("tanh_bwd", "relu_bwd"):
        aux = torch.randn(*g, M, _r8(N), device=dev).to(torch.bfloat16)[..., :N]
        kw.update(aux=aux, aux_mode=lib.AUX_TANH_BWD if epi == "tanh_bwd" else lib.AUX_RELU_BWD)
        ref = ref * (1 - aux.float() ** 2) if epi == "tanh_bwd" else torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    lib.gemm(A, B, a_major=am, b_major=bm, out_f32=of, out_bf16=ob, **kw)
    torch.cuda.synchronize()
    scale = ref.abs().max().item() + 1e-6
    if of is not None:
        assert (of - ref).abs().max().item() <= 1e-4 * scale + 1e-5
        pad = of.as_strided((*g, M, ldf), of.stride())[..., N:]          # the pitch padding must stay untouched
        assert torch.isnan(pad).all() or epi == "res_inplace"
    if ob is not None:
        assert (ob.float() - ref).abs().max().item() <= 8e-3 * scale + 1e-3
        pad = ob.as_strided((*g, M, _r8(N) + 8), ob.stride())[..., N:]
        assert (pad == 0).all()


# ------------------------------------------------------------------------------------------------ get_batch gather
@pytest.mark.parametrize("M,N,B,T", [(2, 1000, 4, 4), (8, 100000, 64, 256), (3, 50, 1, 49)])
def test_gather_batch_bit_exact(lib, M, N, B, T):
    g = torch.Generator().manual_seed(1)
    store = torch.randint(0, 2 ** 31 - 1, (M, N), generator=g, dtype=torch.int64).to(torch.int32)
    ix = torch.randint(0, N - T, (B,), generator=g)
    xb = torch.empty(M, B, T, dtype=torch.int64, device="cuda")
    yb = torch.empty_like(xb)
    lib.gather_batch(store.cuda(), ix.cuda(), T, xb, yb)
    ar = torch.arange(T)
    ref_x = store.long()[:, ix[:, None] + ar]
    ref_y = store.long()[:, ix[:, None] + ar + 1]
    assert torch.equal(xb.cpu(), ref_x) and torch.equal(yb.cpu(), ref_y)


# ------------------------------------------------------------------------------------------------ embedding
@pytest.mark.parametrize("V,B,T,C", [([57, 3], 4, 4, 32), ([1024, 1024, 512], 16, 64, 256), ([30], 2, 5, 36)])
def test_embedding_fwd_bwd(lib, V, B, T, C):
    torch.manual_seed(2)
    M = len(V)
    voff = torch.tensor([0] + list(torch.tensor(V).cumsum(0)), dtype=torch.int64, device="cuda")
    tok = torch.randn(sum(V), C, device="cuda")
    pos = torch.randn(T + 3, C, device="cuda")
    idx = torch.stack([torch.randint(0, V[m], (B * T,)) for m in range(M)]).cuda()
    x = torch.empty(M, B * T, C, device="cuda")
    lib.embed_fwd(idx, tok, voff, pos, x, T)
    ref = torch.stack([tok[voff[m] + idx[m]] + pos[:T].repeat(B, 1) for m in range(M)])
    assert torch.equal(x, ref)  # one fp32 add per element: bit-exact
    dx = torch.randn(M, B * T, C, device="cuda")
    dtok = torch.full_like(tok, float("nan"))
    dpos = torch.zeros(T, C, device="cuda")
    lib.embed_bwd(idx, dx, voff, max(V), dtok, dpos, B, T)
    ref_tok = torch.zeros_like(tok)
    for m in range(M):
        ref_tok.index_add_(0, voff[m] + idx[m], dx[m])
    assert torch.allclose(dtok, ref_tok, rtol=1e-5, atol=1e-5)
    ref_pos = dx.view(M, B, T, C).sum((0, 1))
    assert torch.allclose(dpos, ref_pos, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("G,rows,C", [(2, 16, 32), (3, 192, 256), (1, 100, 512), (2, 128, 1024), (1, 40, 36)])
def test_layernorm_fwd_bwd(lib, G, rows, C):
    torch.manual_seed(3)
    x = (torch.randn(G, rows, C, device="cuda") * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(G, C, device="cuda")).requires_grad_(True)
    beta = (0.1 * torch.randn(G, C, device="cuda")).requires_grad_(True)
    ld = _r8(C) + 8
    y = torch.zeros(G, rows, ld, device="cuda", dtype=torch.bfloat16)[..., :C]
    mean = torch.empty(G * rows, device="cuda")
    rstd = torch.empty(G * rows, device="cuda")
    lib.layernorm_fwd(x.detach(), gamma.detach(), beta.detach(), y, mean, rstd)
    ref = torch.stack([F.layer_norm(x[g], (C,), gamma[g], beta[g], 1e-5) for g in range(G)])
    assert (y.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-6   # bf16 rounding of y
    assert torch.allclose(mean.view(G, rows), x.mean(-1), atol=1e-5)
    assert torch.allclose(rstd.view(G, rows), (x.var(-1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-5)
    dy = torch.randn(G, rows, ld, device="cuda").to(torch.bfloat16)[..., :C]
    din = torch.randn(G, rows, C, device="cuda")
    ref.backward(dy.float())
    dout = torch.empty_like(din)
    dbf = torch.zeros(G, rows, ld, device="cuda", dtype=torch.bfloat16)[..., :C]
    dgamma = torch.empty(G, C, device="cuda")
    dbeta = torch.empty(G, C, device="cuda")
    ws = torch.empty(lib.layernorm_bwd_workspace(G, rows, C), device="cuda")
    lib.layernorm_bwd(dy, x.detach(), gamma.detach(), mean, rstd, din, dout, dbf, dgamma, dbeta, ws)
    sc = x.grad.abs().max().item()
    assert (dout - (din + x.grad)).abs().max().item() <= 1e-5 * sc + 1e-5
    assert (dbf.float() - dout).abs().max().item() <= 2 ** -8 * dout.abs().max().item()
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-4, atol=1e-4 * gamma.grad.abs().max().item())
    assert torch.allclose(dbeta, beta.grad, rtol=1e-4, atol=1e-4 * beta.grad.abs().max().item())
    # dstream_in = None (first contribution to the stream gradient)
    lib.layernorm_bwd(dy, x.detach(), gamma.detach(), mean, rstd, None, dout, None, dgamma, dbeta, ws)
    assert (dout - x.grad).abs().max().item() <= 1e-5 * sc + 1e-5


# ------------------------------------------------------------------------------------------------ cross-entropy
@pytest.mark.parametrize("rows,V", [(16, 57), (16, 3), (4096, 1024), (100, 1000)])
def test_cross_entropy(lib, rows, V):
    torch.manual_seed(4)
    ldl = V + 5
    logits = (torch.randn(rows, ldl, device="cuda") * 3)[:, :V]
    tgt = torch.randint(0, V, (rows,), device="cuda")
    row_loss = torch.empty(rows, device="cuda")
    loss = torch.empty(1, device="cuda")
    dl = torch.zeros(rows, _r8(V), device="cuda", dtype=torch.bfloat16)[:, :V]
    lib.cross_entropy(logits, tgt, row_loss, loss, dl, grad_scale=1.0)
    lg = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lg, tgt)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-6
    sc = lg.grad.abs().max().item()
    assert (dl.float() - lg.grad).abs().max().item() <= 2 ** -8 * sc + 1e-9
    lib.cross_entropy(logits, tgt, row_loss, loss, None)
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-6


# ------------------------------------------------------------------------------------------------ colsum / casts
@pytest.mark.parametrize("G,rows,N", [(2, 16, 28), (3, 600, 136), (1, 8192, 512), (2, 5000, 65)])
def test_colsum(lib, G, rows, N):
    torch.manual_seed(5)
    dy = torch.randn(G, rows, _r8(N) + 8, device="cuda").to(torch.bfloat16)[..., :N]
    out = torch.empty(G, N + 3, device="cuda")[:, :N]
    ws = torch.empty(lib.colsum_workspace(rows, N, G), device="cuda")
    lib.colsum(dy, out, ws)
    ref = dy.float().sum(1)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-4 * math.sqrt(rows))


def test_casts(lib):
    torch.manual_seed(6)
    src = torch.randn(5, 7, 13, device="cuda")
    dst = torch.full((5, 8, 16), 7.0, device="cuda", dtype=torch.bfloat16)
    lib.cast_pad(src, dst)
    ref = torch.zeros(5, 8, 16, device="cuda")
    ref[:, :7, :13] = src
    assert torch.equal(dst, ref.to(torch.bfloat16))
    back = torch.empty(5, 7, 13, device="cuda")
    lib.unpad(ref, back)
    assert torch.equal(back, src)
    flat = torch.randn(100003, device="cuda")
    out = torch.empty(100003, device="cuda", dtype=torch.bfloat16)
    lib.cast_bf16(flat, out)
    assert torch.equal(out, flat.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ AdamW
def test_adamw_matches_torch(lib):
    torch.manual_seed(7)
    n = 100003
    p0 = torch.randn(n, device="cuda")
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    sc = torch.zeros(4, device="cuda")
    for _ in range(5):
        g = torch.randn(n, device="cuda") * 0.1
        ref_p.grad = g.clone()
        opt.step()
        lib.adamw_tick(step, sc, 1e-3)
        lib.adamw_step(p, g, m, v, shadow, sc)
    assert step.item() == 5
    assert (p - ref_p.detach()).abs().max().item() <= 2e-6      # fp32, same operation order as torch's AdamW
    assert torch.equal(shadow, p.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ attention forward
def _ref_attention(q, ks, vs, B, T, H, hs):
    """q [G, B*T, H*hs] fp32; ks/vs lists of same. Returns o [G, B*T, H*hs], per-source o, lse [S, G, B, H, T]."""
    G = q.shape[0]
    qh = q.view(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
    mask = torch.ones(T, T, dtype=torch.bool, device=q.device).tril()
    outs, lses = [], []
    for k, v in zip(ks, vs):
        kh = k.reshape(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
        vh = v.reshape(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
        s = (qh @ kh.transpose(-1, -2)) * hs ** -0.5
        s = s.masked_fill(~mask, float("-inf"))
        lses.append(torch.logsumexp(s, -1))
        outs.append((torch.softmax(s, -1) @ vh).permute(0, 1, 3, 2, 4).reshape(G, B * T, H * hs))
    return sum(outs), torch.stack(outs), torch.stack(lses)


ATTN_CASES = [
    # G, B, T, H, hs, n_src
    (1, 2, 4, 4, 8, 1),       # demo shape
    (2, 3, 64, 8, 32, 1),     # c2 self
    (1, 2, 128, 8, 64, 1),    # c3 self
    (2, 2, 256, 4, 64, 1),    # c4 self
    (1, 2, 256, 2, 64, 3),    # cross, 3 sources
    (1, 1, 1024, 2, 64, 2),   # c5 long context
    (1, 2, 100, 2, 16, 2),    # ragged T
    (1, 3, 5, 2, 8, 1),       # T < block_size (generate)
    (1, 2, 37, 3, 24, 2),     # odd T, head size that is no power of two
]


@pytest.mark.parametrize("G,B,T,H,hs,S", ATTN_CASES)
def test_attention_fwd(lib, G, B, T, H, hs, S):
    torch.manual_seed(8)
    C = H * hs
    q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    # k and v of a source interleaved per head like the cross-attention kv projection output: [.., H, 2*hs]
    kv = [torch.randn(G, B * T, H, 2 * hs, device="cuda").to(torch.bfloat16) for _ in range(S)]
    ks = [t[..., :hs] for t in kv]
    vs = [t[..., hs:] for t in kv]
    o = torch.zeros(G, B * T, C, device="cuda", dtype=torch.bfloat16)
    o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(S, G, B, H, T, device="cuda")
    k_views = [t.view(G, B * T, 2 * C) for t in kv]
    v_views = [t.view(G, B * T, 2 * C)[..., hs:] for t in kv]
    lib.attention_fwd(q, k_views, v_views, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src)
    torch.cuda.synchronize()
    ro, ros, rl = _ref_attention(q.float(), [k.float() for k in ks], [v.float() for v in vs], B, T, H, hs)
    # P is rounded to bf16 before P V and o is stored in bf16: 2^-8 relative each
    assert (o.float() - ro).abs().max().item() <= 1.5e-2 * ro.abs().max().item()
    assert (o_src.float() - ros).abs().max().item() <= 1.5e-2 * ros.abs().max().item()
    assert (lse - rl).abs().max().item() <= 2e-3


ATTN_BWD_CASES = [c for c in ATTN_CASES if c[2] <= 256]


@pytest.mark.parametrize("G,B,T,H,hs,S", ATTN_BWD_CASES)
def test_attention_bwd(lib, G, B, T, H, hs, S):
    torch.manual_seed(9)
    C = H * hs
    q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    kv = [torch.randn(G, B * T, H, 2 * hs, device="cuda").to(torch.bfloat16) for _ in range(S)]
    d_o = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    o = torch.zeros(G, B * T, C, device="cuda", dtype=torch.bfloat16)
    o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(S, G, B, H, T, device="cuda")
    k_views = [t.view(G, B * T, 2 * C) for t in kv]
    v_views = [t.view(G, B * T, 2 * C)[..., hs:] for t in kv]
    lib.attention_fwd(q, k_views, v_views, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src)
    dq = torch.full_like(q, float("nan"))
    dkv = [torch.full((G, B * T, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(S)]
    lib.attention_bwd(q, k_views, v_views, o, d_o, lse, dq, dkv, [t[..., hs:] for t in dkv], B=B, T=T, H=H, hs=hs,
                      head_stride_kv=2 * hs, o_src=o_src if S > 1 else None)
    torch.cuda.synchronize()
    qf = q.float().requires_grad_(True)
    kvf = [t.float().requires_grad_(True) for t in kv]
    ro, _, _ = _ref_attention(qf, [t[..., :hs] for t in kvf], [t[..., hs:] for t in kvf], B, T, H, hs)
    ro.backward(d_o.float())
    # P, dS are rounded to bf16 before the gradient UMMAs and outputs are bf16: ~2^-8 relative per rounding
    tol = 2e-2
    assert (dq.float() - qf.grad).abs().max().item() <= tol * qf.grad.abs().max().item()
    for s in range(S):
        ref = kvf[s].grad.view(G, B * T, 2 * C)
        assert (dkv[s].float() - ref).abs().max().item() <= tol * ref.abs().max().item()


# ------------------------------------------------------------------------------------------------ dropout
# The reference draws its dropout masks from torch's own Philox stream (nn.Dropout, model.py:69, 91, 116, 151, 171), which
# cannot be reproduced outside torch; the kernels use their own counter-based Philox4x32-7 masks. Parity is therefore
# checked against a reference that applies the SAME mask (read back through the mask kernels), plus the mask statistics.
def test_dropout_mask_statistics(lib):
    G, rows, cols, p = 3, 1000, 264, 0.2
    src = torch.ones(G, rows, cols + 8, device="cuda", dtype=torch.bfloat16)[..., :cols]
    dst = torch.full_like(src, float("nan"))
    lib.dropout_bf16(src, dst, p, seed=1234)
    keep = dst > 0
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3                       # 792 k draws: sigma = 4.5e-4
    assert torch.allclose(dst[keep].float(), torch.tensor(1 / (1 - p), device="cuda"), rtol=4e-3)   # bf16 rounding of 1.25
    assert (dst[~keep] == 0).all()
    dst2 = torch.empty_like(dst)
    lib.dropout_bf16(src, dst2, p, seed=1234)
    assert torch.equal(dst, dst2)                                                 # a pure function of (seed, coordinates)
    lib.dropout_bf16(src, dst2, p, seed=1235)
    assert (dst2 != dst).float().mean().item() > 0.2                              # another seed, another mask
    # rows and groups are decorrelated
    assert abs((keep[0] & keep[1]).float().mean().item() - (1 - p) ** 2) < 5e-3


@pytest.mark.parametrize("M,N,K,g,outs", [(300, 200, 136, (2,), "f32"), (300, 264, 136, (), "both"), (4096, 1024, 512, (2,), "f32"),
                                          (100, 72, 40, (), "legacy")])
def test_gemm_epilogue_dropout(lib, M, N, K, g, outs):
    torch.manual_seed(11)
    dev, p, seed = "cuda", 0.25, 987654321
    A = torch.randn(*g, M, _r8(K) + 8, device=dev).to(torch.bfloat16)[..., :K]
    B = torch.randn(*g, N, _r8(K) + 8, device=dev).to(torch.bfloat16)[..., :K]
    bias = torch.randn(*g, N, device=dev)
    res = torch.randn(*g, M, N + 4, device=dev)[..., :N]
    ldf = _r8(N) + (3 if outs == "legacy" else 4)          # N + 3: pitch not TMA-able -> direct-store path
    of = torch.full((*g, M, ldf), float("nan"), device=dev)[..., :N]
    ob = torch.zeros(*g, M, _r8(N) + 8, device=dev, dtype=torch.bfloat16)[..., :N] if outs == "both" else None
    lib.gemm(A, B, bias=bias, act=lib.ACT_RELU, residual=res, out_f32=of, out_bf16=ob, dropout_p=p, dropout_seed=seed)
    ones = torch.ones(max(1, len(g) and g[0]), M, _r8(N), device=dev, dtype=torch.bfloat16)[..., :N]
    mask = torch.empty_like(ones)
    lib.dropout_bf16(ones, mask, p, seed)                  # = keep / (1 - p), bf16
    torch.cuda.synchronize()
    keep = (mask > 0).float().view(*g, M, N)
    ref = torch.relu(A.float() @ B.float().transpose(-1, -2) + bias[..., None, :]) * keep / (1 - p) + res
    scale = ref.abs().max().item()
    assert (of - ref).abs().max().item() <= 1e-4 * scale + 1e-5
    if ob is not None:
        assert (ob.float() - ref).abs().max().item() <= 8e-3 * scale + 1e-3
    assert abs(keep.mean().item() - (1 - p)) < 2e-2


@pytest.mark.parametrize("G,B,T,H,hs,S", [(2, 2, 256, 4, 64, 1), (1, 2, 256, 2, 64, 3), (1, 3, 100, 2, 16, 2), (2, 3, 64, 8, 32, 1)])
def test_attention_dropout_fwd_bwd(lib, G, B, T, H, hs, S):
    torch.manual_seed(12)
    p, seed = 0.2, 55555
    C = H * hs
    q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    kv = [torch.randn(G, B * T, H, 2 * hs, device="cuda").to(torch.bfloat16) for _ in range(S)]
    d_o = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    o = torch.zeros(G, B * T, C, device="cuda", dtype=torch.bfloat16)
    o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(S, G, B, H, T, device="cuda")
    k_views = [t.view(G, B * T, 2 * C) for t in kv]
    v_views = [t.view(G, B * T, 2 * C)[..., hs:] for t in kv]
    lib.attention_fwd(q, k_views, v_views, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src, dropout_p=p,
                      dropout_seed=seed)
    dq = torch.full_like(q, float("nan"))
    dkv = [torch.full((G, B * T, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(S)]
    lib.attention_bwd(q, k_views, v_views, o, d_o, lse, dq, dkv, [t[..., hs:] for t in dkv], B=B, T=T, H=H, hs=hs,
                      head_stride_kv=2 * hs, o_src=o_src if S > 1 else None, dropout_p=p, dropout_seed=seed)
    mask = torch.empty(S, G, B, H, T, T, device="cuda", dtype=torch.uint8)
    lib.dropout_attn_mask(mask, S, G, B, H, T, p, seed)
    torch.cuda.synchronize()
    tril = torch.ones(T, T, dtype=torch.bool, device="cuda").tril()
    assert abs(mask[..., tril].float().mean().item() - (1 - p)) < 1e-2
    # reference with the same mask: softmax, THEN inverted dropout on the probabilities, THEN P V (model.py:67-72)
    qf = q.float().requires_grad_(True)
    kvf = [t.float().requires_grad_(True) for t in kv]
    qh = qf.view(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
    outs = []
    for s in range(S):
        kh = kvf[s][..., :hs].reshape(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
        vh = kvf[s][..., hs:].reshape(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
        sc = (qh @ kh.transpose(-1, -2)) * hs ** -0.5
        pr = torch.softmax(sc.masked_fill(~tril, float("-inf")), -1) * mask[s].float() / (1 - p)
        outs.append((pr @ vh).permute(0, 1, 3, 2, 4).reshape(G, B * T, C))
    ro = sum(outs)
    ro.backward(d_o.float())
    assert (o.float() - ro).abs().max().item() <= 2e-2 * ro.abs().max().item()
    tol = 2.5e-2
    assert (dq.float() - qf.grad).abs().max().item() <= tol * qf.grad.abs().max().item()
    for s in range(S):
        ref = kvf[s].grad.view(G, B * T, 2 * C)
        assert (dkv[s].float() - ref).abs().max().item() <= tol * ref.abs().max().item()


@pytest.mark.parametrize("G,B,H,hs,cap,length,S", [(2, 3, 4, 64, 40, 33, 1), (1, 5, 2, 32, 64, 64, 3), (3, 2, 3, 16, 10, 1, 2),
                                                   (1, 4, 12, 64, 1024, 1000, 3), (2, 2, 4, 8, 7, 5, 1),
                                                   (1, 2, 2, 48, 20, 17, 2), (2, 2, 3, 24, 9, 9, 1), (1, 3, 2, 56, 70, 70, 1), (1, 2, 4, 40, 33, 32, 3)])
def test_attention_decode(lib, G, B, H, hs, cap, length, S):
    """ta_attention_decode (one new query row against a K/V cache, generate's growing phase) vs fp32 torch: per source an
    independent softmax over the cached positions [0, length), outputs summed (model.py:60-73, 136-159, last row)."""
    g = torch.Generator(device="cuda").manual_seed(17)
    C = H * hs
    q = torch.randn(G, B, 3 * C, device="cuda", generator=g).to(torch.bfloat16)              # packed q|k|v row, q first
    caches = [torch.randn(G, B, cap, 2 * C, device="cuda", generator=g).to(torch.bfloat16) for _ in range(S)]   # per head [k|v]
    o = torch.zeros(G, B, C, device="cuda", dtype=torch.bfloat16)
    ks = [c for c in caches]
    vs = [c[..., hs:] for c in caches]
    lib.attention_decode(q[..., :C], ks, vs, o, length=length, H=H, hs=hs, head_stride_kv=2 * hs)
    qh = q[..., :C].float().view(G, B, H, hs)
    ref = 0
    for c in caches:
        kv = c[:, :, :length].float().view(G, B, length, H, 2 * hs)
        sc = torch.einsum("gbhd,gbthd->gbht", qh, kv[..., :hs]) * hs ** -0.5
        ref = ref + torch.einsum("gbht,gbthd->gbhd", torch.softmax(sc, -1), kv[..., hs:])
    ref = ref.reshape(G, B, C)
    assert (o.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    assert ((o.float() - ref).norm() / ref.norm()).item() <= 5e-3      # bf16 rounding of the output only


def test_directional_metrics_kernel_matches_reference_golden(lib):
    """ta_directional_metrics against the reference's own calculate_evaluation_metrics (tests/golden/metrics.pt, made by
    /root/reference/training_utils.py:215-330): value data, percentage data, a non-numeric vocabulary; plus accumulation
    over several calls and a batch larger than one shared-memory pass."""
    from conftest import load_golden
    from tradeaid_b200 import training_utils as tu
    g = load_golden("metrics.pt")
    cu = lambda ts: [t.cuda() for t in ts]  # noqa: E731
    w, l, c, p = tu.calculate_evaluation_metrics(cu(g["logits"]), cu(g["xb"]), cu(g["yb"]), 3, g["vocabs"], g["params"], None)
    rw, rl, rc, rp = g["result"]
    assert w == rw and l == rl and p == rp
    for a, b in zip(c, rc):
        assert abs(a - b) <= 1e-5
    # large batch (3 passes of 1024 samples) against the torch formulation, accumulated twice
    gen = torch.Generator().manual_seed(4)
    B, V = 2500, 333
    logits = torch.randn(B, 3, V, generator=gen)
    vals = torch.randn(V, generator=gen, dtype=torch.float64).round(decimals=1)       # ties in value -> zero directions
    xb, yb = torch.randint(0, V, (B, 5), generator=gen), torch.randint(0, V, (B, 5), generator=gen)
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    cert = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(2):
        lib.directional_metrics(logits.cuda()[:, -1, :], vals.cuda(), xb.cuda()[:, -1], yb.cuda()[:, -1], False, counts, cert)
    last = logits[:, -1, :]
    ref = vals[xb[:, -1]]
    ps, as_ = torch.sign(vals[last.argmax(-1)] - ref), torch.sign(vals[yb[:, -1]] - ref)
    mass = (torch.softmax(last, -1).double() * (torch.sign(vals[None, :] - ref[:, None]) == ps[:, None])).sum()
    assert counts.tolist() == [2 * int((ps == as_).sum()), 2 * int((ps != as_).sum())]
    assert abs(float(cert) - 2 * float(mass)) <= 1e-6 * B


def test_augment_tokens_matches_reference_golden():
    """ta_augment_tokens through TokenStore.augment_modality on the GPU against the reference's own outputs
    (tests/golden/augmentation.pt: values after 1 and 3 cumulative applications, position of Python's `random` stream),
    plus a multi-tile case (several 2048-item tiles, rejection sampling for rand_size 2) against the oracle loop."""
    import random
    from conftest import load_golden
    from oracle import restatement as R
    from tradeaid_b200.feeder import TokenStore
    g = load_golden("augmentation.pt")
    for c in g["cases"]:
        for rounds, key_out, key_tail in ((1, "after1", "tail1"), (3, "after3", "tail3")):
            store = TokenStore([c["data"], c["data"]], "cuda")
            random.seed(c["seed"])
            for _ in range(rounds):
                store.augment_modality(1, c["size"], c["V"])
            assert store.store[1].tolist() == c[key_out], (c["name"], rounds)
            assert store.store[0].tolist() == c["data"]
            assert [random.getrandbits(32) for _ in range(4)] == c[key_tail]
    gen = torch.Generator().manual_seed(6)
    for size, V, N in ((2, 64, 100_003), (1, 1024, 50_000), (3, 9, 20_000)):
        data = torch.randint(0, V, (N,), generator=gen).tolist()
        store = TokenStore([data], "cuda")
        random.seed(77)
        store.augment_modality(0, size, V)
        store.augment_modality(0, size, V)
        tail = random.getrandbits(32)
        random.seed(77)
        ref = R.add_rand_to_data_points(R.add_rand_to_data_points(list(data), size, V), size, V)
        assert store.store[0].tolist() == ref and tail == random.getrandbits(32)


_WIDE_SCRIPT = r"""
import sys, torch
from tradeaid_b200 import lib
torch.manual_seed(0)
outs = []
for (M, N, K, G, act) in [(300, 192, 128, 3, 0), (1024, 256, 128, 8, 0), (640, 512, 1024, 2, lib.ACT_RELU), (512, 128, 64, 1, lib.ACT_TANH),
                          (2048, 1024, 512, 2, 0), (700, 64, 96, 2, 0)]:
    A = torch.randn(G, M, K + 8, device="cuda").to(torch.bfloat16)[..., :K]
    B = torch.randn(G, N, K + 8, device="cuda").to(torch.bfloat16)[..., :K]
    bias = torch.randn(G, N, device="cuda") if act else None
    ob = torch.full((G, M, N + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    lib.gemm(A, B, out_bf16=ob[..., :N], alpha=0.05, **({"bias": bias, "act": act} if act else {}))
    ref = (A.float() @ B.float().transpose(-1, -2)) * 0.05
    if act:
        ref = ref + bias[:, None, :]
        ref = torch.relu(ref) if act == lib.ACT_RELU else torch.tanh(ref)
    assert (ob[..., :N].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3, (M, N, K, G, act)
    assert (ob[..., N:] == 7.0).all(), "wrote past the last column"
    outs.append(ob.cpu())
# the saved-activation epilogues (dX = (dY W) * tanh'(a) / ReLU'(a)): the activation is TMA-loaded into the staging buffer
for (M, N, K, G, mode) in [(300, 192, 256, 3, lib.AUX_TANH_BWD), (1024, 128, 256, 8, lib.AUX_TANH_BWD), (900, 1024, 512, 2, lib.AUX_RELU_BWD),
                           (512, 64, 64, 2, lib.AUX_RELU_BWD)]:
    A = torch.randn(G, M, K + 8, device="cuda").to(torch.bfloat16)[..., :K]
    B = torch.randn(G, K, N + 8, device="cuda").to(torch.bfloat16)[..., :N]
    aux = torch.randn(G, M, N + 8, device="cuda").to(torch.bfloat16)[..., :N]
    ob = torch.full((G, M, N + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    lib.gemm(A, B, b_major=1, aux=aux, aux_mode=mode, out_bf16=ob[..., :N], alpha=0.05)
    ref = (A.float() @ B.float()) * 0.05
    ref = ref * (1 - aux.float() ** 2) if mode == lib.AUX_TANH_BWD else torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    assert (ob[..., :N].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3, (M, N, K, G, mode)
    assert (ob[..., N:] == 7.0).all(), "wrote past the last column"
    outs.append(ob.cpu())
torch.save(outs, sys.argv[1])
"""


def test_gemm_wide_bf16_store_boxes_are_bit_identical(tmp_path):
    """The bf16-only epilogue with [32 x 64] store boxes (TA_GEMM_WIDE=2: wherever N % 64 == 0) against 32-column boxes
    (TA_GEMM_WIDE=0): same accumulators, same rounding, a different way out - every output bit equal, nothing written past
    N or M (ragged M, an N tail tile, 1-CTA 128/256-wide and CTA-pair tiles, bias + activation and tanh' / ReLU' epilogues). The switch is read
    once per process, hence the subprocesses."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for mode in ("0", "2"):
        out = tmp_path / f"wide_{mode}.pt"
        env = dict(os.environ, TA_GEMM_WIDE=mode, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        r = subprocess.run([sys.executable, "-c", _WIDE_SCRIPT, str(out)], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(torch.load(out))
    for a, b in zip(*res):
        assert torch.equal(a, b)
