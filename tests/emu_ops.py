"""CPU emulation of the kernel namespace `tradeaid_b200.lib` — TEST INFRASTRUCTURE ONLY.

Same Python signatures and the same rounding points (bf16 operands / bf16 activation stores, fp32 accumulation and
statistics) as the CUDA kernels, written with plain torch ops on CPU tensors. It lets the CPU test-suite run the
engine's host logic (buffer wiring, grouping, the hand-written backward program, packing) against the oracle without a
GPU. The product never imports this module."""
import math

import torch

ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2
AUX_NONE, AUX_TANH_BWD, AUX_RELU_BWD = 0, 1, 2

_launches = 0


def launch_count():
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _tick(n=1):
    global _launches
    _launches += n


def dropout_bf16(src, dst, p, seed, dropout_step=None):
    raise NotImplementedError("the CPU emulation covers dropout = 0 only (masks come from the CUDA Philox kernels)")


def gemm(A, B, *, a_major=0, b_major=0, out_f32=None, out_bf16=None, bias=None, act=ACT_NONE, aux=None,
         aux_mode=AUX_NONE, residual=None, alpha=1.0):
    _tick()
    assert A.dtype == B.dtype
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and A.stride(-2) % 8 == 0 and B.stride(-2) % 8 == 0
    for t in (A, B):
        for k in range(t.dim() - 2):
            assert t.shape[k] == 1 or t.stride(k) % 8 == 0, "TMA needs 16-byte group strides"
    Af = A.float() if a_major == 0 else A.float().transpose(-1, -2)
    Bf = B.float() if b_major == 0 else B.float().transpose(-1, -2)
    v = (Af @ Bf.transpose(-1, -2)) * alpha
    if bias is not None:
        v = v + bias.unsqueeze(-2)
    if act == ACT_TANH:
        v = torch.tanh(v)
    elif act == ACT_RELU:
        v = torch.relu(v)
    if aux is not None:
        a = aux.float()
        v = v * (1 - a * a) if aux_mode == AUX_TANH_BWD else torch.where(a > 0, v, torch.zeros_like(v))
    if residual is not None:
        v = v + residual
    if out_f32 is not None:
        out_f32.copy_(v)
    if out_bf16 is not None:
        out_bf16.copy_(v.to(out_bf16.dtype))


def _heads(t, B, T, H, hs, head_stride):
    """[G, B*T, *] view with head h at columns [h*head_stride, +hs) -> fp32 [G, B, H, T, hs]."""
    if t.dim() == 2:
        t = t.unsqueeze(0)
    G = t.shape[0]
    v = t.as_strided((G, B, T, H, hs), (t.stride(0), T * t.stride(1), t.stride(1), head_stride, 1), t.storage_offset())
    return v.permute(0, 1, 3, 2, 4)


def _scatter_heads(dst, src, B, T, H, hs, head_stride):
    """inverse of _heads: src [G, B, H, T, hs] written into dst's head columns (bf16 rounding)."""
    if dst.dim() == 2:
        dst = dst.unsqueeze(0)
    G = dst.shape[0]
    v = dst.as_strided((G, B, T, H, hs), (dst.stride(0), T * dst.stride(1), dst.stride(1), head_stride, 1),
                       dst.storage_offset())
    v.copy_(src.permute(0, 1, 3, 2, 4).to(dst.dtype))


def _attn_core(q, k, v, hs):
    T = q.shape[-2]
    s = (q @ k.transpose(-1, -2)) * hs ** -0.5
    mask = torch.ones(T, T, dtype=torch.bool).tril()
    s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse.unsqueeze(-1))
    return p, lse


def attention_fwd(q, ks, vs, o, lse, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None, o_src=None):
    _tick()
    hq = head_stride_q or hs
    hkv = head_stride_kv or hs
    qh = _heads(q, B, T, H, hs, hq).float()
    G = qh.shape[0]
    tot = 0
    lse_v = lse.view(len(ks), G, B, H, T)
    for s, (k, v) in enumerate(zip(ks, vs)):
        kh, vh = _heads(k, B, T, H, hs, hkv).float(), _heads(v, B, T, H, hs, hkv).float()
        p, l = _attn_core(qh, kh, vh, hs)
        # the kernel rounds the un-normalised probabilities to bf16 before P V; emulate the same rounding point
        o_s = (p.to(q.dtype).float() @ vh)
        lse_v[s].copy_(l)
        if o_src is not None:
            _scatter_heads(o_src[s], o_s, B, T, H, hs, hs)
        tot = tot + o_s
    _scatter_heads(o, tot, B, T, H, hs, hs)


def attention_fwd_f32(q, ks, vs, o, lse, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None, o_src=None):
    assert q.dtype == torch.float32 and o.dtype == torch.float32
    attention_fwd(q, ks, vs, o, lse, B=B, T=T, H=H, hs=hs, head_stride_q=head_stride_q, head_stride_kv=head_stride_kv, o_src=o_src)


def embed_fwd_at(idx, tok, voff, pos, x, pos_dev):
    _tick()
    M, rows = idx.shape
    for m in range(M):
        x[m].copy_(tok[voff[m] + idx[m]] + pos[int(pos_dev)])


def kv_append(src, cache, pos_dev):
    _tick()
    cache[:, :, int(pos_dev), :src.shape[-1]].copy_(src)


def attention_decode(q, ks, vs, o, *, length, H, hs, head_stride_q=None, head_stride_kv=None, len_dev=None):
    """One new query row per (group, batch) against cache rows [0, length) of every source (csrc/attention_decode.cu)."""
    _tick()
    if len_dev is not None:
        length = length + int(len_dev)
    hq = head_stride_q or hs
    hkv = head_stride_kv or hs
    G, B = q.shape[0], q.shape[1]
    qh = torch.stack([q[..., h * hq:h * hq + hs] for h in range(H)], 2).float()            # [G, B, H, hs]
    tot = 0
    for k, v in zip(ks, vs):
        kh = torch.stack([k[:, :, :length, h * hkv:h * hkv + hs] for h in range(H)], 2).float()   # [G, B, H, len, hs]
        vh = torch.stack([v[:, :, :length, h * hkv:h * hkv + hs] for h in range(H)], 2).float()
        sc = torch.einsum("gbhd,gbhtd->gbht", qh, kh) * hs ** -0.5
        tot = tot + torch.einsum("gbht,gbhtd->gbhd", torch.softmax(sc, -1), vh)
    o.copy_(tot.reshape(G, B, H * hs).to(o.dtype))


def attention_bwd(q, ks, vs, o, d_o, lse, dq, dks, dvs, *, B, T, H, hs, head_stride_q=None, head_stride_kv=None,
                  o_src=None):
    _tick()
    hq = head_stride_q or hs
    hkv = head_stride_kv or hs
    qh = _heads(q, B, T, H, hs, hq).float()
    doh = _heads(d_o, B, T, H, hs, hs).float()
    G = qh.shape[0]
    lse_v = lse.view(len(ks), G, B, H, T)
    dq_tot = 0
    for s, (k, v) in enumerate(zip(ks, vs)):
        kh, vh = _heads(k, B, T, H, hs, hkv).float(), _heads(v, B, T, H, hs, hkv).float()
        sc = (qh @ kh.transpose(-1, -2)) * hs ** -0.5
        mask = torch.ones(T, T, dtype=torch.bool).tril()
        p = torch.exp(sc - lse_v[s].unsqueeze(-1)).masked_fill(~mask, 0.0)
        osrc = _heads(o_src[s] if o_src is not None else o, B, T, H, hs, hs).float()
        D = (doh * osrc).sum(-1, keepdim=True)
        dp = doh @ vh.transpose(-1, -2)
        ds = (p * (dp - D) * hs ** -0.5).to(q.dtype).float()
        pb = p.to(q.dtype).float()
        _scatter_heads(dvs[s], pb.transpose(-1, -2) @ doh, B, T, H, hs, hkv)
        _scatter_heads(dks[s], ds.transpose(-1, -2) @ qh, B, T, H, hs, hkv)
        dq_tot = dq_tot + ds @ kh
    _scatter_heads(dq, dq_tot, B, T, H, hs, hq)


def gather_batch(store, ix, T, xb, yb):
    _tick()
    ar = torch.arange(T)
    xb.copy_(store.long()[:, ix[:, None] + ar])
    yb.copy_(store.long()[:, ix[:, None] + ar + 1])


def embed_fwd(idx, tok, voff, pos, x, T):
    _tick()
    M, rows = idx.shape
    for m in range(M):
        x[m].copy_(tok[voff[m] + idx[m]] + pos[:T].repeat(rows // T, 1))


def embed_bwd(idx, dx, voff, max_vocab, dtok, dpos, B, T):
    _tick(2)
    M = idx.shape[0]
    dtok.zero_()
    for m in range(M):
        dtok.index_add_(0, voff[m] + idx[m], dx[m])
    dpos.copy_(dx.view(M, B, T, -1).sum((0, 1)))


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps=1e-5):
    _tick()
    G, rows, C = x.shape
    mu = x.mean(-1)
    rs = (x.var(-1, unbiased=False) + eps).rsqrt()
    mean.view(G, rows).copy_(mu)
    rstd.view(G, rows).copy_(rs)
    y.copy_((((x - mu.unsqueeze(-1)) * rs.unsqueeze(-1)) * gamma.unsqueeze(1) + beta.unsqueeze(1)).to(y.dtype))


def layernorm_fwd_f32(x, gamma, beta, y, mean, rstd, eps=1e-5):
    assert y.dtype == torch.float32
    layernorm_fwd(x, gamma, beta, y, mean, rstd, eps)


def split_bf16(src, dst, lo_mask, terms=3, concat_rows=False):
    """csrc/split.cu: slot k of dst = lo (bit k of lo_mask) or hi of src."""
    _tick()
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16
    hi = src.to(torch.bfloat16)
    lo = (src - hi.float()).to(torch.bfloat16)
    R, C = src.shape[-2], src.shape[-1]
    for k in range(terms):
        v = lo if (lo_mask >> k) & 1 else hi
        if concat_rows:
            dst[..., k * R:(k + 1) * R, :].copy_(v)
        else:
            dst[..., k * C:(k + 1) * C].copy_(v)


def convert(src, dst):
    _tick()
    assert src.shape == dst.shape and src.stride(-1) == 1 and dst.stride(-1) == 1 and src.dim() <= 4
    dst.copy_(src.to(dst.dtype))


def layernorm_bwd_workspace(G, rows, C):
    return 1


def layernorm_bwd(dy, x, gamma, mean, rstd, dstream_in, dstream_out, dstream_bf16, dgamma, dbeta, workspace):
    _tick(3)
    G, rows, C = x.shape
    mu, rs = mean.view(G, rows, 1), rstd.view(G, rows, 1)
    xh = (x - mu) * rs
    d = dy.float()
    dgamma.copy_((d * xh).sum(1))
    dbeta.copy_(d.sum(1))
    dg = d * gamma.unsqueeze(1)
    dx = rs * (dg - dg.mean(-1, keepdim=True) - xh * (dg * xh).mean(-1, keepdim=True))
    out = dx if dstream_in is None else dstream_in + dx
    dstream_out.copy_(out)
    if dstream_bf16 is not None:
        dstream_bf16.copy_(out.to(dstream_bf16.dtype))


def cross_entropy(logits, targets, row_loss, loss_mean, dlogits=None, grad_scale=1.0):
    _tick(2)
    lse = torch.logsumexp(logits, -1)
    rl = lse - logits.gather(1, targets[:, None]).squeeze(1)
    row_loss.copy_(rl)
    loss_mean.copy_(rl.mean().reshape(1))
    if dlogits is not None:
        p = torch.exp(logits - lse[:, None])
        p[torch.arange(logits.shape[0]), targets] -= 1.0
        dlogits.copy_((p * (grad_scale / logits.shape[0])).to(dlogits.dtype))


def colsum_workspace(rows, N, G):
    return 1


def colsum(dy, out, workspace):
    _tick(2)
    out.copy_(dy.float().sum(1))


def cast_pad(src, dst):
    _tick()
    G, R, Cs = src.shape
    dst.zero_()
    dst[:, :R, :Cs].copy_(src.to(dst.dtype))


def unpad(src, dst):
    _tick()
    G, R, Cn = dst.shape
    dst.copy_(src[:, :R, :Cn])


def cast_bf16(src, dst):
    _tick()
    dst.copy_(src.to(dst.dtype))


def adamw_tick(step_dev, scalars_dev, lr, beta1=0.9, beta2=0.999, weight_decay=0.01):
    _tick()
    step_dev += 1
    t = int(step_dev.item())
    scalars_dev[0] = lr / (1.0 - beta1 ** t)
    scalars_dev[1] = math.sqrt(1.0 - beta2 ** t)
    scalars_dev[2] = 1.0 - lr * weight_decay


def adamw_step(p, g, m, v, shadow, scalars_dev, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _tick()
    step_size, bc2s, decay = (float(scalars_dev[k]) for k in range(3))
    g = g.float() * grad_scale
    p.mul_(decay)
    m.add_((g - m) * (1.0 - beta1))
    v.mul_(beta2).add_((1.0 - beta2) * g * g)
    p.sub_(step_size * (m / (v.sqrt() / bc2s + eps)))
    if shadow is not None:
        shadow.copy_(p.to(shadow.dtype))


def directional_metrics(last_logits, vocab_values, prev_tokens, actual_tokens, is_percentage, counts, certainty):
    """csrc/metrics.cu on CPU tensors."""
    _tick()
    x = last_logits.float()
    ref = torch.zeros(x.shape[0], dtype=torch.float64) if is_percentage else vocab_values[prev_tokens]
    ps = torch.sign(vocab_values[x.argmax(-1)] - ref)
    as_ = torch.sign(vocab_values[actual_tokens] - ref)
    counts[0] += int((ps == as_).sum())
    counts[1] += int((ps != as_).sum())
    probs = torch.softmax(x, -1).double()
    certainty[0] += float((probs * (torch.sign(vocab_values[None, :] - ref[:, None]) == ps[:, None])).sum())


def augment_workspace_bytes(N, W):
    return 64


def augment_tokens(tokens, words, rand_size, vocab_size, workspace, stats):
    """csrc/augment.cu with torch ops: j-th accepted word <-> j-th eligible token."""
    _tick()
    n = 2 * rand_size + 1
    k = 2 if rand_size == 1 else 3
    w = words.to(torch.int64) & 0xFFFFFFFF
    r = w >> (32 - k)
    acc = (r < n).nonzero().flatten()
    elig = ((tokens > rand_size) & (tokens < vocab_size - rand_size)).nonzero().flatten()
    E, A = elig.numel(), acc.numel()
    stats[0], stats[1], stats[2] = E, A, (int(acc[E - 1]) + 1 if 0 < E <= A else 0)
    if A < E or E == 0:
        return
    delta = torch.tensor([0, 1, -1, 2, -2, 3, -3], dtype=torch.int32)
    tokens[elig] += delta[r[acc[:E]]]
