"""Small-shape launches of every hand-rolled mbarrier / TMEM / TMA kernel of the library, meant to be run under
compute-sanitizer (memcheck, racecheck, synccheck, initcheck): GEMM 1-CTA (64/128/256-wide tiles, both majors, fused epilogues)
and CTA-pair, attention forward (both generations, single and multi source) / backward / decode, LayerNorm fwd/bwd,
cross-entropy, embedding fwd/bwd, gather, AdamW, split/convert, augmentation, metrics. Results are checked loosely against torch
so that a tool-induced slowdown does not hide a wrong answer.  usage: compute-sanitizer --tool X python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tradeaid_b200 import lib  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bf = torch.bfloat16


def gemm_case(M, N, K, G, am, bm, act=0, res=False, f32=True, b16=True):
    A = torch.randn(G, *((M, K) if am == 0 else (K, M)), device=dev).to(bf)
    B = torch.randn(G, *((N, K) if bm == 0 else (K, N)), device=dev).to(bf)
    bias = torch.randn(G, N, device=dev)
    r = torch.randn(G, M, N, device=dev) if res else None
    o32 = torch.zeros(G, M, N, device=dev) if f32 else None
    o16 = torch.zeros(G, M, N, device=dev, dtype=bf) if b16 else None
    lib.gemm(A, B, a_major=am, b_major=bm, out_f32=o32, out_bf16=o16, bias=bias, act=act, residual=r)
    Af = A.float() if am == 0 else A.float().transpose(1, 2)
    Bf = B.float() if bm == 0 else B.float().transpose(1, 2)
    v = Af @ Bf.transpose(1, 2) + bias[:, None]
    v = torch.tanh(v) if act == 1 else torch.relu(v) if act == 2 else v
    if res:
        v = v + r
    out = o32 if f32 else o16.float()
    assert (out - v).abs().max().item() <= 3e-2 * v.abs().max().item() + 1e-2, ("gemm", M, N, K, G, am, bm)


for case in [(128, 64, 64, 1, 0, 0), (200, 136, 96, 2, 0, 0, 1, True), (256, 256, 128, 1, 0, 1, 2), (96, 72, 256, 3, 1, 1),
             (512, 512, 256, 2, 0, 0, 0, True), (512, 256, 512, 1, 1, 1), (384, 300, 64, 1, 0, 0, 1, False, False, True)]:
    gemm_case(*case)
print("gemm ok", flush=True)


def attn_case(G, B, T, H, hs, S, version):
    os.environ["TA_ATTN_FWD"] = str(version)
    C = H * hs
    q = torch.randn(G, B * T, C, device=dev).to(bf)
    kv = [torch.randn(G, B * T, 2 * C, device=dev).to(bf) for _ in range(S)]
    d_o = torch.randn(G, B * T, C, device=dev).to(bf)
    o = torch.zeros_like(q)
    o_src = torch.zeros(S, G, B * T, C, device=dev, dtype=bf)
    lse = torch.zeros(S, G, B, H, T, device=dev)
    ks, vs = kv, [t[..., hs:] for t in kv]
    lib.attention_fwd(q, ks, vs, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src)
    dq = torch.zeros_like(q)
    dkv = [torch.zeros_like(t) for t in kv]
    if T <= 256:
        lib.attention_bwd(q, ks, vs, o, d_o, lse, dq, dkv, [t[..., hs:] for t in dkv], B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs,
                          o_src=o_src if S > 1 else None)
    qh = q.float().view(G, B, T, H, hs).permute(0, 1, 3, 2, 4)
    ref = 0
    mask = torch.ones(T, T, dtype=torch.bool, device=dev).tril()
    for t in kv:
        t5 = t.float().view(G, B, T, H, 2 * hs).permute(0, 1, 3, 2, 4)
        s = (qh @ t5[..., :hs].transpose(-1, -2)) * hs ** -0.5
        ref = ref + torch.softmax(s.masked_fill(~mask, float("-inf")), -1) @ t5[..., hs:]
    ref = ref.permute(0, 1, 3, 2, 4).reshape(G, B * T, C)
    torch.cuda.synchronize()
    assert (o.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item(), ("attention", G, B, T, H, hs, S, version)


# the library reads TA_ATTN_FWD once per process: the generation is chosen by the caller of this script
ver = int(os.environ.get("TA_ATTN_FWD", "2"))
for case in [(1, 2, 100, 2, 16, 1), (2, 1, 256, 2, 64, 1), (1, 2, 256, 2, 64, 3), (1, 1, 40, 3, 32, 2), (1, 1, 384, 1, 64, 1)]:
    attn_case(*case, ver)
print("attention ok (forward generation %d)" % ver, flush=True)

# decode
G, B, H, hs, cap, length = 2, 3, 3, 48, 20, 17
C = H * hs
q = torch.randn(G, B, C, device=dev).to(bf)
cache = torch.randn(G, B, cap, 2 * C, device=dev).to(bf)
o = torch.zeros(G, B, C, device=dev, dtype=bf)
lib.attention_decode(q, [cache], [cache[..., hs:]], o, length=length, H=H, hs=hs, head_stride_kv=2 * hs)
# row ops
G, rows, C = 2, 70, 128
x = torch.randn(G, rows, C, device=dev)
w, b = torch.randn(G, C, device=dev), torch.randn(G, C, device=dev)
y = torch.zeros(G, rows, C, device=dev, dtype=bf)
mean, rstd = torch.zeros(G * rows, device=dev), torch.zeros(G * rows, device=dev)
lib.layernorm_fwd(x, w, b, y, mean, rstd)
dy = torch.randn(G, rows, C, device=dev).to(bf)
dx, dxb = torch.zeros(G, rows, C, device=dev), torch.zeros(G, rows, C, device=dev, dtype=bf)
dg, db = torch.zeros(G, C, device=dev), torch.zeros(G, C, device=dev)
ws = torch.zeros(lib.layernorm_bwd_workspace(G, rows, C), device=dev)
lib.layernorm_bwd(dy, x, w, mean, rstd, None, dx, dxb, dg, db, ws)
logits = torch.randn(50, 57, device=dev)
tg = torch.randint(0, 57, (50,), device=dev)
rl, lm = torch.zeros(50, device=dev), torch.zeros(1, device=dev)
dl = torch.zeros(50, 64, device=dev, dtype=bf)[:, :57]
lib.cross_entropy(logits, tg, rl, lm, dl)
assert abs(float(lm) - float(torch.nn.functional.cross_entropy(logits, tg))) < 1e-4
cs_ws = torch.zeros(lib.colsum_workspace(rows, C, G), device=dev)
out = torch.zeros(G, C, device=dev)
lib.colsum(dy, out, cs_ws)
src = torch.randn(2, 3, 37, 40, device=dev)
d3 = torch.zeros(2, 3, 37, 128, device=dev, dtype=bf)
lib.split_bf16(src, d3[..., :120], lib.SPLIT_A, 3, False)
y2 = torch.zeros(2, 3, 37, 40, device=dev, dtype=bf)
lib.convert(src, y2)
store = torch.randint(0, 50, (3, 500), device=dev, dtype=torch.int32)
ix = torch.randint(0, 400, (6,), device=dev)
xb, yb = torch.zeros(3, 6, 16, dtype=torch.int64, device=dev), torch.zeros(3, 6, 16, dtype=torch.int64, device=dev)
lib.gather_batch(store, ix, 16, xb, yb)
assert torch.equal(xb[1, 2], store[1, int(ix[2]):int(ix[2]) + 16].long())
tok, pos = torch.randn(150, 64, device=dev), torch.randn(16, 64, device=dev)
voff = torch.tensor([0, 50, 100, 150], device=dev)
xe = torch.zeros(3, 96, 64, device=dev)
lib.embed_fwd(xb.view(3, 96), tok, voff, pos, xe, 16)
dtok, dpos = torch.zeros_like(tok), torch.zeros_like(pos)
lib.embed_bwd(xb.view(3, 96), torch.randn(3, 96, 64, device=dev), voff, 50, dtok, dpos, 6, 16)
n = 4096
pp, gg, mm, vv = (torch.randn(n, device=dev) for _ in range(4))
vv.abs_()
sh = torch.zeros(n, device=dev, dtype=bf)
step, sc = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(4, device=dev)
lib.adamw_tick(step, sc, 1e-3)
lib.adamw_step(pp, gg, mm, vv, sh, sc)
torch.cuda.synchronize()
print("row ops ok", flush=True)
