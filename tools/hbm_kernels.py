"""The HBM-bound kernels of the path at c4 sizes (and the decode attention at c5 sizes), each timed alone with CUDA
events: achieved GB/s = ALGORITHMIC bytes / launch time, next to the measured copy bandwidth of
MEASURED_PEAKS.json. Also the ncu target for these kernels:  HBM_ONCE=1 runs every kernel once so that
`ncu --set full -c 40` captures one launch of each.   usage: hbm_kernels.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tradeaid_b200 import lib
dev = "cuda"
ONCE = os.environ.get("HBM_ONCE", "0") == "1"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2: written between timed launches


def timeit(fn, n=10):
    if ONCE:
        fn(); torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


out = []
def report(kernel, what, nbytes, ms):
    gbs = nbytes / ms / 1e6 if ms == ms else None
    out.append(dict(kernel=kernel, shape=what, algorithmic_bytes=nbytes, ms=None if ms != ms else round(ms, 4),
                    achieved_GBs=None if gbs is None else round(gbs, 1), peak_GBs=peak, frac=None if gbs is None else round(gbs / peak, 3)))
    print(json.dumps(out[-1]), flush=True)


M, B, T, C, V, N = 8, 64, 256, 1024, 1024, 45_000_000
rows = B * T
g = torch.Generator(device=dev).manual_seed(0)
# K7 get_batch gather
store = torch.randint(0, V, (M, N), device=dev, generator=g, dtype=torch.int32)
ix = torch.randint(0, N - T - 1, (B,), device=dev, generator=g)
xb = torch.empty(M, B, T, dtype=torch.int64, device=dev); yb = torch.empty_like(xb)
report("gather_batch_kernel", f"M={M} B={B} T={T}", M * B * (T + 1) * 4 + 2 * M * B * T * 8, timeit(lambda: lib.gather_batch(store, ix, T, xb, yb)))
del store
# K1 embedding
tok = torch.randn(M * V, C, device=dev); pos = torch.randn(T, C, device=dev)
voff = torch.arange(0, (M + 1) * V, V, device=dev, dtype=torch.int64)
x = torch.empty(M, rows, C, device=dev)
lib.gather_batch(torch.randint(0, V, (M, 100000), device=dev, generator=g, dtype=torch.int32), ix % 90000, T, xb, yb)
report("embed_fwd_kernel", f"M={M} rows={rows} C={C}", M * rows * (8 + 4 * C + 4 * C), timeit(lambda: lib.embed_fwd(xb.view(M, rows), tok, voff, pos, x, T)))
dtok = torch.empty(M * V, C, device=dev); dpos = torch.empty(T, C, device=dev)
report("embed_bwd (tok + pos kernels)", f"M={M} rows={rows} C={C}", M * rows * C * 4 * 2 + (M * V + T) * C * 4,
       timeit(lambda: lib.embed_bwd(xb.view(M, rows), x, voff, V, dtok, dpos, B, T)))
# K4 LayerNorm
gamma = torch.randn(M, C, device=dev); beta = torch.randn(M, C, device=dev)
y = torch.empty(M, rows, C, device=dev, dtype=torch.bfloat16); mean = torch.empty(M * rows, device=dev); rstd = torch.empty(M * rows, device=dev)
report("ln_fwd_kernel", f"G={M} rows={rows} C={C}", M * rows * C * 6, timeit(lambda: lib.layernorm_fwd(x, gamma, beta, y, mean, rstd)))
dy = torch.randn(M, rows, C, device=dev).to(torch.bfloat16); dx = torch.randn(M, rows, C, device=dev); dxb = torch.empty_like(dy)
dg = torch.empty(M, C, device=dev); db = torch.empty(M, C, device=dev)
ws = torch.empty(lib.layernorm_bwd_workspace(M, rows, C), device=dev)
report("ln_bwd_kernel (+2 partial reduces)", f"G={M} rows={rows} C={C}", M * rows * C * 16, timeit(lambda: lib.layernorm_bwd(dy, x, gamma, mean, rstd, dx, dx, dxb, dg, db, ws)))
# K5 cross-entropy (fwd + dlogits), one modality
logits = torch.randn(rows, V, device=dev); tg = torch.randint(0, V, (rows,), device=dev, generator=g)
rl = torch.empty(rows, device=dev); lm = torch.empty(1, device=dev); dl = torch.empty(rows, V, device=dev, dtype=torch.bfloat16)
report("ce_fwd_bwd_kernel (+row mean)", f"rows={rows} V={V}", rows * V * 6 + rows * 12, timeit(lambda: lib.cross_entropy(logits, tg, rl, lm, dl, 1.0)))
# bias gradients
dyw = torch.randn(M, rows, 4 * C, device=dev).to(torch.bfloat16); cs = torch.empty(M, 4 * C, device=dev)
cw = torch.empty(lib.colsum_workspace(rows, 4 * C, M), device=dev)
report("colsum_partial_kernel (+reduce)", f"G={M} rows={rows} N={4 * C}", M * rows * 4 * C * 2, timeit(lambda: lib.colsum(dyw, cs, cw)))
del dyw, x, dx, dy, dxb, y
# K6 AdamW: c4 has 1.89e9 parameters (30 B each); HBM_ONCE (ncu replays with save/restore) uses an eighth
n = 236_000_000 if ONCE else 1_890_000_000
p = torch.randn(n, device=dev); gr = torch.randn(n, device=dev); m1 = torch.zeros(n, device=dev); v1 = torch.zeros(n, device=dev)
sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
step = torch.zeros(1, dtype=torch.int64, device=dev); sc = torch.zeros(4, device=dev)
lib.adamw_tick(step, sc, 3e-4)
report("adamw_kernel", f"n={n}", n * 30, timeit(lambda: lib.adamw_step(p, gr, m1, v1, sh, sc), n=5))
del p, gr, m1, v1, sh
# K3 attention at the c4 shapes: self (8 modalities in one launch) and cross (one querier, 7 sources). Algorithmic bytes:
# forward q, o + per source k, v (+ the per-source outputs when there is more than one source); backward q, dO, dq + per
# source k, v, dk, dv, o_s (the D table pass reads dO and o_s; its 4-byte-per-row output and the lse tables are left out)
def attn_case(Ga, S):
    Ha, hsa, Ta, Ba = 16, 64, 256, 64
    Ca = Ha * hsa
    t = Ga * Ba * Ta * Ca * 2
    qa = torch.randn(Ga, Ba * Ta, Ca, device=dev).to(torch.bfloat16)
    kva = [torch.randn(Ga, Ba * Ta, 2 * Ca, device=dev).to(torch.bfloat16) for _ in range(S)]
    doa = torch.randn(Ga, Ba * Ta, Ca, device=dev).to(torch.bfloat16)
    oa = torch.zeros_like(qa); osa = torch.zeros(S, Ga, Ba * Ta, Ca, device=dev, dtype=torch.bfloat16) if S > 1 else None
    lsea = torch.zeros(S, Ga, Ba, Ha, Ta, device=dev)
    dqa = torch.zeros_like(qa); dkva = [torch.zeros_like(x) for x in kva]
    ksa, vsa = kva, [x[..., hsa:] for x in kva]
    kw = dict(B=Ba, T=Ta, H=Ha, hs=hsa, head_stride_kv=2 * hsa, o_src=osa)
    what = f"G={Ga} B={Ba} T={Ta} H={Ha} hs={hsa} sources={S}"
    report("attn_fwd2_kernel", what, t * (2 + 2 * S + (S if S > 1 else 0)), timeit(lambda: lib.attention_fwd(qa, ksa, vsa, oa, lsea, **kw)))
    report("attn_bwd_dtab_kernel + attn_bwd_kernel", what, t * (3 + 5 * S),
           timeit(lambda: lib.attention_bwd(qa, ksa, vsa, oa, doa, lsea, dqa, dkva, [x[..., hsa:] for x in dkva], **kw)))
attn_case(8, 1)
attn_case(1, 7)
# K3d decode attention at c5 sizes: 4 modalities, batch 256, 12 heads x 64, 768 cached positions
G, Bc, H, hs, cap, ln = 4, 256, 12, 64, 1024, 768
Cc = H * hs
q = torch.randn(G, Bc, 3 * Cc, device=dev).to(torch.bfloat16)
kv = torch.randn(G, Bc, cap, 2 * Cc, device=dev).to(torch.bfloat16)
o = torch.empty(G, Bc, Cc, device=dev, dtype=torch.bfloat16)
report("attn_decode_kernel<64>", f"G={G} B={Bc} H={H} len={ln}", G * Bc * ln * 2 * Cc * 2,
       timeit(lambda: lib.attention_decode(q[..., :Cc], [kv[..., :Cc]], [kv[..., Cc:]], o, length=ln, H=H, hs=hs)))
del q, kv, o
# f-2: the [2]-quirk augmentation of one modality of the resident token store (c4: 45 M training tokens, rand_size 1:
# 3 candidates -> 2 bits per draw, 4/3 words per accepted draw + slack). Added at the end of round 1: first measured in round 2.
Nt = 45_000_000 if not ONCE else 8_000_000
tokens = torch.randint(0, V, (Nt,), device=dev, generator=g, dtype=torch.int32)
Wn = int(Nt * (4 / 3) * 1.03) + 4096
words = torch.randint(-2 ** 31, 2 ** 31 - 1, (Wn,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
aws = torch.empty(lib.augment_workspace_bytes(Nt, Wn), dtype=torch.uint8, device=dev)
astats = torch.zeros(3, dtype=torch.int64, device=dev)
report("aug_count/scan/compact/apply (ta_augment_tokens)", f"N={Nt} W={Wn}", 2 * (4 * Wn + 4 * Nt) + 4 * Nt + Wn,
       timeit(lambda: lib.augment_tokens(tokens, words, 1, V, aws, astats), n=5))
del tokens, words, aws
# f-1: directional metrics of one evaluation batch (B = 64 last-position rows x V logits): launch-latency sized
Bm = 64
lg = torch.randn(Bm, V, device=dev); vv = torch.linspace(-5, 5, V, dtype=torch.float64, device=dev)
pt = torch.randint(0, V, (Bm,), device=dev, generator=g); at = torch.randint(0, V, (Bm,), device=dev, generator=g)
cnt = torch.zeros(2, dtype=torch.int64, device=dev); cert = torch.zeros(1, dtype=torch.float64, device=dev)
report("directional_metrics_kernel", f"B={Bm} V={V}", Bm * V * 4 + V * 8 + Bm * 16,
       timeit(lambda: lib.directional_metrics(lg, vv, pt, at, False, cnt, cert)))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
