"""Dev tool: time the GEMM shapes of the c4 training step (same epilogues as engine.py) in one process.
usage: gemm_c4.py [tag]   -> one JSON line per shape"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib

CASES = [
    # name, M, N, K, G, am, bm, flags
    ("ffn1_fwd", 16384, 4096, 1024, 8, 0, 0, "bias,relu,bf16"),
    ("ffn2_dx", 16384, 4096, 1024, 8, 0, 1, "aux,bf16"),
    ("ffn2_fwd", 16384, 1024, 4096, 8, 0, 0, "bias,res,f32,bf16"),
    ("ffn1_dx", 16384, 1024, 4096, 8, 0, 1, "bf16"),
    ("ffn_dw", 4096, 1024, 16384, 8, 1, 1, "f32"),
    ("qkv1_fwd", 16384, 1536, 1024, 8, 0, 0, "bias,tanh,bf16"),
    ("qkv1_dx", 16384, 1024, 1536, 8, 0, 1, "bf16"),
    ("proj2_fwd", 16384, 1024, 512, 8, 0, 0, "bias,res,f32"),
    ("proj1_fwd", 16384, 512, 1024, 8, 0, 0, "bias,tanh,bf16"),
    ("proj2_dx", 16384, 512, 1024, 8, 0, 1, "auxt,bf16"),
    ("proj1_dx", 16384, 1024, 512, 8, 0, 1, "bf16"),
    ("ckv_fwd", 16384, 8192, 1024, 1, 0, 0, "bf16"),
    ("ckv_dx", 16384, 1024, 8192, 1, 0, 1, "res,f32,bf16"),
    ("ckv_dw", 8192, 1024, 16384, 1, 1, 1, "f32"),
    ("cq_fwd", 16384, 1024, 1024, 4, 0, 0, "bf16"),
    ("qkv2_fwd", 16384, 64, 32, 384, 0, 0, "bf16"),
    ("qkv2_dx", 16384, 32, 64, 384, 0, 1, "auxt,bf16"),
    ("qkv2_dw", 64, 32, 16384, 384, 1, 1, "f32"),
]
tag = sys.argv[1] if len(sys.argv) > 1 else ""
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
dev = "cuda"
for name, M, N, K, G, am, bm, fl in CASES:
    if only and name not in only:
        continue
    flags = set(fl.split(","))
    A = torch.randn(G, M, K, device=dev).to(torch.bfloat16) if am == 0 else torch.randn(G, K, M, device=dev).to(torch.bfloat16)
    B = torch.randn(G, N, K, device=dev).to(torch.bfloat16) if bm == 0 else torch.randn(G, K, N, device=dev).to(torch.bfloat16)
    kw = {}
    if "bias" in flags: kw["bias"] = torch.randn(G, N, device=dev)
    if "relu" in flags: kw["act"] = lib.ACT_RELU
    if "tanh" in flags: kw["act"] = lib.ACT_TANH
    if "aux" in flags: kw.update(aux=torch.randn(G, M, N, device=dev).to(torch.bfloat16), aux_mode=lib.AUX_RELU_BWD)
    if "auxt" in flags: kw.update(aux=torch.randn(G, M, N, device=dev).to(torch.bfloat16), aux_mode=lib.AUX_TANH_BWD)
    if "res" in flags: kw["residual"] = torch.randn(G, M, N, device=dev)
    if "f32" in flags: kw["out_f32"] = torch.empty(G, M, N, device=dev)
    if "bf16" in flags: kw["out_bf16"] = torch.empty(G, M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2): lib.gemm(A, B, a_major=am, b_major=bm, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): lib.gemm(A, B, a_major=am, b_major=bm, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    byt = G * (M * K * 2 + N * K * 2 + M * N * (2 * ("bf16" in flags) + 4 * ("f32" in flags) + 4 * ("res" in flags) + 2 * ("aux" in flags or "auxt" in flags)))
    print(json.dumps(dict(tag=tag, name=name, M=M, N=N, K=K, G=G, flags=fl, ms=round(ms, 4),
                          tflops=round(2.0 * M * N * K * G / ms / 1e9, 1), gbs=round(byt / ms / 1e6, 1))), flush=True)
    del A, B, kw
    torch.cuda.empty_cache()
