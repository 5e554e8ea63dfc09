"""Dev tool: time one GEMM shape with selectable epilogue pieces. usage: gemm_time.py M N K G am bm flags(csv: bias,relu,tanh,aux,res,f32,bf16)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib
M, N, K, G, am, bm = map(int, sys.argv[1:7]); flags = set(sys.argv[7].split(",")) if len(sys.argv) > 7 else {"bf16"}
dev = "cuda"
A = torch.randn(G, M, K, device=dev).to(torch.bfloat16) if am == 0 else torch.randn(G, K, M, device=dev).to(torch.bfloat16)
B = torch.randn(G, N, K, device=dev).to(torch.bfloat16) if bm == 0 else torch.randn(G, K, N, device=dev).to(torch.bfloat16)
kw = {}
if "bias" in flags: kw["bias"] = torch.randn(G, N, device=dev)
if "relu" in flags: kw["act"] = lib.ACT_RELU
if "tanh" in flags: kw["act"] = lib.ACT_TANH
if "aux" in flags: kw.update(aux=torch.randn(G, M, N, device=dev).to(torch.bfloat16), aux_mode=lib.AUX_RELU_BWD)
if "res" in flags: kw["residual"] = torch.randn(G, M, N, device=dev)
if "f32" in flags: kw["out_f32"] = torch.empty(G, M, N, device=dev)
if "bf16" in flags: kw["out_bf16"] = torch.empty(G, M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3): lib.gemm(A, B, a_major=am, b_major=bm, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib.gemm(A, B, a_major=am, b_major=bm, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps(dict(shape=sys.argv[1:8], dbg=os.environ.get("TA_GEMM_DEBUG", "0"), ms=round(ms, 4), tflops=round(2.0 * M * N * K * G / ms / 1e9, 1))))
