"""Dev tool: run ONE GEMM case on the GPU against torch fp32 matmul of the same bf16 operands.
usage: gemm_check.py M N K G0 G1 a_major b_major [epi]   (epi: none|bias_tanh|bias_relu|res|tanh_bwd|relu_bwd)"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib

def main():
    M, N, K, G0, G1, am, bm = map(int, sys.argv[1:8])
    epi = sys.argv[8] if len(sys.argv) > 8 else "none"
    torch.manual_seed(0)
    dev = "cuda"
    g = (G1, G0)
    def r8(x): return (x + 7) // 8 * 8
    # operands with padded pitches (exercise ld != cols)
    def mk(rows, cols):
        buf = torch.randn(*g, rows, r8(cols) + 8, device=dev).to(torch.bfloat16)
        return buf[..., :cols]
    A = mk(M, K) if am == 0 else mk(K, M)
    B = mk(N, K) if bm == 0 else mk(K, N)
    Af = A.float() if am == 0 else A.float().transpose(-1, -2)
    Bf = B.float() if bm == 0 else B.float().transpose(-1, -2)
    ref = Af @ Bf.transpose(-1, -2)
    kw = {}
    if epi in ("bias_tanh", "bias_relu"):
        bias = torch.randn(*g, N, device=dev)
        kw.update(bias=bias, act=lib.ACT_TANH if epi == "bias_tanh" else lib.ACT_RELU)
        ref = ref * 0.05 + bias[..., None, :]
        ref = torch.tanh(ref) if epi == "bias_tanh" else torch.relu(ref)
        kw["alpha"] = 0.05
    elif epi == "res":
        res = torch.randn(*g, M, N, device=dev)
        bias = torch.randn(*g, N, device=dev)
        kw.update(residual=res, bias=bias)
        ref = ref + bias[..., None, :] + res
    elif epi in ("tanh_bwd", "relu_bwd"):
        aux = torch.randn(*g, M, r8(N), device=dev).to(torch.bfloat16)[..., :N]
        kw.update(aux=aux, aux_mode=lib.AUX_TANH_BWD if epi == "tanh_bwd" else lib.AUX_RELU_BWD)
        ref = ref * (1 - aux.float() ** 2) if epi == "tanh_bwd" else torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    of = torch.full((*g, M, N + 3), float("nan"), device=dev)[..., :N]
    ob = torch.zeros(*g, M, r8(N), device=dev, dtype=torch.bfloat16)[..., :N]
    lib.gemm(A, B, a_major=am, b_major=bm, out_f32=of, out_bf16=ob, **kw)
    torch.cuda.synchronize()
    err = (of - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    errb = (ob.float() - ref).abs().max().item()
    ok = err <= 2e-3 * scale + 1e-4 and errb <= 1e-2 * scale + 1e-3
    # timing
    t = None
    if ok and M * N * K * G0 * G1 >= 1 << 24:
        for _ in range(3):
            lib.gemm(A, B, a_major=am, b_major=bm, out_bf16=ob, **{k: v for k, v in kw.items()})
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.gemm(A, B, a_major=am, b_major=bm, out_bf16=ob, **{k: v for k, v in kw.items()})
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10
    out = dict(case=sys.argv[1:], ok=bool(ok), max_err=err, max_err_bf16=errb, ref_scale=scale, ms=t,
               tflops=(2.0 * M * N * K * G0 * G1 / (t * 1e-3) / 1e12) if t else None)
    print(json.dumps(out))
    sys.exit(0 if ok else 1)

main()
