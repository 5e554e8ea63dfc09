import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tradeaid_b200.engine import Engine, KVCache
from tradeaid_b200.params import Dims
from oracle import restatement as R
M, cross, V, B, T, C, H, L, n0 = 4, [True, False, False, True], [64]*4, 2, 300, 768, 12, 1, 290
for scale in (1.0, 2.0, 3.0):
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
    eng.layout.init_(eng.master, generator=torch.Generator(device="cuda").manual_seed(2))
    eng.master.mul_(scale)
    g = torch.Generator().manual_seed(3)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    sd = {k: v.cpu().clone() for k, v in eng.layout.named_views(eng.master).items()}
    ref, _ = R.forward(sd, R.Cfg(T, C, H, L, V, cross), list(idx.cpu()), None)
    full, _ = eng.forward(idx, None, save=False)
    full = [x.clone().cpu() for x in full]
    cache = KVCache(eng, B)
    eng.forward(idx[:, :, :n0].contiguous(), None, save=False, cache=cache)
    out = []
    for t in range(n0, T):
        lg, _ = eng.forward(idx[:, :, t:t + 1].contiguous(), None, save=False, cache=cache)
        lg = [x.cpu() for x in lg]
        e_df = max(((a[:, 0] - b[:, t]).norm() / b[:, t].norm()).item() for a, b in zip(lg, full))
        e_dr = max(((a[:, 0] - b[:, t]).norm() / b[:, t].norm()).item() for a, b in zip(lg, ref))
        e_fr = max(((a[:, t] - b[:, t]).norm() / b[:, t].norm()).item() for a, b in zip(full, ref))
        out.append((t, round(e_df, 4), round(e_dr, 4), round(e_fr, 4)))
    print("scale", scale, "(t, decode-vs-full, decode-vs-oracle, full-vs-oracle)", out)
