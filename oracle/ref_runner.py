"""Driver of the UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) through its own public API and stock code
path — test / bench infrastructure, never imported by the product:

    model.MultimodalTransformer(num_modalities, vocab_sizes, all_modality_params).to(device)      (main.py:463)
    training_utils.get_batch('train', 1) with the module globals main.py injects                   (main.py:387-396, 641)
    logits, losses = m(xb, yb); sum(losses).backward(); torch.optim.AdamW(m.parameters(), lr).step() (main.py:642-650)

on the synthetic streams of tradeaid_b200.synthetic (the same data the CUDA arm trains on). `device='cpu'` is the reference's
CPU path (bench.py --impl reference, cpu_baseline.kind = "reference"); `device='cuda'` is the reference-eager-on-B200
comparator of BASELINE.md section 3 (fp32, TF32 matmuls, or torch.autocast(bf16)) — the only "Blackwell implementation" the
reference has. None of this repository's model, kernels or engine is on that path."""
import importlib
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "model.py")) and os.path.isfile(os.path.join(REF, "training_utils.py"))


class ReferenceRun:
    def __init__(self, cfg, batch, device, rows, lr=3e-4, seed=0, mode="fp32"):
        from tradeaid_b200 import synthetic as syn
        if not available():
            raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
        if REF not in sys.path:
            sys.path.insert(0, REF)
        for name in ("config_utils", "model", "training_utils", "data_utils"):
            sys.modules.pop(name, None)
        self.config_utils = importlib.import_module("config_utils")
        T, V = cfg["block_size"], cfg["vocab"]
        self.M, self.B, self.T = len(V), batch, T
        self.config_utils._config_cache = dict(device=device, batch_size=batch, block_size=T, n_embd=cfg["n_embd"],
                                               n_head=cfg["n_head"], n_layer=cfg["n_layer"], dropout=0.0, eval_iters=1,
                                               fixed_values=[], output_file_name="", project_file_path="./")
        self.model_mod = importlib.import_module("model")
        self.tu = importlib.import_module("training_utils")
        assert os.path.dirname(os.path.abspath(self.model_mod.__file__)) == REF, "not the reference's model.py"
        streams = syn.make_streams(cfg, "cpu", rows)
        n_train = int(rows * 0.9)
        tu = self.tu
        tu.all_train_sets = [s[:n_train].tolist() for s in streams]          # python lists, as main.py hands them over
        tu.all_val_sets = [s[n_train:].clone() for s in streams]
        tu.all_full_datasets = None
        tu.all_vocabularies = [list(range(v)) for v in V]
        tu.all_modality_params = syn.modality_params(cfg)
        tu.all_file_info = [[n_train]] * self.M
        tu.file_lengths = [n_train]
        tu.num_modalities = self.M
        tu.is_percents = False
        torch.manual_seed(seed)
        self.m = self.model_mod.MultimodalTransformer(self.M, list(V), tu.all_modality_params).to(device)
        tu.m = self.m
        self.opt = torch.optim.AdamW(self.m.parameters(), lr=lr)
        self.device, self.mode = device, mode
        if mode == "tf32":
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        torch.manual_seed(1)

    def step(self):
        """The loop body of main.py:641-650."""
        xb_list, yb_list = self.tu.get_batch('train', 1)
        if self.mode == "autocast":
            with torch.autocast(torch.device(self.device).type, dtype=torch.bfloat16):
                logits_list, losses_list = self.m(xb_list, yb_list)
        else:
            logits_list, losses_list = self.m(xb_list, yb_list)
        total_loss = sum(losses_list)
        self.opt.zero_grad(set_to_none=True)
        total_loss.backward()
        self.opt.step()
        return total_loss

    def timed(self, steps, warmup):
        """seconds per step (wall clock with a device synchronize on both sides), last loss"""
        cuda = torch.device(self.device).type == "cuda"
        loss = None
        for _ in range(warmup):
            loss = self.step()
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = self.step()
        if cuda:
            torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps, float(loss.detach())
