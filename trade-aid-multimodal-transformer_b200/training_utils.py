"""Drop-in replacement for the public surface of /root/reference/training_utils.py:

    generate_batch_starting_indices(...)   (:33)    same validation, messages and RNG consumption
    get_batch(split, is_training)          (:333)   token store resident in HBM, one gather kernel (K7)
    calculate_evaluation_metrics(...)      (:215)   same counts, computed from the last-position logits in one pass
    estimate_loss(current_step, max_steps) (:387)   same prints and log lines

and the module globals main.py injects (main.py:387-396, 485): all_train_sets, all_val_sets, all_vocabularies,
all_modality_params, all_file_info, file_lengths, num_modalities, is_percents, m.
"""
from __future__ import annotations

import numbers
import os
from datetime import datetime

import torch
import torch.nn.functional as F

from .feeder import TokenStore, generate_batch_starting_indices  # noqa: F401  (re-exported)

__all__ = ['generate_batch_starting_indices', 'get_batch', 'estimate_loss', 'calculate_evaluation_metrics']

_explicit_cfg = None


def configure(**kw):
    """Standalone alternative to the reference's config_utils cache: device, batch_size, block_size, eval_iters,
    output_file_name, project_file_path."""
    global _explicit_cfg
    _explicit_cfg = dict(output_file_name='', project_file_path='')
    _explicit_cfg.update(kw)


def _cfg():
    if _explicit_cfg is not None:
        return _explicit_cfg
    from config_utils import _get_config
    return _get_config()


def _get(key):
    return _cfg()[key]


# ------------------------------------------------------------------------------------------------ feeder
_stores = {}
_ops_override = None   # tests inject a CPU emulation of the kernel namespace; the product leaves this None


def _store(split):
    """Device-resident token store of a split, rebuilt when main.py (re)assigns the global lists."""
    src = all_train_sets if split == 'train' else all_val_sets
    key = (split, id(src), tuple(id(s) for s in src), str(_get('device')))
    ent = _stores.get(split)
    if ent is None or ent[0] != key:
        dev = _get('device')
        ent = (key, TokenStore(src, dev, ops=_ops_override))
        _stores[split] = ent
    return ent[1]


batch_shard = None   # (rank, world) in data-parallel PARITY mode; None: every process draws its own batches


def get_batch(split, is_training, shard=True):
    """training_utils.py:333-384. Returns (xb_list, yb_list): per modality int64 [batch_size, block_size] on `device`.
    With `batch_shard = (rank, world)` (data-parallel parity mode, fixed GLOBAL batch_size): every rank draws the same
    batch_size indices from its identically seeded host generator - the CPU RNG stream stays bit-identical to the
    single-process reference - and gathers only rows [rank*B/world, (rank+1)*B/world); the shards of all ranks
    concatenate to the single-process batch. `shard=False` returns the whole batch regardless (estimate_loss: an
    evaluation batch is never split by rows - its iterations are dealt to the ranks instead)."""
    train_store = _store('train')
    for r in range(num_modalities):
        this_rand_size = all_modality_params[r][2]
        if this_rand_size is not None and is_training == 1:
            # The reference reads the augmentation size from index 2 of the legacy list, which holds `has_header`
            # (training_utils.py:353; schema.py:213-248): True acts as rand_size 1 - a +-1 random walk applied to the whole
            # training list in place, cumulatively, on every call; False fails add_rand_to_data_points' range check.
            # Kept as shipped, but on the device-resident store (TokenStore.augment_modality: same values, same
            # consumption of Python's `random` stream, same errors); the host lists are not rewritten - the reference
            # reads them nowhere else.
            train_store.augment_modality(r, this_rand_size, len(all_vocabularies[r]))
    store = train_store if split == 'train' else _store('val')
    T, B = _get('block_size'), _get('batch_size')
    ix = generate_batch_starting_indices(store.N, T, B, split, file_lengths, is_percents)
    if batch_shard is not None and shard:
        rank, world = batch_shard
        if B % world != 0:
            raise ValueError(f"batch_size ({B}) must be divisible by the number of data-parallel ranks ({world})")
        ix = ix[rank * (B // world):(rank + 1) * (B // world)]
    xb, yb = store.gather(ix, T)
    return [xb[r] for r in range(num_modalities)], [yb[r] for r in range(num_modalities)]


# ------------------------------------------------------------------------------------------------ metrics
_vocab_tables = {}


def _vocab_table(vocab, device):
    """Values of a numeric vocabulary as a device fp64 tensor (python floats are doubles: the comparisons of
    training_utils.py:198-211 stay exact). Cached per vocabulary object."""
    key = (id(vocab), len(vocab), str(device))
    ent = _vocab_tables.get(key)
    if ent is None or ent[0] is not vocab:
        ent = (vocab, torch.tensor([float(v) for v in vocab], dtype=torch.float64, device=device))
        _vocab_tables[key] = ent
    return ent[1]


def _metric_ops(device):
    if _ops_override is not None:
        return _ops_override
    if torch.device(device).type != 'cuda':
        raise RuntimeError("tradeaid_b200.training_utils computes the directional metrics with a CUDA kernel "
                           "(ta_directional_metrics); logits must live on the GPU - there is no CPU path")
    from . import lib
    return lib


def _accumulate_metrics(logits_list, xb_list, yb_list, num_modalities, all_vocabularies, all_modality_params, counts, cert,
                        processed):
    """One ta_directional_metrics launch per numeric modality: counts int64 [M, 2] (wins, losses) and cert fp64 [M] are
    device accumulators; processed (host list) counts the batches that took part (training_utils.py:262)."""
    for i in range(num_modalities):
        if not (len(logits_list) > i and len(yb_list) > i):
            continue
        vocab = all_vocabularies[i]
        is_pct = bool(all_modality_params[i][3])
        numeric = all(isinstance(v, numbers.Number) for v in vocab)
        y = yb_list[i]
        if not (numeric and y.ndim >= 2 and y.shape[1] >= (1 if is_pct else 2) and y.shape[0] > 0):
            continue
        processed[i] += 1
        if not is_pct and xb_list[i].shape[1] < 1:
            continue  # no previous value: direction undefined for every sample (training_utils.py:278-287)
        last = logits_list[i][:, -1, :]
        if last.dtype != torch.float32:
            last = last.float()
        dev = last.device
        ops = _metric_ops(dev)
        prev = None if is_pct else xb_list[i][:, -1].to(dev)
        ops.directional_metrics(last, _vocab_table(vocab, dev), prev, y[:, -1].to(dev), is_pct, counts[i], cert[i:i + 1])


def calculate_evaluation_metrics(logits_list, xb_list, yb_list, num_modalities, all_vocabularies, all_modality_params,
                                 all_file_info):
    """training_utils.py:215-330 on the device: directional wins / losses of the arg-max prediction at the last position
    against the actual next value, and the probability mass on the predicted direction — one kernel per modality and
    ONE read-back per call (the reference does batch x vocabulary .item() calls)."""
    dev = logits_list[0].device if len(logits_list) else 'cpu'
    counts = torch.zeros(num_modalities, 2, dtype=torch.int64, device=dev)
    cert = torch.zeros(num_modalities, dtype=torch.float64, device=dev)
    processed = [0] * num_modalities
    _accumulate_metrics(logits_list, xb_list, yb_list, num_modalities, all_vocabularies, all_modality_params, counts, cert,
                        processed)
    c, ce = counts.tolist(), cert.tolist()
    return [r[0] for r in c], [r[1] for r in c], [float(x) for x in ce], [min(p, 1) for p in processed]


eval_process_group = None   # data-parallel runs set this (torch.distributed group): see estimate_loss


def _all_reduce_eval(pg, dev, loss_sum, n_losses, counts, cert, tot_proc):
    """Sum of the per-rank evaluation totals: ONE all-reduce of 2 + 4M float64 scalars (loss sum, number of losses; per
    modality wins, losses, certainty mass, batches processed) - SURVEY.md 8e. Counts are exact in float64 (< 2^53)."""
    import torch.distributed as dist
    M = num_modalities
    pack = torch.zeros(2 + 4 * M, dtype=torch.float64, device=dev)
    pack[0], pack[1] = loss_sum, n_losses
    if counts is not None:
        pack[2:2 + 2 * M] = counts.to(torch.float64).reshape(-1)
        pack[2 + 2 * M:2 + 3 * M] = cert
    pack[2 + 3 * M:] = torch.tensor(tot_proc, dtype=torch.float64)
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=pg)
    host = pack.tolist()
    counts_h = [[int(host[2 + 2 * i]), int(host[3 + 2 * i])] for i in range(M)]
    return host[0], int(host[1]), counts_h, host[2 + 2 * M:2 + 3 * M], [int(v) for v in host[2 + 3 * M:]]


def estimate_loss(current_step=None, max_steps=None):
    """training_utils.py:387-520: eval_iters batches of each split through the model in eval mode; mean summed loss,
    directional metrics, same console and log-file lines.
    Data parallel (`eval_process_group` set): the eval_iters iterations are dealt round-robin to the ranks.
    EVERY rank still draws every batch (get_batch is a host index draw + one gather launch), so the identically seeded host
    RNG streams stay in step and the union of the evaluated batches is exactly the single-process set; the forward and the
    metrics run only on the owner. The totals are summed with one all-reduce per split: counts are identical to the
    single-process run, the mean loss up to the order of a float64 sum. Rank 0 prints and writes the log lines."""
    out = {}
    m.eval()
    eval_iters, batch_size = _get('eval_iters'), _get('batch_size')
    pg = eval_process_group
    if pg is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    else:
        world, rank = 1, 0
    say = print if rank == 0 else (lambda *a, **k: None)
    for state in ['train', 'val']:
        current_time = datetime.now().strftime("%H:%M:%S")
        step_info = f'Step {current_step}/{max_steps} | ' if current_step is not None else ''
        batch_calc = f' * {batch_size} batches = {eval_iters * batch_size} samples'
        say(f'Evaluation: {step_info}{state.title()} set ({eval_iters} iterations{batch_calc}) | {current_time}')
        tot_proc = [0] * num_modalities
        counts = cert = None
        total_losses = []
        for k in range(eval_iters):
            # whole batches even in data-parallel parity mode (`batch_shard`): the owner of iteration k evaluates all
            # batch_size samples of it, so the totals are those of the single-process run
            xb_list, yb_list = get_batch(state, 0, shard=False)
            if k % world != rank:
                continue
            with torch.no_grad():
                logits_list, losses_list = m(xb_list, yb_list)
            if losses_list and all(l is not None for l in losses_list):
                total_losses.append(sum(losses_list).detach())
            else:
                say(f"Warning: Iteration {k} losses not calculated, skipping")
            if counts is None:   # device accumulators: one read-back per split instead of B*V .item() calls per batch
                dev = logits_list[0].device
                counts = torch.zeros(num_modalities, 2, dtype=torch.int64, device=dev)
                cert = torch.zeros(num_modalities, dtype=torch.float64, device=dev)
            _accumulate_metrics(logits_list, xb_list, yb_list, num_modalities, all_vocabularies, all_modality_params,
                                counts, cert, tot_proc)
        total_losses = torch.stack(total_losses).tolist() if total_losses else []
        if world > 1:
            dev = torch.device(_get('device'))
            loss_sum, n_losses, c, tot_cert, tot_proc = _all_reduce_eval(pg, dev, sum(total_losses), len(total_losses), counts,
                                                                         cert, tot_proc)
            out[state] = loss_sum / n_losses if n_losses else float('nan')
        else:
            c = counts.tolist() if counts is not None else [[0, 0]] * num_modalities
            tot_cert = cert.tolist() if cert is not None else [0.0] * num_modalities  # noqa: F841 (reference keeps it too)
            out[state] = sum(total_losses) / len(total_losses) if total_losses else float('nan')
        tot_ok, tot_bad = [r[0] for r in c], [r[1] for r in c]
        disp = 'Train Set' if state == 'train' else 'Val Set'
        if rank != 0:
            continue
        print(f"\nDIRECTIONAL METRICS - {disp} (Correct/Total)")
        for i in range(num_modalities):
            name = all_modality_params[i][9] if all_modality_params[i][9] else f"Modality {i+1}"
            if tot_proc[i] > 0:
                total = tot_ok[i] + tot_bad[i]
                if total > 0:
                    print(f"  - {name:<30}{tot_ok[i]}/{total} ({round((tot_ok[i] / total) * 100, 1)}%)")
                else:
                    print(f"  - {name}: No directional predictions")
            else:
                print(f"  - {name}: No data processed (non-numeric)")
        cfg = _cfg()
        if cfg['output_file_name'] != '':
            path = cfg['project_file_path'] + 'output/' + cfg['output_file_name']
            with open(path, 'a', encoding='utf-8') as f:
                for i in range(num_modalities):
                    name = all_modality_params[i][9] if all_modality_params[i][9] else f"Modality {i+1}"
                    if tot_proc[i] > 0:
                        total = tot_ok[i] + tot_bad[i]
                        acc = f"{round((tot_ok[i] / total) * 100, 1)}%" if total > 0 else "N/A"
                        f.write(f"   DIRECTIONAL PREDICTION {disp} - {name}: Correct={tot_ok[i]:,} | Incorrect={tot_bad[i]:,} | Accuracy={acc}\n")
                    else:
                        f.write(f"   DIRECTIONAL PREDICTION {disp} - {name}: Correct=0 | Incorrect=0 | Accuracy=N/A\n")
                if state == 'train':
                    f.write("\n")
        if state == 'train':
            print()
    m.train()
    return out


# Globals set by main.py after data loading (training_utils.py:523-534)
all_full_datasets = None
all_train_sets = None
all_val_sets = None
all_vocabularies = None
all_modality_params = None
all_file_info = None
file_lengths = None
num_modalities = None
is_percents = False
m = None
