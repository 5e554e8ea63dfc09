"""The feeder of the hot path: training_utils.get_batch (/root/reference/training_utils.py:333-384) with the token
store resident in HBM and the window gather done by one kernel (K7, ta_gather_batch).

Bit-exactness contract (BASELINE.json north_star): for a given seed the starting indices AND the int64 token ids are
identical to the reference's. Indices are therefore drawn exactly as the reference draws them — one `torch.randint`
on the default CPU generator per call, same bounds, same order (training_utils.py:104, 150) — and only the O(B*T*M)
gather runs on the device. What the reference does per call on the host (torch.tensor(list) of the WHOLE train set,
training_utils.py:363; python slicing + stack, :378-379) is replaced by a one-time upload.
"""
from __future__ import annotations

import torch


def split_file_lengths(data_size, split, file_lengths):
    """Lengths of the files (or file parts) that make up the current split (training_utils.py:109-135): the train split
    consumes files from the front, the val split from the back; the last piece may be a partial file."""
    out, acc = [], 0
    n = len(file_lengths)
    for f in range(n):
        this = file_lengths[f] if split == "train" else file_lengths[n - 1 - f]
        acc += this
        out.append(this if acc <= data_size else data_size - (acc - this))
        if acc >= data_size:
            if split == "val":  # accumulated from the back: restore file order (:132-134)
                out.reverse()
            break
    return out


def generate_batch_starting_indices(data_size, block_size, batch_size, split, file_lengths, is_percents):
    """Same contract, validation, error messages and RNG consumption as the reference function of the same name
    (training_utils.py:33-181); the per-sample python search (:156-178) is a vectorised searchsorted."""
    if not isinstance(data_size, int) or data_size <= 0:
        raise TypeError("'data_size' must be a positive integer.")
    if not isinstance(block_size, int) or block_size <= 0:
        raise TypeError("'block_size' must be a positive integer.")
    if block_size >= data_size:
        raise ValueError("'block_size' cannot be equal to or greater than 'data_size'.")
    if not isinstance(batch_size, int) or batch_size <= 0:
        raise TypeError("'batch_size' must be a positive integer.")
    if not isinstance(split, str) or (split != 'train' and split != 'val'):
        raise ValueError("'split' must be 'train' or 'val'.")
    if not isinstance(file_lengths, list) or not len(file_lengths) >= 1:
        raise TypeError("'file_lengths' must be a list containing at least 1 element.")
    if not isinstance(is_percents, bool):
        raise TypeError("'is_percents' must be a boolean.")
    span = block_size + 1
    first = 1 if is_percents else 0
    if len(file_lengths) == 1:
        return torch.randint(first, data_size - span + 1, (batch_size,))
    lens = torch.tensor(split_file_lengths(data_size, split, file_lengths), dtype=torch.long)
    valid = (lens - span - first + 1).clamp_min(0)
    total = int(valid.sum())
    if total <= 0:
        raise ValueError("No valid starting positions available for the given block size and file lengths.")
    r = torch.randint(total, (batch_size,))
    ends = torch.cumsum(valid, 0)
    k = torch.searchsorted(ends, r, right=True)
    file_start = torch.cumsum(lens, 0) - lens
    return file_start[k] + (r - (ends[k] - valid[k])) + first


class TokenStore:
    """int32 token ids of every modality of one split, [M, N] in device memory, + the staging for the indices."""

    def __init__(self, data_list, device, ops=None):
        if ops is None:
            from . import lib as ops
        self.ops = ops
        self.device = torch.device(device)
        rows = []
        for d in data_list:
            t = d if torch.is_tensor(d) else torch.tensor(d, dtype=torch.long)
            if t.numel() and (int(t.max()) > 2 ** 31 - 1 or int(t.min()) < 0):
                raise ValueError("token ids must fit int32")
            rows.append(t.to(torch.int32))
        n = {int(r.numel()) for r in rows}
        if len(n) != 1:
            raise ValueError(f"all modalities must have the same length, got {sorted(n)}")
        self.N = n.pop()
        self.M = len(rows)
        self.store = torch.stack(rows).to(self.device)
        self._ix_host = None

    def update_modality(self, r, data):
        t = data if torch.is_tensor(data) else torch.tensor(data, dtype=torch.long)
        self.store[r].copy_(t.to(torch.int32))

    def augment_modality(self, r, rand_size, vocab_size):
        """data_utils.add_rand_to_data_points (data_utils.py:293-358) applied IN PLACE to modality r of the resident
        store, bit-exact to the reference's host loop and to its use of Python's module-level `random` generator:
        the Mersenne-Twister words the loop would draw are pre-drawn in bulk (random.getrandbits(32 * W): its 32-bit
        limbs are the successive outputs), matched on the device with the eligible tokens (ta_augment_tokens), and the
        generator is then re-positioned to exactly the number of words the reference would have consumed.
        Same argument checks and messages as the reference (rand_size comes from all_modality_params[r][2])."""
        import random
        if not isinstance(rand_size, (int, type(None))):
            raise TypeError("rand_size must be an integer or null.")
        if rand_size is not None and (rand_size < 1 or rand_size > 3):
            raise ValueError("rand_size must be an integer between 1 and 3, or null.")
        if not isinstance(vocab_size, int) or vocab_size <= 0:
            raise TypeError("vocab_size must be a positive integer.")
        if rand_size is None:
            return
        if self.N == 0:
            raise ValueError("numeric_data cannot be empty.")
        s = int(rand_size)
        n = 2 * s + 1
        k = n.bit_length()
        state0 = random.getstate()
        W = int(self.N * (2 ** k / n) * 1.03) + 4096       # expected words if every token were eligible, + slack
        stats = torch.zeros(3, dtype=torch.int64, device=self.device)
        while True:
            random.setstate(state0)
            big = random.getrandbits(32 * W)
            words = torch.frombuffer(bytearray(big.to_bytes(4 * W, "little")), dtype=torch.int32).to(self.device)
            ws = torch.empty(self.ops.augment_workspace_bytes(self.N, W), dtype=torch.uint8, device=self.device)
            self.ops.augment_tokens(self.store[r], words, s, int(vocab_size), ws, stats)
            eligible, accepted, consumed = stats.tolist()
            if accepted >= eligible:
                break
            W *= 2                                           # (astronomically unlikely) not enough accepted draws: redo
        random.setstate(state0)
        if consumed:
            random.getrandbits(32 * consumed)                # leave the generator where the reference's loop leaves it

    RING = 4  # pinned index staging buffers: the host may run this many steps ahead of the device

    def gather(self, ix, block_size, out=None):
        """ix: int64 [B] on the HOST (as torch.randint made it). Returns (xb, yb): int64 [M, B, T] device tensors —
        fresh ones per call like the reference, or the caller's static pair `out` (CUDA-graph replay).
        Asynchronous: pinned staging -> cudaMemcpyAsync -> gather kernel, all on the current stream."""
        B = int(ix.shape[0])
        pin = self.device.type == "cuda"
        if self._ix_host is None or self._ix_host[0].shape[0] != B:
            self._ix_host = [torch.empty(B, dtype=torch.int64, pin_memory=pin) for _ in range(self.RING)]
            self._ix_dev = torch.empty(B, dtype=torch.int64, device=self.device)
            self._ix_event = [torch.cuda.Event() if pin else None for _ in range(self.RING)]
            self._ix_used = [False] * self.RING
            self._slot = 0
        k = self._slot
        self._slot = (k + 1) % self.RING
        if self._ix_used[k] and self._ix_event[k] is not None:
            self._ix_event[k].synchronize()   # the async H2D that last read this staging buffer has completed
        self._ix_host[k].copy_(ix)
        self._ix_dev.copy_(self._ix_host[k], non_blocking=True)
        if out is None:
            out = (torch.empty(self.M, B, block_size, dtype=torch.int64, device=self.device),
                   torch.empty(self.M, B, block_size, dtype=torch.int64, device=self.device))
        xb, yb = out
        self.ops.gather_batch(self.store, self._ix_dev, block_size, xb, yb)
        if self._ix_event[k] is not None:
            self._ix_event[k].record()
            self._ix_used[k] = True
        return xb, yb
