"""tradeaid_b200 — B200-native (sm_100a) engine for the hot path of tsnuk/trade-AId-multimodal-transformer:
model.py forward/backward + training_utils.get_batch. Host code is Python/PyTorch (device memory, streams,
torch.distributed); every hot-path op is a hand-written CUDA kernel in libtradeaid_b200.so behind a C ABI
(include/tradeaid_b200.h). There is no CPU fallback: importing ``tradeaid_b200.lib`` raises if the library is missing."""
__version__ = "0.1.0"
