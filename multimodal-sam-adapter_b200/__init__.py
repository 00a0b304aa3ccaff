"""mmsam_b200 — B200-native MM SAM-Adapter encoder forward path (hand-written sm_100a kernels behind
a C ABI; PyTorch only for device memory, streams and torch.distributed)."""
__version__ = "0.1.0"
