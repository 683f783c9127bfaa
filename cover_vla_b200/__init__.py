"""cover_vla_b200 - B200-native (sm_100a) sample-and-verify path of CoVer-VLA.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in
hand-written CUDA kernels behind the C ABI declared in include/coverb200.h.
"""
__all__ = ["lib"]
