"""audiolab_b200 -- B200-native spectral demix engine behind AudioLab's Separate API.

Hot path: pad-and-chunk -> STFT -> mask net -> (mask (.) spec) -> iSTFT ->
overlap-add, plus the polyphase resampler around it, as hand-written sm_100a kernels behind
the C ABI in include/audiolab_b200.h.  Host code is Python/PyTorch (device memory, streams,
torch.distributed); there is no CPU fallback.
"""
__version__ = "0.1.0"
