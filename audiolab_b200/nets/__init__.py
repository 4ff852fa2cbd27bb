"""Mask networks (PyTorch modules) that sit between the STFT and iSTFT kernels."""
