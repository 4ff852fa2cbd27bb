"""The oracle against the golden vectors produced by the REFERENCE's own mdxnet.py
(tests/golden/make_mdx_golden.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import mdx as omdx
from oracle.metrics import max_abs_err
from oracle.synth import synth_mix

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class FakeOrtNet:
    """Same function as FakeOrt in make_mdx_golden.py."""

    def __call__(self, x):
        x = x.numpy()
        return torch.tensor((0.75 * x + 0.1 * x[:, ::-1] + 0.05 * np.tanh(x) + 0.01 * x * x).astype(np.float32))


def test_stft_istft_match_reference_golden():
    g = np.load(os.path.join(GOLD, "mdx_stft.npz"))
    for tag in ("a", "b"):
        n_fft, dim_f, dtl, seed = [int(v) for v in g[f"{tag}_cfg"]]
        cfg = omdx.MdxConfig(n_fft=n_fft, dim_f=dim_f, dim_t_log2=dtl)
        spec = omdx.MdxSpec(cfg)
        waves = synth_mix(2 * cfg.chunk_size, seed=seed).reshape(2, 2, cfg.chunk_size).transpose(1, 0, 2)
        x = torch.tensor(np.ascontiguousarray(waves))
        spek = spec.stft(x)
        assert max_abs_err(spek.numpy()[:, :, ::7, :], g[f"{tag}_spek_sub"]) == 0.0
        sums = np.array([float(spek.double().sum()), float(spek.double().abs().sum())])
        assert np.allclose(sums, g[f"{tag}_spek_sum"], rtol=0, atol=0)
        assert max_abs_err(spec.istft(spek).numpy(), g[f"{tag}_istft"]) == 0.0


def test_demix_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "mdx_demix.npz"))
    for tag in ("a", "b", "c"):
        n_fft, dim_f, dtl, n, chunks, margin, den, seed = [int(v) for v in g[f"{tag}_cfg"]]
        cfg = omdx.MdxConfig(n_fft=n_fft, dim_f=dim_f, dim_t_log2=dtl, denoise=bool(den))
        out = omdx.demix_segments(synth_mix(n, seed=seed), FakeOrtNet(), cfg, chunks=chunks, margin=margin)
        assert out.shape == g[f"{tag}_out"].shape
        assert max_abs_err(out, g[f"{tag}_out"]) == 0.0      # same torch ops, same machine class: bit-exact


def test_windowed_form_reduces_to_identity_with_identity_net():
    cfg = omdx.MdxConfig(n_fft=2048, dim_f=1025, dim_t_log2=5)
    mix = synth_mix(70000, seed=1)
    out = omdx.demix_windowed(mix, omdx.identity_net, cfg)
    assert max_abs_err(out, mix) < 2e-6
