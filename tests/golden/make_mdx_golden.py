"""Generate MDX-Net golden vectors by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_mdx_golden.py

It imports /root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py unmodified.  That
file imports ``librosa`` and ``soundfile`` at top level (mdxnet.py:6-8) and
``onnxruntime`` inside ``Predictor.__init__`` (:92); none of them is installed here and
none is touched by the spectral code we pin (``ConvTDFNetTrim.stft/istft`` :41-75,
``Predictor.demix`` :109-141, ``Predictor.demix_base`` :143-197), so empty stub modules
are registered for the two top-level imports and ``Predictor`` is built without
running ``__init__``; a seeded numpy mask function stands in for the ONNX session
(``_ort.run(None, {"input": spek})[0]``, mdxnet.py:171-176).

Outputs (small, committed): tests/golden/mdx_stft.npz, tests/golden/mdx_demix.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.synth import synth_mix  # noqa: E402  (input generator only)


def load_reference_module():
    for name in ("librosa", "soundfile"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("ref_mdxnet", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeOrt:
    """Deterministic stand-in for the ONNX session: y = 0.75*x + 0.1*flip_channels(x) + 0.05*tanh(x).

    Non-linear and not odd-symmetric-only, so the denoise branch (:168-172) is exercised.
    """

    def run(self, _names, feeds):
        x = feeds["input"]
        return [(0.75 * x + 0.1 * x[:, ::-1] + 0.05 * np.tanh(x) + 0.01 * x * x).astype(np.float32)]


def main():
    ref = load_reference_module()
    torch.manual_seed(0)

    # ---- (1) ConvTDFNetTrim.stft / istft, small dim_t so the fixture stays small -------------
    out = {}
    for tag, (n_fft, dim_f, dim_t_log2) in {"a": (6144, 3072, 4), "b": (2048, 1024, 5)}.items():
        m = ref.ConvTDFNetTrim(device=ref.cpu, model_name="Conv-TDF", target_name="vocals", L=11,
                               dim_f=dim_f, dim_t=dim_t_log2, n_fft=n_fft)
        seed = 100 + ord(tag)
        waves = synth_mix(2 * m.chunk_size, seed=seed).reshape(2, 2, m.chunk_size).transpose(1, 0, 2)
        waves = np.ascontiguousarray(waves)          # [N=2, 2, chunk]
        x = torch.tensor(waves)
        spek = m.stft(x)
        back = m.istft(spek)
        # keep a strided subset of the spectrum (every 7th bin) + the full round trip
        out[f"{tag}_cfg"] = np.array([n_fft, dim_f, dim_t_log2, seed], dtype=np.int64)
        out[f"{tag}_spek_sub"] = spek.numpy()[:, :, ::7, :].astype(np.float32)
        out[f"{tag}_spek_sum"] = np.array([float(spek.double().sum()), float(spek.double().abs().sum())])
        out[f"{tag}_istft"] = back.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "mdx_stft.npz"), **out)

    # ---- (2) Predictor.demix / demix_base with a fake ORT session -----------------------------
    class Args:
        pass

    res = {}
    for tag, (n_fft, dim_f, dim_t_log2, n, chunks, margin, denoise) in {
        "a": (6144, 3072, 4, 20001, 0, 44100, False),      # single segment, 3 chunks, ragged tail
        "b": (6144, 3072, 4, 50003, 1, 4410, True),        # margin segmentation + denoise branch
        "c": (2048, 1024, 4, 0, 0, 44100, False),          # n = 2*gen_size -> n % gen_size == 0 -> a full extra pad chunk
    }.items():
        args = Args()
        args.dim_f, args.dim_t, args.n_fft = dim_f, dim_t_log2, n_fft
        args.margin, args.chunks, args.denoise = margin, chunks, denoise
        pred = object.__new__(ref.Predictor)
        pred.args = args
        pred.model_ = ref.get_models(device=ref.cpu, dim_f=dim_f, dim_t=dim_t_log2, n_fft=n_fft)
        pred.model = FakeOrt()
        seed = 200 + ord(tag)
        if tag == "c":
            gen = pred.model_.chunk_size - n_fft
            n = 2 * gen
        mix = synth_mix(n, seed=seed)
        sources = pred.demix(mix)                    # [1, 2, n] float64
        res[f"{tag}_cfg"] = np.array([n_fft, dim_f, dim_t_log2, n, chunks, margin, int(denoise), seed],
                                     dtype=np.int64)
        res[f"{tag}_out"] = np.asarray(sources)[0].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "mdx_demix.npz"), **res)
    for f in ("mdx_stft.npz", "mdx_demix.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
