"""``separate_music`` -- the call ``Separate.process_audio`` makes (reference:
/root/reference/modules/separator/stem_separator.py:949-1001 separate_music, :847-946
predict_with_model, :357-413 ensemble loop, :241-262 _blend_tracks, :625-677 _save_all_stems).

Scope: the vocals/instrumental ensemble over the RoFormer / MDX-Net models on
the hot path, kept ON THE DEVICE between models -- the reference writes a PCM_16 temp WAV, runs
``separator.separate`` and re-loads the outputs for every model x file (:264-355).  Options that need
architectures outside the hot path (VR de-noise / de-reverb, MDX23C drum split, ...) raise
``NotImplementedError`` instead of silently doing nothing.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .separator import Separator
from .wavio import read_wav, write_wav

logger = logging.getLogger(__name__)

# (model, vocals weight, instrumental weight) -- stem_separator.py:379-387 / :873-879
ENSEMBLE = [
    ("vocals_mel_band_roformer.ckpt", 8.6, 16.0),
    ("model_bs_roformer_ep_368_sdr_12.9628.ckpt", 8.4, 16.0),
    ("melband_roformer_big_beta4.ckpt", 8.5, 16.0),
    ("MDX23C-8KFFT-InstVoc_HQ.ckpt", 7.2, 14.9),
    ("UVR-MDX-NET-Voc_FT.onnx", 6.9, 14.9),
]

_OUT_OF_SCOPE = {
    "separate_bg_vocals": False, "separate_drums": False, "separate_woodwinds": False, "alt_bass_model": False,
    "reverb_removal": "Nothing", "echo_removal": "Nothing", "delay_removal": "Nothing",
    "crowd_removal": "Nothing", "noise_removal": "Nothing", "store_reverb_ir": False,
}


def blend_tracks(tracks: List[torch.Tensor], weights: List[float]) -> torch.Tensor:
    """stem_separator.py:241-262: weighted mean over models, then peak-normalise to 1."""
    n = max(t.shape[-1] for t in tracks)
    combined = torch.zeros((tracks[0].shape[0], n), dtype=torch.float32, device=tracks[0].device)
    total = max(sum(weights), 1e-6)
    for i, t in enumerate(tracks):
        w = weights[i] if i < len(weights) else 1.0
        combined[:, : t.shape[-1]] += t * float(w)
    combined = combined / total
    peak = combined.abs().max()
    return torch.where(peak > 0, combined / peak, combined)


def separate_music(input_dict: Dict[str, List[str]], callback: Optional[Callable] = None,
                   separator: Optional[Separator] = None, **kwargs) -> List[str]:
    """{output_folder: [input paths]} -> list of written stem paths.  ``callback(fraction, desc, total)``."""
    for key, off in _OUT_OF_SCOPE.items():
        if kwargs.get(key, off) not in (off, None):
            raise NotImplementedError(f"{key}={kwargs[key]!r} needs a model family outside this engine's scope")
    if not kwargs.get("vocals_only", True):
        raise NotImplementedError("multi-stem (htdemucs_6s / drumsep / woodwinds) orchestration is not wired up; "
                                  "use Separator.load_model('htdemucs_ft.yaml') directly")
    strength = int(kwargs.get("ensemble_strength", 2))
    models = [m for m in ENSEMBLE if not m[0].startswith("MDX23C")][:strength]
    files = []
    for out_folder, paths in input_dict.items():
        for p in paths:
            if os.path.isfile(p):
                files.append((out_folder, p))
    if not files:
        return []
    sep = separator or Separator(log_level=logging.ERROR, invert_using_spec=True, use_autocast=True,
                                 model_file_dir=kwargs.get("model_file_dir", "models/audio_separator"),
                                 allow_random_init=bool(kwargs.get("allow_random_init", False)))
    total_steps = len(models) * len(files) + 1 + len(files)
    step = 0
    if callback is not None:
        callback(0, "Starting ensemble separation...", total_steps)
    mixes = []
    for _, p in files:
        audio, sr = read_wav(p)
        mixes.append(sep.prepare_mix(torch.from_numpy(audio), sr))
    per_file = [dict(vocals=[], instrumental=[]) for _ in files]
    wv, wi = [], []
    for name, w_voc, w_inst in models:
        sep.load_model(name)
        wv.append(w_voc)
        wi.append(w_inst)
        for i, mix in enumerate(mixes):
            stems = sep.separate_tensor(mix)
            per_file[i]["vocals"].append(stems["Vocals"])
            per_file[i]["instrumental"].append(stems["Instrumental"])
            step += 1
            if callback is not None:
                callback(step / total_steps, f"{name}: {os.path.basename(files[i][1])}", total_steps)
    outputs: List[str] = []
    for (out_folder, p), res in zip(files, per_file):
        os.makedirs(out_folder, exist_ok=True)
        base = os.path.splitext(os.path.basename(p))[0]
        for stem, tag, w in (("vocals", "(Vocals)", wv), ("instrumental", "(Instrumental)", wi)):
            blended = blend_tracks(res[stem], w)
            path = os.path.join(out_folder, f"{base}_{tag}.wav")
            write_wav(path, blended.cpu().numpy(), sep.sample_rate, subtype="FLOAT")
            outputs.append(path)
        step += 1
        if callback is not None:
            callback(step / total_steps, f"Saved stems for {base}", total_steps)
    if callback is not None:
        callback(1.0, "Separation complete", total_steps)
    return outputs
