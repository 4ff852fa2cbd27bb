"""``separate_music`` -- the call ``Separate.process_audio`` makes (reference:
/root/reference/modules/separator/stem_separator.py:949-1001 separate_music, :847-946
predict_with_model, :357-457 ensemble loop + de-bleed, :173-239 _residual_subtract, :241-262 _blend_tracks,
:459-503 6-stem stage, :625-677 _save_all_stems).

Scope: the vocals/instrumental ensemble over the RoFormer / MDX-Net models on
the hot path, the post-blend residual de-bleed, and the htdemucs_6s multi-stem stage, all kept ON THE
DEVICE between stages -- the reference writes a PCM_16 temp WAV, runs ``separator.separate`` and
re-loads the outputs for every model x file (:264-355).  ``pcm16_handoff=True`` reproduces that
quantisation (:57-75) so the pipeline can be compared with the reference's sample for sample.
The post-ensemble transform chain (:777-839, call order :903-934) runs the reverb / echo removal (Mel-band RoFormer
checkpoints) and crowd removal (MDX-Net) models the same way; options that need architectures outside the hot path (the VR
de-noise model, MDX23C drum split, background-vocal split, ...) raise ``NotImplementedError``
instead of silently doing nothing.  ``store_reverb_ir`` writes ``impulse_response.ir`` (reverb_ir.py) from the de-reverb
pass's two outputs (:823-829).
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from .separator import Separator, _arch_of
from .wavio import read_wav, write_wav

logger = logging.getLogger(__name__)

# (model, vocals weight, instrumental weight) -- stem_separator.py:379-387; `[:ensemble_strength]` of THIS list
ENSEMBLE = [
    ("vocals_mel_band_roformer.ckpt", 8.6, 16.0),
    ("model_bs_roformer_ep_368_sdr_12.9628.ckpt", 8.4, 16.0),
    ("melband_roformer_big_beta4.ckpt", 8.5, 16.0),
    ("MDX23C-8KFFT-InstVoc_HQ.ckpt", 7.2, 14.9),
    ("UVR-MDX-NET-Voc_FT.onnx", 6.9, 14.9),
    ("Kim_Vocal_2.onnx", 6.9, 14.9),
    ("Kim_Vocal_1.onnx", 6.8, 14.9),
]
SUPPORTED_ARCHS = ("bs_roformer", "mel_roformer", "mdx", "htdemucs")

# stem key -> label of the saved file `{base}__{label}.wav` (stem_separator.py:637-654)
STEM_LABELS = {
    "vocals": "(Vocals)", "instrumental": "(Instrumental)", "drums": "(Drums)", "bass": "(Bass)",
    "guitar": "(Guitar)", "piano": "(Piano)", "other": "(Other)",
}

_OUT_OF_SCOPE = {
    "separate_bg_vocals": False, "separate_drums": False, "separate_woodwinds": False, "alt_bass_model": False,
}
# (`delay_removal` is accepted and ignored, like the reference: its chain only looks at `echo_removal`, :797)
TRANSFORM_SETTINGS = ("Nothing", "All", "All Vocals", "Main Vocals")


def should_apply_transform(stem_name: str, setting: str) -> bool:
    """stem_separator.py:679-700: which stems ("(vocals)", "(instrumental)", "(bg_vocals ...)") a setting covers."""
    if setting == "All":
        return True
    if setting == "All Vocals":
        return "vocals)" in stem_name.lower()
    if setting == "Main Vocals":
        return "vocals)" in stem_name and "(bg_vocals" not in stem_name.lower()
    return False


def transformations(opts: Dict) -> List[tuple]:
    """(model file, label of the output to keep, setting) in the reference's order (:795-800, default models :147-149)."""
    return [
        ("dereverb_mel_band_roformer_anvuew_sdr_19.1729.ckpt", "No Reverb", opts.get("reverb_removal", "Nothing")),
        (opts.get("delay_removal_model", "dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt"), "dry",
         opts.get("echo_removal", "Nothing")),
        (opts.get("crowd_removal_model", "UVR-MDX-NET_Crowd_HQ_1.onnx"), "No Crowd", opts.get("crowd_removal", "Nothing")),
        (opts.get("noise_removal_model", "UVR-DeNoise.pth"), "No Noise", opts.get("noise_removal", "Nothing")),
    ]


def apply_transform_chain(sep, wav: torch.Tensor, stem_label: str, opts: Dict, skip_transforms=(), pcm16: bool = False,
                          on_step: Optional[Callable] = None, ir_path: Optional[str] = None) -> torch.Tensor:
    """stem_separator.py:777-839 on the device: every transform whose setting covers this stem loads its model, separates
    the CURRENT array and keeps the output whose name carries the transform's label (of two outputs: the first if it
    carries the label, else the second).  `pcm16` reproduces the PCM_16 temp WAV each model input goes through (:811)."""
    current = wav
    for model_file, out_label, setting in transformations(opts):
        if out_label in skip_transforms or not should_apply_transform(f"({stem_label})", setting):
            continue
        arch = _arch_of(model_file)
        if arch not in SUPPORTED_ARCHS:
            raise NotImplementedError(f"{out_label} removal with {model_file}: the {arch} architecture is outside this "
                                      "engine's scope (SURVEY.md section 8 row f3)")
        sep.load_model(model_file)
        stems = sep.separate_tensor(pcm16_roundtrip(current) if pcm16 else current)
        names = list(stems)
        key = out_label.replace(" ", "").lower()
        chosen = None
        if len(names) == 2:
            chosen = names[0] if key in names[0].replace(" ", "").lower() else names[1]
        else:
            chosen = next((n for n in names if key in n.replace(" ", "").lower()), None)
        if chosen is not None:
            current = stems[chosen]
            # :823-829: the reverb the de-reverb pass took out of the VOCALS, as an impulse response next to the stems
            if out_label == "No Reverb" and stem_label.lower() == "vocals" and ir_path and len(names) == 2:
                from .reverb_ir import extract_reverb
                alt = names[1] if chosen == names[0] else names[0]
                try:
                    extract_reverb(stems[chosen], stems[alt], sep.sample_rate, ir_path)
                except Exception as e:             # the reference logs and carries on (:828-829)
                    logger.error(f"Error extracting IR: {e}")
        if on_step is not None:
            on_step(f"TRANSFORM: {out_label} on {stem_label}")
    return current


def ensemble_models(strength: int):
    """The first `strength` entries of the reference's list; a model whose architecture this engine does not carry
    raises (the reference would run it) instead of being skipped silently."""
    chosen = ENSEMBLE[: max(0, int(strength))]
    for name, _, _ in chosen:
        arch = _arch_of(name)
        if arch not in SUPPORTED_ARCHS:
            raise NotImplementedError(
                f"ensemble_strength={strength} selects {name} ({arch}), which is outside this engine's scope "
                "(MDX-Net, BS/Mel-RoFormer, HTDemucs); use ensemble_strength <= 3")
    return chosen


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


def pcm16_roundtrip(x: torch.Tensor) -> torch.Tensor:
    """What survives ``write_temp_wav`` (sf.write subtype PCM_16, stem_separator.py:57-75) followed by a float load:
    round(x * 32768) clipped to int16, / 32768."""
    return torch.clamp(torch.round(x * 32768.0), -32768.0, 32767.0) / 32768.0


def residual_subtract(base: torch.Tensor, component: torch.Tensor, sr: int, max_shift_ms: float = 12.0) -> torch.Tensor:
    """stem_separator.py:173-239 on tensors (any device): per channel, align `component` to `base` by the lag of the
    largest cross-correlation within +-max_shift_ms over the first <= 1 s, least-squares gain clipped to [0, 1.25],
    subtract.  [C, n] x [C, m] -> [C, n] (samples past min(n, m) are `base`'s).

    The correlation and the two dot products are accumulated in float64: the lag is an argmax and must not depend on
    the summation order of a float32 reduction."""
    if base.dim() == 1:
        base = torch.stack((base, base))
    if component.dim() == 1:
        component = torch.stack((component, component))
    max_shift = max(0, int((max_shift_ms / 1000.0) * float(sr)))
    n = min(base.shape[-1], component.shape[-1])
    residual = base.clone()
    for ch in range(base.shape[0]):
        ref, sig = base[ch, :n], component[ch, :n]
        best = 0
        if max_shift > 0 and n > 0:
            probe = min(n, 44100)
            rp = F.pad(ref[:probe].double(), (max_shift, max_shift))
            # np.correlate(ref, sig, "full")[center + k] = sum_n ref[n + k] * sig[n],  k in [-max_shift, max_shift]
            corr = F.conv1d(rp[None, None], sig[:probe].double()[None, None])[0, 0]
            best = int(torch.argmax(corr)) - max_shift
        if best > 0:        # component lags the reference: pad the front
            aligned = torch.cat((torch.zeros(best, dtype=sig.dtype, device=sig.device), sig[:-best]))
        elif best < 0:      # component leads: pad the end
            aligned = torch.cat((sig[-best:], torch.zeros(-best, dtype=sig.dtype, device=sig.device)))
        else:
            aligned = sig
        a64 = aligned.double()
        alpha = float(torch.dot(ref.double(), a64)) / (float(torch.dot(a64, a64)) + 1e-8)
        alpha = min(max(alpha, 0.0), 1.25)
        residual[ch, :n] = ref - alpha * aligned
    return torch.nan_to_num(residual, nan=0.0, posinf=0.0, neginf=0.0)


def _cosine_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.reshape(-1).double(), b.reshape(-1).double()
    return float(torch.dot(a, b).abs() / (a.norm() * b.norm() + 1e-8))


def debleed_instrumental(mix: torch.Tensor, vocals: torch.Tensor, instrumental: torch.Tensor, sr: int,
                         residual_blend: float) -> torch.Tensor:
    """Post-blend de-bleed of stem_separator.py:414-456: the gain-matched residual ``mix - alpha * vocals`` is blended
    into the instrumental only when it is measurably less correlated with the vocals; a near-silent instrumental is
    replaced by the residual."""
    resid = residual_subtract(mix, vocals, sr)
    m = min(resid.shape[-1], instrumental.shape[-1])
    inst, res_m, voc = instrumental[:, :m], resid[:, :m], vocals[:, :m]
    out = instrumental
    if _cosine_abs(res_m, voc) + 1e-6 < _cosine_abs(inst, voc) - 0.01:
        blend = min(max(float(residual_blend), 0.0), 1.0)
        refined = (1.0 - blend) * inst + blend * res_m
        peak = float(refined.abs().max())
        if peak > 0.99:
            refined = refined * (0.99 / peak)
        out = refined
    if float(out.abs().max()) < 1e-6:
        peak = float(resid.abs().max())
        out = resid / peak if peak > 1.0 else resid
    return out


def separate_music(input_dict: Dict[str, List[str]], callback: Optional[Callable] = None,
                   separator: Optional[Separator] = None, **kwargs) -> List[str]:
    """{output_folder: [input paths]} -> list of written stem paths.  ``callback(fraction, desc, total)``."""
    for key, off in _OUT_OF_SCOPE.items():
        if kwargs.get(key, off) not in (off, None):
            raise NotImplementedError(f"{key}={kwargs[key]!r} needs a model family outside this engine's scope")
    for key in ("reverb_removal", "echo_removal", "crowd_removal", "noise_removal"):
        if kwargs.get(key, "Nothing") not in TRANSFORM_SETTINGS:
            raise ValueError(f"{key}={kwargs[key]!r}: one of {TRANSFORM_SETTINGS}")
    vocals_only = bool(kwargs.get("vocals_only", True))
    strength = int(kwargs.get("ensemble_strength", 2))
    models = ensemble_models(strength)
    residual_blend = float(kwargs.get("residual_blend", 0.4))
    if strength <= 2:                       # stem_separator.py:388-390
        residual_blend = min(residual_blend, 0.2)
    pcm16 = bool(kwargs.get("pcm16_handoff", False))
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
    # progress accounting of predict_with_model (:885-888: reverb, crowd and noise removal count, echo removal does not)
    trans_opts = [kwargs.get(k, "Nothing") for k in ("reverb_removal", "crowd_removal", "noise_removal")]
    transform_steps = (sum(o in ("All", "All Vocals", "Main Vocals") for o in trans_opts) + sum(o == "All" for o in trans_opts)) * len(files)
    total_steps = len(models) * len(files) + transform_steps + (0 if vocals_only else len(files)) + 1 + len(files)
    step = 0

    def advance(desc: str):
        nonlocal step
        step += 1
        if callback is not None:
            callback(min(step / total_steps, 1.0), desc, total_steps)

    if callback is not None:
        callback(0, "Starting ensemble separation...", total_steps)
    mixes = []
    for _, p in files:
        audio, sr = read_wav(p)
        mixes.append(sep.prepare_mix(torch.from_numpy(audio), sr))
    # what each model sees: the reference hands the mix over as a PCM_16 temp WAV (write_temp_wav)
    model_inputs = [pcm16_roundtrip(m) for m in mixes] if pcm16 else mixes
    results = [dict(vocals_list=[], instrumental_list=[]) for _ in files]
    wv, wi = [], []
    for name, w_voc, w_inst in models:
        sep.load_model(name)
        wv.append(w_voc)
        wi.append(w_inst)
        for i, mix in enumerate(model_inputs):
            stems = sep.separate_tensor(mix)
            results[i]["vocals_list"].append(stems["Vocals"])
            results[i]["instrumental_list"].append(stems["Instrumental"])
            advance(f"[Ensemble] {os.path.basename(files[i][1])} => {name}")
    for i, res in enumerate(results):
        res["vocals"] = blend_tracks(res.pop("vocals_list"), wv)
        res["instrumental"] = blend_tracks(res.pop("instrumental_list"), wi)
        res["instrumental"] = debleed_instrumental(mixes[i], res["vocals"], res["instrumental"], sep.sample_rate,
                                                   residual_blend)
    # transform chain (:903-934; the background-vocal split between its two halves is out of scope): reverb removal runs the
    # WHOLE chain on the vocals first; with crowd or noise removal set the chain runs again on the vocals without its reverb
    # step, and on the instrumental
    if kwargs.get("reverb_removal", "Nothing") != "Nothing":
        for i, res in enumerate(results):
            base = os.path.basename(files[i][1])
            ir_path = None
            if kwargs.get("store_reverb_ir", False):
                os.makedirs(files[i][0], exist_ok=True)
                ir_path = os.path.join(files[i][0], "impulse_response.ir")
            res["vocals"] = apply_transform_chain(sep, res["vocals"], "vocals", kwargs, pcm16=pcm16, ir_path=ir_path,
                                                  on_step=lambda d, b=base: advance(f"{d} for {b}"))
    if any(kwargs.get(k, "Nothing") != "Nothing" for k in ("crowd_removal", "noise_removal")):
        for i, res in enumerate(results):
            base = os.path.basename(files[i][1])
            res["vocals"] = apply_transform_chain(sep, res["vocals"], "vocals", kwargs, skip_transforms=("No Reverb",),
                                                  pcm16=pcm16, on_step=lambda d, b=base: advance(f"{d} for {b}"))
            res["instrumental"] = apply_transform_chain(sep, res["instrumental"], "instrumental", kwargs, pcm16=pcm16,
                                                        on_step=lambda d, b=base: advance(f"{d} for {b}"))
    if not vocals_only:
        # 6-stem stage on the full mix (stem_separator.py:459-503); vocals / instrumental stay the ensemble's
        sep.load_model("htdemucs_6s.yaml")
        for i, mix in enumerate(model_inputs):
            stems = sep.separate_tensor(mix)
            for key in ("drums", "bass", "guitar", "piano", "other"):
                results[i][key] = stems[key.capitalize()]
            advance(f"6-stem separation completed for {os.path.basename(files[i][1])}.")
    advance("Saving all stems...")
    outputs: List[str] = []
    for (out_folder, p), res in zip(files, results):
        os.makedirs(out_folder, exist_ok=True)
        base = os.path.splitext(os.path.basename(p))[0]
        for key, label in STEM_LABELS.items():
            wav = res.get(key)
            if wav is None or wav.numel() == 0 or float(wav.abs().max()) < 1e-6:     # silent stems are not written
                continue
            path = os.path.join(out_folder, f"{base}__{label}.wav")
            write_wav(path, wav.cpu().numpy(), sep.sample_rate, subtype="FLOAT")
            outputs.append(path)
        advance(f"Stems saved for {base}.")
    if callback is not None:
        callback(1.0, "Separation complete", total_steps)
    return outputs
