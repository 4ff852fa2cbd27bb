"""``Separator`` -- duck type of ``audio_separator.separator.Separator`` as AudioLab uses it.

Callee surface the reference's orchestrator expects (call sites
/root/reference/modules/separator/stem_separator.py:102-107 ctor, :124 download_model_files,
:394 load_model, :281-282 separate -> file names relative to output_dir, :399-400 output_dir /
model_instance.output_dir; /root/reference/handlers/reverb.py:382 ``output_dir=`` ctor kwarg).
Output names carry ``(Vocals)`` / ``(Instrumental)`` (stem_separator.py:325-329) or the 4/6 stem
names (:491-500).

Underneath there is an in-memory path ``separate_tensor`` (device tensors in and out), which is what
bench.py and the orchestrator use; ``separate(path)`` wraps it with WAV I/O.
There is no CPU path: constructing a Separator without a CUDA device raises.
"""
from __future__ import annotations

import logging
import os
from dataclasses import replace
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import spectral as sp
from .configs import HTDemucsConfig, MdxConfig, RoformerConfig
from .demix import HTDemucsDemixer, MdxDemixer, RoformerDemixer
from .wavio import read_wav, write_wav

# (primary, secondary) stem names of the single-target models the orchestrator's transform chain loads; the reference picks
# the output to keep by looking for its label ("No Reverb", "dry", "No Crowd") in the file names audio-separator derives
# from these (stem_separator.py:795-822, debug_reverb :1046-1056 lists the dry / wet strings of the de-reverb models)
STEMS_OF = {
    "dereverb_mel_band_roformer_anvuew_sdr_19.1729.ckpt": ("noreverb", "reverb"),
    "dereverb_mel_band_roformer_less_aggressive_anvuew_sdr_18.8050.ckpt": ("noreverb", "reverb"),
    "dereverb-echo_mel_band_roformer_sdr_10.0169.ckpt": ("dry", "No dry"),
    "dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt": ("dry", "No dry"),
    "UVR-MDX-NET_Crowd_HQ_1.onnx": ("No Crowd", "Crowd"),
}

DEMUCS_4 = ["Drums", "Bass", "Other", "Vocals"]
DEMUCS_6 = ["Drums", "Bass", "Other", "Vocals", "Guitar", "Piano"]


def _arch_of(model_filename: str) -> str:
    name = model_filename.lower()
    if name.endswith(".onnx"):
        return "mdx"
    if "roformer" in name:
        return "mel_roformer" if ("mel_band" in name or "melband" in name) else "bs_roformer"
    if name.startswith("htdemucs") or name.startswith("hdemucs"):
        return "htdemucs"
    if name.endswith(".pth"):
        return "vr"
    if "mdx23c" in name:
        return "mdx23c"
    raise ValueError(f"cannot infer the architecture of {model_filename!r}")


class _ModelInstance:
    """What ``separator.model_instance`` exposes to the orchestrator and to patch_separate.py."""

    def __init__(self, arch: str, model_path: str, device, logger, output_dir):
        self.arch = arch
        self.model_path = model_path
        self.torch_device = device
        self.logger = logger
        self.output_dir = output_dir
        self.model_run: Optional[Callable] = None
        self.segment_size = None
        self.dim_t = None
        self.demixer = None
        self.primary_stem = "Vocals"
        self.secondary_stem = "Instrumental"
        self.stem_names: List[str] = []


class Separator:
    def __init__(self, log_level=logging.INFO, model_file_dir: str = "/tmp/audio-separator-models/",
                 output_dir: Optional[str] = None, output_format: str = "WAV", normalization_threshold: float = 0.9,
                 invert_using_spec: bool = False, sample_rate: int = 44100, use_autocast: bool = False,
                 mdx_params: Optional[dict] = None, mdxc_params: Optional[dict] = None,
                 demucs_params: Optional[dict] = None, allow_random_init: bool = False, device: str = "cuda:0",
                 model_overrides: Optional[dict] = None):
        self.logger = logging.getLogger(__name__)
        self.logger.setLevel(log_level)
        if not torch.cuda.is_available():
            raise RuntimeError("audiolab_b200.Separator needs a CUDA device (B200); there is no CPU fallback")
        self.torch_device = torch.device(device)
        self.model_file_dir = model_file_dir
        self.output_dir = output_dir or os.getcwd()
        self.output_format = output_format
        self.normalization_threshold = normalization_threshold
        self.invert_using_spec = invert_using_spec
        self.sample_rate = sample_rate
        self.use_autocast = use_autocast
        self.mdx_params = {"hop_length": 1024, "segment_size": 256, "overlap": 0.25, "batch_size": 1,
                           "enable_denoise": False, **(mdx_params or {})}
        self.mdxc_params = {"segment_size": 256, "override_model_segment_size": False, "batch_size": 4,
                            "overlap": 4, "pitch_shift": 0, **(mdxc_params or {})}
        self.demucs_params = {"segment_size": "Default", "shifts": 2, "overlap": 0.25, "segments_enabled": True,
                              **(demucs_params or {})}
        self.allow_random_init = allow_random_init or os.environ.get("AUDIOLAB_B200_RANDOM_INIT") == "1"
        self.model_overrides = model_overrides or {}
        self.model_instance: Optional[_ModelInstance] = None
        self.model_name: Optional[str] = None
        self._resample_taps = {}

    # ------------------------------------------------------------------------------------------
    def download_model_files(self, model_filename: str):
        """The reference downloads weights here (stem_separator.py:123-124).  No network in this build:
        report whether the file is already under model_file_dir."""
        path = os.path.join(self.model_file_dir, model_filename)
        if not os.path.exists(path):
            self.logger.debug("model file %s not present (no download in this environment)", path)
        return path

    def _net_dtype(self) -> torch.dtype:
        """16-bit operand format of the mask network under use_autocast.  Default float16: the same tensor-core rate as
        bfloat16 with 11 instead of 8 significand bits, which is what meets BASELINE.json's SI-SDR >= 60 dB against the fp32
        reference (63.6 dB at the real size; bfloat16 operands give 44.7 dB, profiles/r02i_parity_fullsize_fp16.jsonl) --
        and what upstream's `torch.autocast("cuda")` defaults to.  `mdxc_params={"compute_dtype": "bf16"}` or
        AUDIOLAB_B200_NET_DTYPE=bf16 selects bfloat16 (fp32's exponent range; float16 outputs saturate at +-65504)."""
        name = str(self.mdxc_params.get("compute_dtype", os.environ.get("AUDIOLAB_B200_NET_DTYPE", "fp16"))).lower()
        if name in ("fp16", "float16", "half"):
            return torch.float16
        if name in ("bf16", "bfloat16"):
            return torch.bfloat16
        raise ValueError(f"compute_dtype {name!r}: expected 'bf16' or 'fp16'")

    def _state_dict_or_none(self, path: str):
        if os.path.exists(path) and not path.endswith((".onnx", ".yaml")):
            sd = torch.load(path, map_location="cpu", weights_only=True)
            return sd.get("state_dict", sd) if isinstance(sd, dict) else sd
        if not self.allow_random_init:
            raise FileNotFoundError(
                f"{path} not found and allow_random_init is False (set allow_random_init=True or "
                "AUDIOLAB_B200_RANDOM_INIT=1 to run with seeded random weights)")
        return None

    def _demucs_state_dict_or_none(self, path: str):
        """Weights of an HTDemucs model name.  Upstream names are bag-of-models ``.yaml`` files
        (``models: [<signature>, ...]``, optional ``weights:``) next to ``<signature>.th`` packages.  A single-model
        bag (htdemucs_6s, htdemucs) is resolved to its ``.th`` file; a real bag (htdemucs_ft: four fine-tuned models,
        one per source) is NOT averaged here -- it raises instead of silently running something else.  A missing file
        falls back to seeded random weights only with allow_random_init."""
        if path.endswith(".yaml") and os.path.exists(path):
            import yaml
            with open(path) as fh:
                bag = yaml.safe_load(fh) or {}
            sigs = list(bag.get("models") or [])
            if len(sigs) != 1:
                raise NotImplementedError(
                    f"{path}: bag of {len(sigs)} models (per-source weighted average, demucs.apply.BagOfModels) is not "
                    "implemented; use a single-model bag (htdemucs.yaml, htdemucs_6s.yaml)")
            th = os.path.join(os.path.dirname(path), f"{sigs[0]}.th")
            if not os.path.exists(th):
                matches = [f for f in os.listdir(os.path.dirname(path)) if f.startswith(str(sigs[0])) and f.endswith(".th")]
                if not matches:
                    raise FileNotFoundError(f"{path} names model {sigs[0]!r} but no {sigs[0]}*.th file is beside it")
                th = os.path.join(os.path.dirname(path), matches[0])
            path = th
        if os.path.exists(path) and not path.endswith(".yaml"):
            pkg = torch.load(path, map_location="cpu", weights_only=True)
            if isinstance(pkg, dict) and "state" in pkg:          # demucs.states package: {klass, args, kwargs, state}
                pkg = pkg["state"]
            return pkg.get("state_dict", pkg) if isinstance(pkg, dict) else pkg
        if not self.allow_random_init:
            raise FileNotFoundError(
                f"{path} not found and allow_random_init is False (set allow_random_init=True or "
                "AUDIOLAB_B200_RANDOM_INIT=1 to run with seeded random weights)")
        return None

    def load_model(self, model_filename: str = "model_bs_roformer_ep_368_sdr_12.9628.ckpt"):
        arch = _arch_of(model_filename)
        path = os.path.join(self.model_file_dir, model_filename)
        inst = _ModelInstance(arch, path, self.torch_device, self.logger, self.output_dir)
        ov = dict(self.model_overrides.get(model_filename, self.model_overrides.get(arch, {})))
        net = ov.pop("net", None)
        if arch in ("bs_roformer", "mel_roformer"):
            from .nets.roformer import RoformerMaskNet
            cfg = RoformerConfig(kind="mel" if arch == "mel_roformer" else "bs",
                                 num_overlap=int(self.mdxc_params["overlap"]))
            cfg = replace(cfg, **ov)
            if net is None:
                sd = self._state_dict_or_none(path)
                torch.manual_seed(4321)
                net = RoformerMaskNet(cfg)
                if sd is not None:
                    net.load_state_dict(sd, strict=True)
            net = net.to(self.torch_device).eval()
            net.set_compute_dtype(self._net_dtype() if self.use_autocast else torch.float32)
            inst.demixer = RoformerDemixer(cfg, net, batch_size=int(self.mdxc_params["batch_size"]))
            inst.stem_names = ["Vocals"] if cfg.num_stems == 1 else [f"Stem{i}" for i in range(cfg.num_stems)]
            inst.run = lambda mix: inst.demixer.demix(mix)
        elif arch == "mdx":
            from .nets.tfc_tdf import TfcTdfNet
            seg = int(self.mdx_params["segment_size"])
            cfg = MdxConfig(hop=int(self.mdx_params["hop_length"]), dim_t_log2=int(np.log2(seg)),
                            overlap=float(self.mdx_params["overlap"]), denoise=bool(self.mdx_params["enable_denoise"]),
                            zero_low_bins=3)
            cfg = replace(cfg, **ov)
            if net is None and os.path.isfile(path):
                # the released .onnx itself, as a device-resident torch module (nets/onnx_graph.py: own protobuf reader + graph
                # executor; the reference's onnx2torch branch, handlers/patch_separate.py:54-63, without the onnx packages)
                from .nets.onnx_graph import OnnxGraphNet
                net = OnnxGraphNet.from_file(path)
                if net.dim_f:
                    cfg = replace(cfg, dim_f=int(net.dim_f))
                if net.dim_t and net.dim_t != cfg.dim_t:
                    if net.dim_t & (net.dim_t - 1):
                        raise ValueError(f"{path}: model dim_t {net.dim_t} is not a power of two")
                    self.logger.warning(f"{model_filename}: segment_size {seg} != the model's dim_t {net.dim_t}; using the model's")
                    seg = int(net.dim_t)
                    cfg = replace(cfg, dim_t_log2=int(np.log2(seg)))
            if net is None:
                # Without the .onnx: a released model as a converted state dict `<name>.pt` / `.pth` next to the `.onnx` name
                # (KUIELab module names, nets/tfc_tdf.py::ConvTdfNet); otherwise seeded random weights of the released shape.
                from .nets.tfc_tdf import ConvTdfNet
                stem = os.path.splitext(path)[0]
                converted = next((stem + ext for ext in (".pt", ".pth", ".ckpt") if os.path.exists(stem + ext)), None)
                if converted is not None:
                    sd = torch.load(converted, map_location="cpu", weights_only=True)
                    net = ConvTdfNet.from_state_dict(sd.get("state_dict", sd) if isinstance(sd, dict) else sd, cfg.dim_f)
                elif not self.allow_random_init:
                    raise FileNotFoundError(f"{path}: neither the .onnx file nor a converted state dict "
                                            f"({stem}.pt); pass model_overrides[...]['net'] or allow_random_init=True")
                else:
                    torch.manual_seed(4321)
                    full = self.mdx_params.get("full_size_net", True)
                    if full and cfg.dim_t % 32 != 0:
                        raise ValueError(f"segment_size {cfg.dim_t}: the released (L = 11) TFC-TDF net halves the time axis 5 times; "
                                         "use a multiple of 32 (default 256), or mdx_params['full_size_net'] = False")
                    net = ConvTdfNet(cfg.dim_f) if full else TfcTdfNet(cfg.dim_f)
            net = net.to(self.torch_device).eval()
            inst.segment_size, inst.dim_t = seg, cfg.dim_t
            autocast = self.use_autocast

            def model_run(spek, _net=net):
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    return _net(spek)

            inst.model_run = model_run
            inst.demixer = MdxDemixer(cfg, lambda s: inst.model_run(s), batch_size=max(1, int(self.mdx_params["batch_size"])))
            inst.stem_names = ["Vocals"]
            inst.run = lambda mix: inst.demixer.demix_windowed(mix)[None]
        elif arch == "htdemucs":
            from .nets.htdemucs import HTDemucsCore
            six = "6s" in model_filename
            shifts = int(self.demucs_params["shifts"])
            cfg = HTDemucsConfig(num_sources=6 if six else 4, shifts=shifts, overlap=float(self.demucs_params["overlap"]))
            cfg = replace(cfg, **ov)
            if net is None:
                sd = self._demucs_state_dict_or_none(path)
                torch.manual_seed(4321)
                net = HTDemucsCore(num_sources=cfg.num_sources)
                if sd is not None:
                    net.load_state_dict(sd, strict=True)
            net = net.to(self.torch_device).eval()
            autocast = self.use_autocast

            def core(x, xt, _net=net):
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    return _net(x, xt)

            inst.demixer = HTDemucsDemixer(cfg, core)
            inst.stem_names = DEMUCS_6 if six else DEMUCS_4
            inst.run = lambda mix: inst.demixer.demix(mix)
        else:
            raise NotImplementedError(
                f"{model_filename}: the {arch} architecture is outside this engine's scope "
                "(north star: MDX-Net, BS/Mel-RoFormer, HTDemucs)")
        if model_filename in STEMS_OF and len(inst.stem_names) == 1:
            inst.primary_stem, inst.secondary_stem = STEMS_OF[model_filename]
            inst.stem_names = [inst.primary_stem]
        self.model_instance = inst
        self.model_name = os.path.splitext(model_filename)[0]
        return inst

    # ------------------------------------------------------------------------------------------
    def _taps(self, up: int, down: int) -> torch.Tensor:
        key = (up, down)
        if key not in self._resample_taps:
            self._resample_taps[key] = torch.from_numpy(sp.resample_taps(up, down)).to(self.torch_device)
        return self._resample_taps[key]

    def prepare_mix(self, audio: torch.Tensor, sr: int) -> torch.Tensor:
        """[channels, n] at `sr` -> stereo fp32 device tensor at self.sample_rate (K3 polyphase resampler)."""
        x = audio.to(self.torch_device, dtype=torch.float32, non_blocking=True)
        if x.dim() == 1:
            x = x[None]
        if x.shape[0] == 1:
            x = x.repeat(2, 1)
        x = x[:2].contiguous()
        if sr != self.sample_rate:
            g = int(np.gcd(self.sample_rate, sr))
            up, down = self.sample_rate // g, sr // g
            x = sp.resample_poly(x, up, down, taps=self._taps(up, down))
        return x

    @torch.no_grad()
    def separate_tensor(self, mix: torch.Tensor, sr: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """In-memory fast path: mix [2, n] -> {stem name: [2, n]} on the device."""
        if self.model_instance is None:
            raise RuntimeError("load_model() first")
        inst = self.model_instance
        mix = self.prepare_mix(mix, sr or self.sample_rate)
        stems = inst.run(mix)                                  # [S, 2, n]
        out = {name: stems[i] for i, name in enumerate(inst.stem_names)}
        if len(inst.stem_names) == 1:
            if self.invert_using_spec and inst.arch == "mdx":
                # upstream MDXSeparator.separate: match-mix pass + spec_utils.invert_stem
                out[inst.secondary_stem] = inst.demixer.secondary_by_inversion(mix, stems[0])
            else:
                out[inst.secondary_stem] = sp.sub(mix, stems[0])   # `mix - primary`
        return out

    def separate(self, audio_file_path: str) -> List[str]:
        """WAV in -> WAV stems in output_dir; returns file names relative to output_dir
        (stem_separator.py:281-282)."""
        if self.model_instance is None:
            raise RuntimeError("load_model() first")
        audio, sr = read_wav(audio_file_path)
        x = torch.from_numpy(audio)
        peak = float(x.abs().max()) if x.numel() else 0.0
        if peak > self.normalization_threshold:                # upstream spec_utils.normalize
            x = x * (self.normalization_threshold / peak)
        stems = self.separate_tensor(x, sr)
        base = os.path.splitext(os.path.basename(audio_file_path))[0]
        out_dir = self.model_instance.output_dir or self.output_dir
        os.makedirs(out_dir, exist_ok=True)
        names = []
        for stem, wav in stems.items():
            fname = f"{base}_({stem})_{self.model_name}.wav"
            write_wav(os.path.join(out_dir, fname), wav.cpu().numpy(), self.sample_rate, subtype="FLOAT")
            names.append(fname)
        return names
