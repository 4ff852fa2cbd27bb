"""``Separate`` -- the wrapper AudioLab's pipeline calls for stem separation.

Drop-in for /root/reference/wrappers/separate.py: same class attributes (:22-25), same
``allowed_kwargs`` keys and defaults (:32-138), same ``process_audio(inputs, callback=None, **kwargs)``
signature (:233) and the same behaviour around the separation call: special-file skip (:247-272),
``separation_info.json`` cache keyed by config + SHA-256 of every stem (:274-315, 363-373), one
``separate_music`` call for all uncached projects (:337-341), extra-stem deletion (:376-386).
Only the engine underneath changes (audiolab_b200.orchestrator -> sm_100a kernels).
"""
from __future__ import annotations

import hashlib
import json
import logging
import os
import shutil
import threading
from typing import Any, Dict, List

from ..orchestrator import separate_music
from ..project_files import ProjectFiles
from .base_wrapper import BaseWrapper, TypedInput

logger = logging.getLogger(__name__)

_TARGETS = ["Nothing", "Main Vocals", "All Vocals", "All"]


def _flag(default: bool, description: str, **kw) -> TypedInput:
    return TypedInput(default=default, description=description, type=bool, gradio_type="Checkbox", **kw)


def _choice(default: str, description: str, choices: List[str]) -> TypedInput:
    return TypedInput(default=default, description=description, type=str, choices=choices, gradio_type="Dropdown")


class Separate(BaseWrapper):
    title = "Separate"
    priority = 1
    default = True
    required = False
    description = ("Separate audio into distinct stems with optional background vocal splitting "
                   "and audio transformations (reverb, echo, delay, crowd, noise removal).")
    file_operation_lock = threading.Lock()

    allowed_kwargs = {
        "delete_extra_stems": _flag(True, "Automatically delete intermediate stem files after processing."),
        "separate_bg_vocals": _flag(False, "Separate background vocals from main vocals."),
        "bg_vocal_layers": TypedInput(default=1, le=10, ge=1, type=int, gradio_type="Slider", render=False,
                                      description="Number of background vocal layers to separate."),
        "vocals_only": _flag(True, "Enable to separate only the main vocals and instrumental, disable for additional stems."),
        "store_reverb_ir": _flag(False, "Store the impulse response for reverb removal. Will be used to re-apply reverb later."),
        "separate_drums": _flag(False, "Separate the drum track."),
        "separate_woodwinds": _flag(False, "Separate the woodwind instruments."),
        "alt_bass_model": _flag(False, "Use an alternative bass model."),
        "reverb_removal": _choice("Nothing", "Apply reverb removal.", _TARGETS),
        "echo_removal": _choice("Nothing", "Apply echo/delay removal.", _TARGETS),
        "crowd_removal": _choice("Nothing", "Apply crowd noise removal.", _TARGETS),
        "noise_removal": _choice("Nothing", "Apply general noise removal.", _TARGETS),
        "noise_removal_model": _choice("UVR-DeNoise.pth", "Choose the model used for noise removal.",
                                       ["UVR-DeNoise.pth", "UVR-DeNoise-Lite.pth"]),
        "delay_removal_model": _choice("dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt",
                                       "Select the model for echo/delay removal.",
                                       ["dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt",
                                        "dereverb-echo_mel_band_roformer_sdr_10.0169.ckpt", "UVR-DeEcho-DeReverb.pth"]),
        "crowd_removal_model": _choice("UVR-MDX-NET_Crowd_HQ_1.onnx", "Select the model for crowd noise removal.",
                                       ["UVR-MDX-NET_Crowd_HQ_1.onnx",
                                        "mel_band_roformer_crowd_aufr33_viperx_sdr_8.7144.ckpt"]),
    }

    # keys of the cache config and the kwarg default each falls back to (separate.py:274-291)
    _CONFIG_DEFAULTS = {
        "vocals_only": True, "separate_drums": False, "separate_woodwinds": False, "alt_bass_model": False,
        "separate_bg_vocals": True, "bg_vocal_layers": 1, "reverb_removal": "Nothing", "echo_removal": "Nothing",
        "delay_removal": "Nothing", "crowd_removal": "Nothing", "noise_removal": "Nothing",
        "delay_removal_model": "dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt",
        "noise_removal_model": "UVR-DeNoise.pth", "crowd_removal_model": "UVR-MDX-NET_Crowd_HQ_1.onnx",
        "store_reverb_ir": True,
    }

    # engine hook: tests / benchmarks inject a configured Separator (e.g. random-init weights)
    engine_kwargs: Dict[str, Any] = {}

    def process_audio(self, inputs: List[ProjectFiles], callback=None, **kwargs: Dict[str, Any]) -> List[ProjectFiles]:
        filtered = {k: v for k, v in kwargs.items() if k in self.allowed_kwargs}
        final_projects: List[ProjectFiles] = []
        to_separate = []

        for project in inputs:
            project.base_name = os.path.splitext(os.path.basename(project.src_file))[0]
            out_dir = os.path.join(project.project_dir, "stems")
            os.makedirs(out_dir, exist_ok=True)
            cache_file = os.path.join(out_dir, "separation_info.json")

            fname, fdir = os.path.basename(project.src_file), os.path.dirname(project.src_file)
            if fname.startswith(("TTS_", "ZONOS_")) or any(d in fdir for d in ("tts", "zonos", "stable_audio")):
                stem_base, ext = os.path.splitext(fname)
                new_path = os.path.join(out_dir, f"{stem_base}(Vocals){ext}")
                if not os.path.exists(new_path):
                    shutil.copyfile(project.src_file, new_path)
                project.add_output("stems", [new_path])
                final_projects.append(project)
                logger.info("Skipping separation for special file %s", project.src_file)
                continue

            config = {"file": project.src_file}
            config.update({k: filtered.get(k, d) for k, d in self._CONFIG_DEFAULTS.items()})
            if self._cache_hit(cache_file, config, project):
                final_projects.append(project)
            else:
                to_separate.append((project, config))

        if to_separate:
            input_dict: Dict[str, List[str]] = {}
            by_dir = {}
            for proj, config in to_separate:
                stem_dir = os.path.join(proj.project_dir, "stems")
                input_dict.setdefault(stem_dir, []).append(proj.src_file)
                by_dir[os.path.normpath(stem_dir)] = (proj, config)
            stems = separate_music(input_dict=input_dict, callback=callback, **filtered, **self.engine_kwargs)
            results: Dict[str, List[str]] = {}
            for s in stems:
                results.setdefault(os.path.normpath(os.path.dirname(s)), []).append(s)
            for stem_dir, (proj, config) in by_dir.items():
                if stem_dir not in results:
                    logger.warning("No separation results found for project %s", proj.src_file)
                    continue
                proj.add_output("stems", results[stem_dir])
                final_projects.append(proj)
                info = {"config": config, "stems": [{"path": p, "hash": self._hash_file(p)} for p in results[stem_dir]]}
                try:
                    with open(os.path.join(stem_dir, "separation_info.json"), "w") as f:
                        json.dump(info, f, indent=2)
                except OSError as e:
                    logger.warning("Error writing cache file in %s: %s", stem_dir, e)

        if filtered.get("delete_extra_stems", True):
            for project in final_projects:
                out_dir = os.path.join(project.project_dir, "stems")
                keep = project.file_dict.get("stems", [])
                for fname in os.listdir(out_dir):
                    full = os.path.join(out_dir, fname)
                    if fname in ("separation_info.json", "impulse_response.ir"):
                        continue
                    if full not in keep:
                        self.del_stem(full)
        return final_projects

    def _cache_hit(self, cache_file: str, config: dict, project: ProjectFiles) -> bool:
        if not os.path.exists(cache_file):
            return False
        try:
            with open(cache_file, "r") as f:
                cached = json.load(f)
            if cached.get("config") != config:
                return False
            stems = []
            for info in cached.get("stems", []):
                path = info.get("path")
                if not os.path.exists(path) or self._hash_file(path) != info.get("hash"):
                    return False
                stems.append(path)
            project.add_output("stems", stems)
            return True
        except Exception as e:  # unreadable cache == no cache, like the reference (:313-314)
            logger.warning("Error reading cache file %s: %s", cache_file, e)
            return False

    def del_stem(self, path: str) -> bool:
        try:
            with self.file_operation_lock:
                if os.path.exists(path):
                    os.remove(path)
                    return True
        except OSError as e:
            logger.warning("Error deleting %s: %s", path, e)
        return False

    def _hash_file(self, filepath: str) -> str:
        h = hashlib.sha256()
        try:
            with open(filepath, "rb") as f:
                for chunk in iter(lambda: f.read(65536), b""):
                    h.update(chunk)
        except OSError as e:
            logger.warning("Error hashing file %s: %s", filepath, e)
        return h.hexdigest()
