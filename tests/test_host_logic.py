"""CPU tests: host-side logic, product nets vs oracle nets, the C ABI's exported symbols."""
import ctypes
import dataclasses
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import htdemucs as oht
from oracle import roformer as oro
from oracle.metrics import max_abs_err
from oracle.synth import synth_mix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from audiolab_b200 import _lib
    from audiolab_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "audiolab_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(al_\w+)\s*\(", header, flags=re.M))
    assert {"al_stft", "al_istft", "al_ola_gather", "al_resample_poly", "al_plan_create"} <= declared
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    handle.al_version.restype = ctypes.c_int
    assert handle.al_version() >= 100
    assert set(_lib.SIGNATURES) - _lib.OPTIONAL <= declared | {"al_gemm_bf16"}


def test_no_cpu_fallback():
    from audiolab_b200 import demix, spectral
    from audiolab_b200.configs import RoformerConfig
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        spectral.resample_poly(torch.zeros(2, 100))
    with pytest.raises(RuntimeError):
        from audiolab_b200.separator import Separator
        Separator()


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import audiolab_b200, audiolab_b200.demix, audiolab_b200.separator, "
            "audiolab_b200.wrappers, audiolab_b200.sharding, audiolab_b200.nets.roformer, "
            "audiolab_b200.nets.htdemucs, audiolab_b200.nets.tfc_tdf; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "audiolab_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


@pytest.mark.parametrize("kind", ["bs", "mel"])
def test_product_roformer_net_loads_oracle_state_dict_and_matches(kind):
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.nets.roformer import RoformerMaskNet
    oc = oro.RoformerConfig(kind=kind, dim=32, depth=2, heads=2, dim_head=16, chunk_size=441 * 40, num_stems=2)
    om = oro.build_roformer(oc)
    pm = RoformerMaskNet(RoformerConfig(**dataclasses.asdict(oc))).eval()
    pm.load_state_dict(om.state_dict(), strict=True)
    x = torch.tensor(synth_mix(441 * 40, seed=2))[None]
    spec, _ = om._stft(x)
    mref = om.mask_from_spec(spec)
    b, fs, t, _ = spec.shape
    sp_il = torch.view_as_complex(spec.contiguous()).reshape(b, fs // 2, 2, t).permute(0, 3, 1, 2).contiguous()
    got = pm.mask(sp_il).permute(0, 1, 3, 4, 2).reshape(b, 2, fs, t)
    assert max_abs_err(got, mref) < 1e-6


def test_schedules_match_oracle():
    from audiolab_b200.demix import hamming_sym, roformer_schedule, triangle_weight
    for n, C, step in [(100, 40, 10), (352800 * 3 + 17, 352800, 88200), (40, 40, 10), (95, 40, 10)]:
        sched = oro.chunk_schedule(n, C, step)
        offs, mult = roformer_schedule(n, C, step)
        assert list(zip(offs, mult)) == sched
    assert np.array_equal(hamming_sym(1000), oro.hamming_sym(1000).astype(np.float32))
    from scipy.signal import windows
    assert np.allclose(oro.hamming_sym(999), windows.hamming(999), atol=1e-15)
    assert np.array_equal(triangle_weight(1001), oht.triangle_weight(1001).numpy())


def test_mel_membership_matches_oracle():
    from audiolab_b200.nets.roformer import mel_band_membership
    assert np.array_equal(mel_band_membership(44100, 2048, 60), oro.mel_band_masks(44100, 2048, 60))


def test_wav_roundtrip(tmp_path):
    from audiolab_b200.wavio import read_wav, write_wav
    x = synth_mix(5000, seed=1)
    write_wav(str(tmp_path / "f.wav"), x, 44100, "FLOAT")
    y, sr = read_wav(str(tmp_path / "f.wav"))
    assert sr == 44100 and np.array_equal(x, y)
    write_wav(str(tmp_path / "p.wav"), x, 48000, "PCM_16")
    y, sr = read_wav(str(tmp_path / "p.wav"))
    assert sr == 48000 and np.abs(x - y).max() <= 1.0 / 32768 + 1e-7


def test_blend_tracks_matches_reference_formula():
    from audiolab_b200.orchestrator import blend_tracks
    a, b = synth_mix(1000, seed=1), synth_mix(900, seed=2)
    got = blend_tracks([torch.tensor(a), torch.tensor(b)], [8.6, 8.4]).numpy()
    comb = np.zeros((2, 1000), np.float32)          # stem_separator.py:241-262
    comb[:, :1000] += a * 8.6
    comb[:, :900] += b * 8.4
    comb = comb / max(8.6 + 8.4, 1e-6)
    comb /= np.max(np.abs(comb))
    assert np.allclose(got, comb, atol=1e-6)


def test_separate_wrapper_surface_and_cache(tmp_path, monkeypatch):
    """Same attributes / kwargs as wrappers/separate.py:22-138, cache + special-file behaviour :233-388."""
    from audiolab_b200 import project_files
    from audiolab_b200.wrappers import BaseWrapper, Separate
    from audiolab_b200.wrappers import separate as sep_mod
    from audiolab_b200.wavio import write_wav
    monkeypatch.setattr(project_files, "output_path", str(tmp_path / "outputs"))
    w = Separate()
    assert isinstance(w, BaseWrapper) and Separate() is w
    assert (w.title, w.priority, w.default, w.required) == ("Separate", 1, True, False)
    assert set(w.allowed_kwargs) == {
        "delete_extra_stems", "separate_bg_vocals", "bg_vocal_layers", "vocals_only", "store_reverb_ir",
        "separate_drums", "separate_woodwinds", "alt_bass_model", "reverb_removal", "echo_removal",
        "crowd_removal", "noise_removal", "noise_removal_model", "delay_removal_model", "crowd_removal_model"}
    assert w.allowed_kwargs["vocals_only"].default is True
    src = tmp_path / "song.wav"
    write_wav(str(src), synth_mix(4000, seed=3), 44100, "PCM_16")
    tts = tmp_path / "TTS_hello.wav"
    write_wav(str(tts), synth_mix(1000, seed=4), 44100, "PCM_16")
    calls = []

    def fake_separate_music(input_dict, callback=None, **kwargs):
        calls.append((input_dict, kwargs))
        out = []
        for folder, files in input_dict.items():
            for f in files:
                base = os.path.splitext(os.path.basename(f))[0]
                for tag in ("(Vocals)", "(Instrumental)"):
                    p = os.path.join(folder, f"{base}_{tag}.wav")
                    write_wav(p, synth_mix(100, seed=5), 44100)
                    out.append(p)
                open(os.path.join(folder, "tmp_extra.wav"), "wb").write(b"x")
        if callback is not None:
            callback(1.0, "done", 3)
        return out

    monkeypatch.setattr(sep_mod, "separate_music", fake_separate_music)
    seen = []
    projects = [project_files.ProjectFiles(str(src)), project_files.ProjectFiles(str(tts))]
    res = w.process_audio(projects, callback=lambda f, d, t: seen.append((f, d, t)), vocals_only=True, bogus=1)
    assert len(res) == 2 and len(calls) == 1 and seen == [(1.0, "done", 3)]
    assert "bogus" not in calls[0][1]
    by_name = {os.path.basename(p.src_file): p for p in res}
    stems = by_name["song.wav"].last_outputs
    assert sorted(os.path.basename(s) for s in stems) == ["song_(Instrumental).wav", "song_(Vocals).wav"]
    stem_dir = os.path.dirname(stems[0])
    assert not os.path.exists(os.path.join(stem_dir, "tmp_extra.wav"))          # delete_extra_stems
    info = json.load(open(os.path.join(stem_dir, "separation_info.json")))
    assert info["config"]["vocals_only"] is True and len(info["stems"]) == 2
    assert by_name["TTS_hello.wav"].last_outputs[0].endswith("TTS_hello(Vocals).wav")
    # second run: cache hit -> no new separate_music call
    res2 = w.process_audio([project_files.ProjectFiles(str(src))], vocals_only=True)
    assert len(calls) == 1 and sorted(res2[0].last_outputs) == sorted(stems)
    # changed config -> cache miss
    w.process_audio([project_files.ProjectFiles(str(src))], vocals_only=True, bg_vocal_layers=2)
    assert len(calls) == 2
    # tampered stem -> cache miss
    open(stems[0], "ab").write(b"0")
    w.process_audio([project_files.ProjectFiles(str(src))], vocals_only=True, bg_vocal_layers=2)
    assert len(calls) == 3


def test_orchestrator_rejects_out_of_scope_options():
    from audiolab_b200.orchestrator import separate_music
    with pytest.raises(NotImplementedError):
        separate_music({"/tmp/x": []}, separate_bg_vocals=True)
    with pytest.raises(ValueError):
        separate_music({"/tmp/x": ["/nonexistent.wav"]}, reverb_removal="Everything")
    assert separate_music({"/tmp/x": ["/nonexistent.wav"]}, reverb_removal="All") == []   # in scope since the transform chain
    with pytest.raises(NotImplementedError):          # MDX23C is the 4th model of the reference's list
        separate_music({"/tmp/x": []}, ensemble_strength=4)
    assert separate_music({"/tmp/x": ["/nonexistent.wav"]}) == []
    assert separate_music({"/tmp/x": ["/nonexistent.wav"]}, vocals_only=False) == []


def test_separator_arch_resolution():
    from audiolab_b200.separator import _arch_of
    assert _arch_of("model_bs_roformer_ep_368_sdr_12.9628.ckpt") == "bs_roformer"
    assert _arch_of("vocals_mel_band_roformer.ckpt") == "mel_roformer"
    assert _arch_of("melband_roformer_big_beta4.ckpt") == "mel_roformer"
    assert _arch_of("UVR-MDX-NET-Voc_FT.onnx") == "mdx"
    assert _arch_of("htdemucs_6s.yaml") == "htdemucs"
    assert _arch_of("17_HP-Wind_Inst-UVR.pth") == "vr"


class _TorchPlan:
    """CPU stand-in for spectral.StftPlan with al_stft / al_istft semantics (chunk gather with zeros outside the track,
    reflection about the chunk ends by `center_pad`, FRAME_MAJOR layout) on torch.fft -- host-logic tests only."""

    def __init__(self, n_fft, hop):
        self.n_fft, self.hop, self.win = n_fft, hop, torch.hann_window(n_fft)

    def stft(self, track, *, chunk_len, n_chunks=1, off0=0, off_step=0, n_valid=None, center_pad=None, n_frames=None,
             layout=0, **_):
        assert layout == 0
        ch, n = track.shape
        n_valid = n if n_valid is None else n_valid
        center_pad = self.n_fft // 2 if center_pad is None else center_pad
        out = []
        for c in range(n_chunks):
            idx = torch.arange(off0 + c * off_step, off0 + c * off_step + chunk_len)
            ok = (idx >= 0) & (idx < n_valid)
            chunk = torch.zeros(ch, chunk_len)
            chunk[:, ok] = track[:, idx[ok]]
            j = torch.arange(-center_pad, (n_frames - 1) * self.hop - center_pad + self.n_fft)
            j = torch.where(j < 0, -j, j)
            j = torch.where(j >= chunk_len, 2 * (chunk_len - 1) - j, j)
            S = torch.stft(chunk[:, j], self.n_fft, self.hop, window=self.win, center=False, return_complex=True)
            out.append(S.transpose(1, 2))                                   # [ch, T, F]
        return torch.cat(out, 0).contiguous()

    def istft(self, spec, *, n_chunks, channels, layout=0, **_):
        assert layout == 0 and n_chunks == 1
        w = torch.istft(spec.transpose(1, 2), self.n_fft, self.hop, window=self.win, center=True)
        return w[None, None]                                                # [1, 1, ch, hop * (T - 1)]


def test_mdx_secondary_stem_by_spectral_inversion_host_logic(monkeypatch):
    """MdxDemixer.invert_stem (device path: two al_stft with zero centre padding, the in-tree
    "invert_p" arithmetic, al_istft) against the oracle's torch / librosa-convention restatement -- the kernels are
    replaced by their torch definitions, so this pins the host logic (padding convention, layout, length, sign)."""
    import audiolab_b200.demix as demix
    from oracle import mdx as omdx
    monkeypatch.setattr(demix, "_check_mix", lambda m, channels=2: m.contiguous().float())
    monkeypatch.setattr(demix.MdxDemixer, "_inv_plan", _TorchPlan(2048, 1024))
    d = object.__new__(demix.MdxDemixer)
    for n in (44100, 1024 * 7, 5000):
        mix = torch.tensor(synth_mix(n, seed=n))
        stem = 0.6 * mix.flip(0) + 0.1 * torch.tensor(synth_mix(n, seed=n + 1))
        got = d.invert_stem(mix, stem)
        ref = omdx.invert_stem(mix.numpy(), stem.numpy())
        m = ref.shape[1]
        assert got.shape == (2, n) and m == 1024 * (n // 1024)
        assert float((got[:, :m] - torch.tensor(ref)).abs().max()) <= 2e-5
        assert not got[:, m:].any()
    # a stem equal to the mix inverts to (almost) silence; a silent stem inverts to minus-minus the mix = the mix's STFT round trip
    mix = torch.tensor(synth_mix(20480, seed=3))
    assert float(d.invert_stem(mix, mix).abs().max()) <= 1e-5
    back = d.invert_stem(mix, torch.zeros_like(mix))
    assert float((back[:, 1024:-1024] - mix[:, 1024:-1024]).abs().max()) <= 1e-4


def test_gemm_args_struct_layout_matches_the_header(tmp_path):
    """The ctypes mirror of `al_gemm_args` (netops.GemmArgs) against the C header, field by field (gcc, no GPU)."""
    import ctypes
    import shutil
    import subprocess

    from audiolab_b200.netops import GemmArgs
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    fields = [f[0] for f in GemmArgs._fields_]
    src = tmp_path / "layout.c"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{root}/include/audiolab_b200.h"', 'int main(void) {',
             '  printf("%zu\\n", sizeof(al_gemm_args));']
    lines += [f'  printf("%zu\\n", offsetof(al_gemm_args, {f}));' for f in fields]
    lines += ['  return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(GemmArgs)
    assert out[1:] == [getattr(GemmArgs, f).offset for f in fields]
