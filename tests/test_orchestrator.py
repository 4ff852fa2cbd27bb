"""Orchestrator rows of SURVEY.md section 8 (a5, a16, f1, f2): ensemble selection, blend, residual de-bleed, PCM_16
hand-off and the 6-stem stage of audiolab_b200/orchestrator.py against the numpy restatement of the reference
(oracle/debleed.py, stem_separator.py:173-262, 414-456).  The functions are device-agnostic torch code: the CPU run
checks the arithmetic, the GPU run (test_orchestrator_gpu) the same on the device."""
import os

import numpy as np
import pytest
import torch

from audiolab_b200 import orchestrator as orch
from oracle import debleed as ref
from oracle.synth import synth_mix


def _case(seed, n=60000, lag=0, gain=0.8):
    rs = np.random.RandomState(seed)
    voc = synth_mix(n, seed=seed)
    inst = synth_mix(n, seed=seed + 100) * 0.7
    shifted = np.roll(voc, lag, axis=1)
    mix = (inst + gain * shifted + 0.001 * rs.randn(2, n)).astype(np.float32)
    return mix, voc.astype(np.float32), inst.astype(np.float32)


@pytest.mark.parametrize("lag,gain", [(0, 0.8), (37, 1.0), (-211, 0.5), (500, 2.0), (3, -0.5)])
def test_residual_subtract_matches_reference(lag, gain):
    mix, voc, _ = _case(11 + abs(lag), lag=lag, gain=gain)
    want = ref.residual_subtract(mix, voc, 44100)
    got = orch.residual_subtract(torch.from_numpy(mix), torch.from_numpy(voc), 44100).numpy()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-6 * max(1.0, scale)


def test_residual_subtract_ragged_lengths_and_mono():
    mix, voc, _ = _case(5, n=30000)
    want = ref.residual_subtract(mix, voc[:, :25000], 44100)
    got = orch.residual_subtract(torch.from_numpy(mix), torch.from_numpy(voc[:, :25000]), 44100).numpy()
    assert got.shape == mix.shape and np.abs(got - want).max() <= 2e-6
    want = ref.residual_subtract(mix[0], voc[0], 44100)
    got = orch.residual_subtract(torch.from_numpy(mix[0]), torch.from_numpy(voc[0]), 44100).numpy()
    assert got.shape == (2, 30000) and np.abs(got - want).max() <= 2e-6


@pytest.mark.parametrize("bleed", [0.0, 0.3])
def test_debleed_matches_reference(bleed):
    mix, voc, inst = _case(21, gain=1.0)
    est_inst = (inst + bleed * voc).astype(np.float32)       # an instrumental estimate with vocal bleed
    for blend in (0.2, 0.4, 1.5):
        want = ref.debleed_instrumental(mix, voc, est_inst, 44100, blend)
        got = orch.debleed_instrumental(torch.from_numpy(mix), torch.from_numpy(voc), torch.from_numpy(est_inst), 44100,
                                        blend).numpy()
        assert np.abs(got - want).max() <= 2e-6
    silent = np.zeros_like(inst)
    want = ref.debleed_instrumental(mix, voc, silent, 44100, 0.4)
    got = orch.debleed_instrumental(torch.from_numpy(mix), torch.from_numpy(voc), torch.from_numpy(silent), 44100, 0.4).numpy()
    assert np.abs(want).max() > 1e-3 and np.abs(got - want).max() <= 2e-6


def test_pcm16_roundtrip_matches_reference():
    x = (np.random.RandomState(3).uniform(-1.2, 1.2, size=(2, 5000))).astype(np.float32)
    x[0, :4] = [1.0, -1.0, 0.5 / 32768, 1.5 / 32768]
    assert np.array_equal(orch.pcm16_roundtrip(torch.from_numpy(x)).numpy(), ref.pcm16_roundtrip(x))


def test_ensemble_selection_follows_the_reference_list():
    assert [m[0] for m in orch.ensemble_models(1)] == ["vocals_mel_band_roformer.ckpt"]
    assert [m[0] for m in orch.ensemble_models(3)][-1] == "melband_roformer_big_beta4.ckpt"
    assert len(orch.ENSEMBLE) == 7 and orch.ENSEMBLE[5][0] == "Kim_Vocal_2.onnx" and orch.ENSEMBLE[6][1] == 6.8
    with pytest.raises(NotImplementedError):
        orch.ensemble_models(4)


class _StubSeparator:
    """Separator duck type on the CPU: 'separates' by fixed filters so that the orchestration is checkable."""
    sample_rate = 44100

    def __init__(self):
        self.loaded = []
        self.inputs = []

    def prepare_mix(self, audio, sr):
        return audio.float()

    def load_model(self, name):
        self.loaded.append(name)
        self.name = name

    def separate_tensor(self, mix):
        self.inputs.append(mix.clone())
        if self.name.startswith("htdemucs"):
            return {k: mix * (0.1 * (i + 1)) for i, k in enumerate(["Drums", "Bass", "Other", "Vocals", "Guitar", "Piano"])}
        k = 0.5 + 0.1 * len(self.loaded)
        return {"Vocals": mix * k, "Instrumental": mix * (1 - k)}


def test_separate_music_orchestration(tmp_path):
    from audiolab_b200.wavio import read_wav, write_wav
    mix = synth_mix(20000, seed=9)
    src = tmp_path / "song.wav"
    write_wav(str(src), mix, 44100, "FLOAT")
    out_dir = str(tmp_path / "stems")
    stub = _StubSeparator()
    seen = []
    outs = orch.separate_music({out_dir: [str(src)]}, callback=lambda f, d, t: seen.append((f, t)), separator=stub,
                               vocals_only=False, ensemble_strength=2, pcm16_handoff=True)
    assert stub.loaded == ["vocals_mel_band_roformer.ckpt", "model_bs_roformer_ep_368_sdr_12.9628.ckpt", "htdemucs_6s.yaml"]
    names = sorted(os.path.basename(o) for o in outs)
    assert names == sorted(f"song__({s}).wav" for s in ("Vocals", "Instrumental", "Drums", "Bass", "Guitar", "Piano", "Other"))
    assert seen[0][0] == 0 and seen[-1][0] == 1.0 and all(a[0] <= b[0] for a, b in zip(seen, seen[1:]))
    # every model saw the PCM_16 image of the mix (write_temp_wav hand-off), not the float mix
    q = ref.pcm16_roundtrip(mix)
    assert all(np.array_equal(x.numpy(), q) for x in stub.inputs)
    # vocals = reference blend of the two model outputs
    want_v = ref.blend_tracks([q * 0.6, q * 0.7], [8.6, 8.4])
    got_v, _ = read_wav(os.path.join(out_dir, "song__(Vocals).wav"))
    assert np.abs(got_v - want_v).max() <= 1e-6
    want_i = ref.debleed_instrumental(mix, want_v, ref.blend_tracks([q * 0.4, q * 0.3], [16.0, 16.0]), 44100, 0.2)
    got_i, _ = read_wav(os.path.join(out_dir, "song__(Instrumental).wav"))
    assert np.abs(got_i - want_i).max() <= 2e-6


@pytest.mark.parametrize("stem", ["(vocals)", "(Vocals)", "(instrumental)", "(bg_vocals)", "(bg_vocals) (vocals)", "(drums)"])
@pytest.mark.parametrize("setting", ["Nothing", "All", "All Vocals", "Main Vocals", "Something else"])
def test_should_apply_transform_truth_table(stem, setting):
    from oracle import transforms as tref
    assert orch.should_apply_transform(stem, setting) == tref.should_apply_transform(stem, setting)


class _ChainSeparator(_StubSeparator):
    """Transform models scale their input by a model-specific factor; output names follow audiolab_b200.separator.STEMS_OF."""
    FACT = {"dereverb": 0.9, "dereverb-echo": 0.8, "UVR-MDX-NET_Crowd": 0.7}

    def separate_tensor(self, mix):
        from audiolab_b200.separator import STEMS_OF
        if self.name in STEMS_OF:
            self.inputs.append(mix.clone())
            prim, sec = STEMS_OF[self.name]
            k = next(v for p, v in sorted(self.FACT.items(), key=lambda kv: -len(kv[0])) if self.name.startswith(p))
            return {prim: mix * k, sec: mix * (1 - k)}
        return super().separate_tensor(mix)


@pytest.mark.parametrize("opts", [
    dict(reverb_removal="Main Vocals"),
    dict(reverb_removal="All Vocals", echo_removal="All Vocals", crowd_removal="All"),
    dict(crowd_removal="All"),
    dict(echo_removal="All"),                      # the reference only enters the chain for reverb / crowd / noise: no-op
    dict(reverb_removal="All", crowd_removal="Main Vocals", delay_removal="All"),
])
def test_transform_chain_follows_the_reference_order_and_output_choice(tmp_path, opts):
    """separate_music's transform stage against the oracle restatement of stem_separator.py:777-839 / :903-934 with the same
    stub models: order of the loaded models, the chain running twice on the vocals when reverb AND crowd removal are set,
    and which of a model's two outputs is kept."""
    from audiolab_b200.separator import STEMS_OF
    from audiolab_b200.wavio import read_wav, write_wav
    from oracle import transforms as tref
    mix = synth_mix(12000, seed=11)
    src = tmp_path / "song.wav"
    write_wav(str(src), mix, 44100, "FLOAT")
    out_dir = str(tmp_path / "stems")
    stub = _ChainSeparator()
    orch.separate_music({out_dir: [str(src)]}, separator=stub, vocals_only=True, ensemble_strength=1, **opts)
    # oracle side: the same fake models on numpy arrays
    log = []

    def run_model(model_file, arr):
        log.append(model_file)
        prim, sec = STEMS_OF[model_file]
        k = next(v for p, v in sorted(_ChainSeparator.FACT.items(), key=lambda kv: -len(kv[0])) if model_file.startswith(p))
        return [(f"song_({prim})_{model_file}", arr * np.float32(k)), (f"song_({sec})_{model_file}", arr * np.float32(1 - k))]

    voc0 = ref.blend_tracks([mix * np.float32(0.6)], [8.6])       # the stub's ensemble output with one model loaded
    inst0 = ref.debleed_instrumental(mix, voc0, ref.blend_tracks([mix * np.float32(0.4)], [16.0]), 44100, 0.2)
    results = {"song": {"vocals": voc0, "instrumental": inst0}}
    tref.transform_stage(results, opts, run_model)
    assert stub.loaded[1:] == log
    got_v, _ = read_wav(os.path.join(out_dir, "song__(Vocals).wav"))
    got_i, _ = read_wav(os.path.join(out_dir, "song__(Instrumental).wav"))
    assert np.abs(got_v - results["song"]["vocals"]).max() <= 1e-6
    assert np.abs(got_i - results["song"]["instrumental"]).max() <= 2e-6


def test_transform_chain_refuses_models_outside_the_scope(tmp_path):
    from audiolab_b200.wavio import write_wav
    src = tmp_path / "song.wav"
    write_wav(str(src), synth_mix(6000, seed=12), 44100, "FLOAT")
    with pytest.raises(NotImplementedError):       # UVR-DeNoise.pth is a VR-architecture model
        orch.separate_music({str(tmp_path / "o"): [str(src)]}, separator=_ChainSeparator(), ensemble_strength=1, noise_removal="All")
    with pytest.raises(ValueError):
        orch.separate_music({str(tmp_path / "o"): [str(src)]}, separator=_ChainSeparator(), ensemble_strength=1, crowd_removal="Everything")
    with pytest.raises(NotImplementedError):
        orch.separate_music({str(tmp_path / "o"): [str(src)]}, separator=_ChainSeparator(), ensemble_strength=1, separate_drums=True)


def _reverb_case(n=30000, sr=8000, seed=5):
    rs = np.random.RandomState(seed)
    dry = (rs.randn(n, 2) * np.exp(-np.arange(n) / 9000.0)[:, None]).astype(np.float32) * 0.3
    t = np.arange(int(0.3 * sr))
    ir = np.exp(-t / (0.04 * sr)) * rs.randn(t.size) * 0.2
    ir[0] = 1.0
    wet = np.stack([np.convolve(dry[:, c], ir)[:n] for c in range(2)], axis=1).astype(np.float32)
    # an envelope whose dB curve IS the model the reference fits (a exp(-b t) + c, b = 3 -> "decay time" 1 s): a well-posed fit
    tt = np.arange(n) / sr
    env = 10.0 ** ((40.0 * np.exp(-3.0 * tt) - 60.0) / 20.0)
    wet = (np.sign(wet + 1e-12) * env[:, None] / np.sqrt(2.0)).astype(np.float32)
    return dry, wet, sr


def test_reverb_ir_extraction_matches_the_reference_restatement(tmp_path):
    """audiolab_b200/reverb_ir.py (torch.fft, fp64) against oracle/reverb_ir.py (numpy restatement of handlers/reverb.py:113-172):
    pre-delay, RT60 fit, Wiener-deconvolved impulse response and its statistics; the JSON file of the reference."""
    import json
    from audiolab_b200 import reverb_ir
    from oracle import reverb_ir as rref
    dry, wet, sr = _reverb_case()
    want = rref.extract_params(dry, wet, sr)
    got = reverb_ir.extract_reverb_params(torch.from_numpy(dry.T.copy()), torch.from_numpy(wet.T.copy()), sr)
    assert set(got) == set(want)
    for k in ("sample_rate", "pre_delay"):
        assert got[k] == want[k]
    assert abs(got["decay_time"] - 1.0) < 1e-3 and abs(got["decay_time"] - want["decay_time"]) <= 1e-6
    # fp64 on both sides; the division by |H|^2 + 1e-6 amplifies the last-bit differences of two FFT libraries
    for k in ("early_reflection_ratio", "late_reverb_ratio", "diffusion", "spectral_centroid"):
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), k
    ir_w = np.array(want["impulse_response"])
    assert np.abs(np.array(got["impulse_response"]) - ir_w).max() <= 1e-6 * np.abs(ir_w).max()
    p = reverb_ir.extract_reverb(torch.from_numpy(dry.T.copy()), torch.from_numpy(wet.T.copy()), sr, str(tmp_path / "ir.json"))
    assert json.load(open(p))["decay_time"] == got["decay_time"]


def test_store_reverb_ir_writes_the_impulse_response_of_the_vocals_pass(tmp_path):
    import json
    from audiolab_b200.wavio import write_wav
    src = tmp_path / "song.wav"
    write_wav(str(src), synth_mix(9000, seed=13), 44100, "FLOAT")
    out_dir = str(tmp_path / "stems")
    orch.separate_music({out_dir: [str(src)]}, separator=_ChainSeparator(), ensemble_strength=1, reverb_removal="All",
                        store_reverb_ir=True)
    params = json.load(open(os.path.join(out_dir, "impulse_response.ir")))
    assert params["sample_rate"] == 44100 and len(params["impulse_response"]) == 9000   # shorter than 2 s: the whole signal


@pytest.mark.gpu
def test_orchestrator_gpu():
    """The same de-bleed arithmetic on the device, and the real Separator through separate_music (random-init nets)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    mix, voc, inst = _case(31, lag=123, gain=0.9)
    est = (inst + 0.3 * voc).astype(np.float32)
    want = ref.debleed_instrumental(mix, voc, est, 44100, 0.4)
    got = orch.debleed_instrumental(torch.from_numpy(mix).cuda(), torch.from_numpy(voc).cuda(), torch.from_numpy(est).cuda(),
                                    44100, 0.4).cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6


@pytest.mark.gpu
def test_transform_chain_gpu(tmp_path):
    """The transform stage with the real Separator on the device (small random-init networks through model_overrides): the
    de-reverb Mel-band RoFormer and the crowd MDX-Net load under the reference's file names, their outputs carry the stem
    names the chain looks for, and the vocals that come out are the chain applied to the ensemble's vocals."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from audiolab_b200.separator import Separator
    from audiolab_b200.wavio import read_wav, write_wav
    small = dict(dim=64, depth=1, heads=2, dim_head=32, chunk_size=441 * 60)
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=False,
                    mdx_params={"segment_size": 16, "full_size_net": False},
                    model_overrides={"mel_roformer": small, "bs_roformer": small})
    mix = synth_mix(441 * 150, seed=21)
    src = tmp_path / "clip.wav"
    write_wav(str(src), mix, 44100, "FLOAT")
    out_dir = str(tmp_path / "stems")
    outs = orch.separate_music({out_dir: [str(src)]}, separator=sep, ensemble_strength=1, reverb_removal="Main Vocals",
                               crowd_removal="All")
    assert sorted(os.path.basename(o) for o in outs) == ["clip__(Instrumental).wav", "clip__(Vocals).wav"]
    # the same stages by hand
    x = sep.prepare_mix(torch.from_numpy(mix), 44100)
    sep.load_model("vocals_mel_band_roformer.ckpt")
    st = sep.separate_tensor(x)
    voc = orch.blend_tracks([st["Vocals"]], [8.6])
    inst = orch.debleed_instrumental(x, voc, orch.blend_tracks([st["Instrumental"]], [16.0]), 44100, 0.2)
    for name, key in (("dereverb_mel_band_roformer_anvuew_sdr_19.1729.ckpt", "noreverb"), ("UVR-MDX-NET_Crowd_HQ_1.onnx", "No Crowd"),
                      ("UVR-MDX-NET_Crowd_HQ_1.onnx", "No Crowd")):
        sep.load_model(name)
        assert list(sep.separate_tensor(voc))[0] == key
        voc = sep.separate_tensor(voc)[key]
    sep.load_model("UVR-MDX-NET_Crowd_HQ_1.onnx")
    inst = sep.separate_tensor(inst)["No Crowd"]
    got_v, _ = read_wav(os.path.join(out_dir, "clip__(Vocals).wav"))
    got_i, _ = read_wav(os.path.join(out_dir, "clip__(Instrumental).wav"))
    assert np.abs(got_v - voc.cpu().numpy()).max() <= 1e-6
    assert np.abs(got_i - inst.cpu().numpy()).max() <= 1e-6


@pytest.mark.gpu
def test_reverb_ir_extraction_gpu():
    """The same extraction with the stems on the device (cuFFT fp64) against the numpy restatement of the reference."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from audiolab_b200 import reverb_ir
    from oracle import reverb_ir as rref
    dry, wet, sr = _reverb_case(n=120000, sr=44100, seed=6)
    want = rref.extract_params(dry, wet, sr)
    got = reverb_ir.extract_reverb_params(torch.from_numpy(dry.T.copy()).cuda(), torch.from_numpy(wet.T.copy()).cuda(), sr)
    assert got["pre_delay"] == want["pre_delay"] and abs(got["decay_time"] - want["decay_time"]) <= 1e-5
    for k in ("early_reflection_ratio", "late_reverb_ratio", "diffusion", "spectral_centroid"):
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), k
    ir_w = np.array(want["impulse_response"])
    assert np.abs(np.array(got["impulse_response"]) - ir_w).max() <= 1e-6 * np.abs(ir_w).max()
