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
