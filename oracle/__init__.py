"""CPU oracle for the AudioLab source-separation spectral hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``audiolab_b200/`` may import this
package.  The only legal importers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there
only as the checker (or the timed CPU baseline), never as the product path.

What it restates (pure PyTorch fp32 on the CPU, numpy/scipy where the reference
uses them):

* ``oracle.mdx``       -- the in-tree UVR5 MDX-Net twin,
                          /root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py
                          (ConvTDFNetTrim.stft :41-56, .istft :58-75,
                          Predictor.demix :109-141, Predictor.demix_base :143-197)
                          plus the windowed overlap-add form of upstream
                          ``audio_separator`` MDXSeparator.demix.
* ``oracle.roformer``  -- BS-RoFormer / Mel-Band RoFormer forward (STFT -> bands ->
                          axial transformers -> complex mask (.) STFT -> iSTFT) and the
                          MSST-style chunk loop of MDXCSeparator.demix.
* ``oracle.htdemucs``  -- HTDemucs ``_spec`` / ``_ispec`` / ``apply_model`` split loop.
* ``oracle.resample``  -- scipy.signal.resample_poly(x, 147, 160) (48 k -> 44.1 k).

PARITY PINNING
--------------
* MDX-Net path: PINNED.  ``tests/golden/make_mdx_golden.py`` imports the
  reference's own ``mdxnet.py`` from /root/reference (librosa / soundfile /
  onnxruntime, which that file imports but which the spectral code never calls,
  are stubbed), runs ``ConvTDFNetTrim.stft/istft`` and ``Predictor.demix`` on
  seeded inputs and commits the outputs under ``tests/golden/``.  The oracle is
  checked against those vectors in ``tests/test_oracle_golden.py``.
* RoFormer / HTDemucs / resample paths: PARITY UNPINNED by any reference test.
  The arithmetic lives in the third-party dependency ``audio-separator[gpu]>=0.32.0``
  (/root/reference/setup.sh:96), which is neither vendored in /root/reference nor
  installed here, and the reference ships no tests or golden vectors
  (SURVEY.md section 4).  These are restatements of the published upstream
  algorithms; the ground truth named by BASELINE.json is
  ``torch.stft`` / ``torch.istft`` themselves, which is what these functions call.
"""

from . import metrics, synth  # noqa: F401
