"""CPU oracle for the deep-audio-mixer hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of the arithmetic the reference
(apelykh/deep-audio-mixer) performs on its STFT -> dB -> gains -> mix path.  It is
the checker the CUDA kernels are compared against; it is never the thing shipped
or measured as the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing
under ``deep_audio_mixer_b200/`` imports it, and that package raises if its CUDA
library is missing instead of falling back to this code.

Pinning status (see DESIGN.md "Oracle"):

* features / mix / model tail / DummyModel: PINNED.  The reference is pure Python;
  it was imported unmodified from ``/root/reference`` in the build container (with
  empty stub modules for the absent ``soundfile`` / ``pyloudnorm`` / ``librosa``,
  see ``tools/ref_harness.py``) and its outputs on seeded inputs are committed as
  ``tests/golden/*.npz`` by ``tools/gen_golden.py``.  ``tests/test_oracle_golden.py``
  checks every oracle function against those vectors, and against the
  known-answer values recorded in the reference's notebooks.
* loudness (ITU-R BS.1770 as implemented by pyloudnorm): CORROBORATED, not pinned by the
  reference.  ``pyloudnorm`` is a third-party dependency that is absent from the reference
  tree and from this image (no pinned version: the reference has no requirements file;
  2020-era => pyloudnorm 0.1.0), and no reference test holds a LUFS value.
  ``oracle/loudness.py`` restates its published algorithm.  Pins: (1) an independent
  implementation that IS installed here -- ``torchaudio.functional.loudness``, whose
  documentation states it follows pyloudnorm -- agrees to ~1e-14 LU at 44.1 kHz on seven
  signal classes including absolute- and relative-gated ones
  (``tests/test_oracle_golden.py::test_loudness_matches_torchaudio``); (2) the linearity
  property recorded in ``experiments.ipynb:649-678``; (3) the BS.1770 conformance values.
"""
