"""CPU port of the reference's hot path for timing (TEST / BENCH INFRASTRUCTURE, see oracle/__init__.py).

``bench.py --impl reference`` and the ``cpu_baseline`` leg time THIS module on the GPU box's host
cores: the reference itself is pure Python and cannot travel (no /root/reference on the box), so the
baseline is a port that makes the same library calls the reference makes, in the same order:

  features   torch.stft(n_fft, hop, torch.hann_window, return_complex=True) -> torch.abs ->
             20*log10(clamp(., 1e-5))           data/dataset.py:145-155 (per stem per chunk, like
             _process_on_the_fly :192-205 and mix_song_smooth inference_utils.py:115-119)
  mono       np.mean over the channel axis      data/dataset.py:181-183
  gains      np.power(10.0, 0.5*m)              data/dataset_utils.py:46-50
  mask, mix  interpolate_mask loop, track*mask, np.sum(np.array(...), axis=0)
             inference_utils.py:12-41,143; evaluation.py:59-60
  bus sum    summed_track += track              data/medleydb_data_surgery.ipynb:321-326
  loudness   pyloudnorm restatement (oracle/loudness.py), normalise evaluation.py:63-64
"""
import numpy as np
import torch

from . import loudness as OL
from . import mix as OM


def compute_features_torch(audio: np.ndarray, window_size: int = 2048, hop_length: int = 1024) -> torch.Tensor:
    """The reference's compute_features body, call for call (window rebuilt every call included)."""
    spectrum = torch.stft(torch.from_numpy(audio), n_fft=window_size, hop_length=hop_length,
                          window=torch.hann_window(window_size), return_complex=True)
    magnitudes = torch.abs(spectrum)
    return torch.log10(torch.clamp(magnitudes, min=1e-5)) * 20.0


def segment_step(raw: np.ndarray, group, m: np.ndarray, chunk_samples: int, hop: int, sr: int = 44100,
                 target: float = -20.0, feature_dtype=None):
    """One segment [S, C, N] float32 through bus sum -> features -> gains -> mix -> loudness ->
    normalise.  m [n_chunks, 4] are the head outputs used for the gains.  ``feature_dtype`` = np.float64
    featurises in double precision like the training path (data/dataset.py reads float64); the default
    follows the audio (float32, the inference path).  Returns (features [n_chunks, 4, F, T], normalised
    mix [C, N] float64, loudness)."""
    S, C, N = raw.shape
    buses = OM.bus_sum(raw, group, 4)                                           # [4, C, N] float32
    n_chunks = N // chunk_samples
    feats = []
    for k in range(n_chunks):
        per = []
        for s in range(4):
            chunk = buses[s][:, k * chunk_samples:(k + 1) * chunk_samples]
            mono = np.mean(chunk, axis=0)
            if feature_dtype is not None:
                mono = mono.astype(feature_dtype)
            per.append(compute_features_torch(mono, 2048, hop))
        feats.append(torch.stack(per))
    feats = torch.stack(feats)
    gains = np.power(10.0, 0.5 * m.astype(np.float32)).astype(np.float64)       # [n_chunks, 4]
    mixed = [buses[s] * OM.interpolate_mask(gains[:, s], N) for s in range(4)]
    mix = np.sum(np.array(mixed), axis=0)                                        # [C, N] float64
    L = OL.integrated_loudness(mix.T, sr)
    return feats, OM.normalize_loudness(mix, L, target), L


def batch_step(raw: np.ndarray, group, m: np.ndarray, chunk_samples: int, hop: int, sr: int = 44100):
    """raw [B, S, C, N]; m [B * n_chunks, 4].  Returns (mixes [B, C, N], loudness [B])."""
    B = raw.shape[0]
    nc = raw.shape[-1] // chunk_samples
    out, loud = [], []
    for b in range(B):
        _, mix, L = segment_step(raw[b], group, m[b * nc:(b + 1) * nc], chunk_samples, hop, sr)
        out.append(mix)
        loud.append(L)
    return np.stack(out), np.array(loud)
