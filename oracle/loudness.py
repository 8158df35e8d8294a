"""Oracle: ITU-R BS.1770-4 integrated loudness as pyloudnorm computes it (TEST INFRASTRUCTURE).

PARITY: no reference vector exists ("unpinned" by the reference), corroborated against the
installed ``torchaudio.functional.loudness`` to 1e-9 LU (tests/test_oracle_golden.py::
test_loudness_matches_torchaudio).  The arithmetic lives in the third-party package ``pyloudnorm`` (imported at
/root/reference/data/dataset.py:5, evaluation.py:7, models/baselines/mean_loudness_model.py:1,
data/listening_test_data_preparation.py:4).  It is not vendored in the reference, not installed
in this image, and the reference pins no version (no requirements file; the 2020 commit dates
imply pyloudnorm 0.1.0).  This file restates that release's published algorithm
(``Meter.integrated_loudness`` with the default "K-weighting" filter class and 0.4 s blocks,
``IIRfilter`` RBJ-cookbook biquads normalised by a0, applied with scipy.signal.lfilter, and
``normalize.loudness``).  What anchors it: the reference's call sites
(``meter.integrated_loudness(x.T)`` on [N, C] float arrays), the linearity values recorded in
experiments.ipynb:649-678 (L(g*x) = L(x) + 20 log10 g), and BS.1770 conformance numbers
(997 Hz 0 dBFS sine -> -3.01 LKFS within the 0.04 LU passband-gain offset of the RBJ form).
"""
import numpy as np
import scipy.signal

G_CHANNEL = [1.0, 1.0, 1.0, 1.41, 1.41]
BLOCK_SIZE = 0.400
OVERLAP = 0.75
GAMMA_A = -70.0


def biquad_coefficients(G: float, Q: float, fc: float, rate: float, filter_type: str):
    """pyloudnorm.IIRfilter.generate_coefficients (RBJ audio-EQ cookbook), returns (b, a)/a0."""
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    if filter_type == 'high_shelf':
        b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
        b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
        b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
        a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
        a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
        a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
    elif filter_type == 'high_pass':
        b0 = (1 + np.cos(w0)) / 2
        b1 = -(1 + np.cos(w0))
        b2 = (1 + np.cos(w0)) / 2
        a0 = 1 + alpha
        a1 = -2 * np.cos(w0)
        a2 = 1 - alpha
    else:
        raise ValueError(filter_type)
    return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0


def k_weighting_filters(rate: float):
    """Meter(rate) default filter chain: high-shelf (4 dB, Q 1/sqrt2, 1500 Hz) then
    high-pass (Q 0.5, 38 Hz)."""
    return [biquad_coefficients(4.0, 1 / np.sqrt(2), 1500.0, rate, 'high_shelf'),
            biquad_coefficients(0.0, 0.5, 38.0, rate, 'high_pass')]


def block_bounds(num_samples: int, rate: float):
    """Gating-block sample bounds exactly as pyloudnorm evaluates them in floating point."""
    T_g = BLOCK_SIZE
    step = 1.0 - OVERLAP
    T = num_samples / rate
    num_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
    lo = [int(T_g * (j * step) * rate) for j in range(num_blocks)]
    hi = [int(T_g * (j * step + 1) * rate) for j in range(num_blocks)]
    return np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64)


def k_weight(data: np.ndarray, rate: float) -> np.ndarray:
    """data [N, C] float -> K-weighted float64 [N, C]."""
    x = np.array(data, dtype=np.float64, copy=True)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    for b, a in k_weighting_filters(rate):
        for ch in range(x.shape[1]):
            x[:, ch] = scipy.signal.lfilter(b, a, x[:, ch])
    return x


def block_energies(data: np.ndarray, rate: float) -> np.ndarray:
    """z[c, j] = mean square of the K-weighted signal over gating block j."""
    x = k_weight(data, rate)
    if x.shape[0] < BLOCK_SIZE * rate:
        raise ValueError('Audio must have length greater than the block size.')
    lo, hi = block_bounds(x.shape[0], rate)
    z = np.zeros((x.shape[1], len(lo)))
    for i in range(x.shape[1]):
        for j in range(len(lo)):
            z[i, j] = (1.0 / (BLOCK_SIZE * rate)) * np.sum(np.square(x[lo[j]:hi[j], i]))
    return z


def gated_loudness(z: np.ndarray) -> float:
    """Absolute (-70 LUFS) and relative (-10 LU) gating over block energies z[c, j]."""
    C, J = z.shape
    G = G_CHANNEL
    with np.errstate(divide='ignore', invalid='ignore'):
        l = [-0.691 + 10.0 * np.log10(np.sum([G[i] * z[i, j] for i in range(C)])) for j in range(J)]
        J_g = [j for j, l_j in enumerate(l) if l_j >= GAMMA_A]
        z_avg = [np.mean([z[i, j] for j in J_g]) if J_g else np.nan for i in range(C)]
        gamma_r = -0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg[i] for i in range(C)])) - 10.0
        J_g = [j for j, l_j in enumerate(l) if (l_j > gamma_r and l_j > GAMMA_A)]
        z_avg = np.nan_to_num(np.array([np.mean([z[i, j] for j in J_g]) if J_g else np.nan
                                        for i in range(C)]))
        return float(-0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg[i] for i in range(C)])))


def integrated_loudness(data: np.ndarray, rate: float = 44100) -> float:
    """pyln.Meter(rate).integrated_loudness(data) for data [N, C] (or [N])."""
    return gated_loudness(block_energies(data, rate))


def evaluate_loudness(tracks: dict, rate: float = 44100, keys=('bass', 'drums', 'vocals', 'other')):
    """LoudnessEvaluator.evaluate_loudness (/root/reference/evaluation.py:39-46)."""
    per = [integrated_loudness(tracks[k].T, rate) for k in keys]
    avg = sum(per) / len(per)
    return [p - avg for p in per]


def loudness_dict_diff(a, b) -> float:
    """LoudnessEvaluator._calculate_diff_between_loudness_dicts (evaluation.py:48-53)."""
    return float(np.mean(np.abs(np.array(list(a)) - np.array(list(b)))))
