"""Oracle: gains -> masks -> mixdown, and the models' tail (TEST INFRASTRUCTURE).

Restates with numpy/scipy:

* ``interpolate_mask``               /root/reference/inference_utils.py:12-41
* ``mix_song_smooth``                /root/reference/inference_utils.py:105-145
* ``scalar_dB_to_amplitude`` etc.    /root/reference/data/dataset_utils.py:39-50
* stem sum                           /root/reference/evaluation.py:59-60,
                                     /root/reference/data/listening_test_data_preparation.py:20-21
* model tail ``masked = sum m_i x_i`` /root/reference/models/model_scalar_1s.py:260-275,
                                     model_scalar_2s.py:117-132, model_resnet.py:120-126
* ``DummyModel.forward``             /root/reference/models/baselines/dummy_model.py:23-34
* ``RandomModel.forward``            /root/reference/models/baselines/random_model.py:10-14
* ``pyln.normalize.loudness``        x * 10**((target - L)/20)   (evaluation.py:64)
* ``sum_stems`` (S raw stems -> bus) /root/reference/data/medleydb_data_surgery.ipynb:305-332

Reference quirks kept on purpose: the last chunk is never analysed;
gains are stretched by int(N/K), not by the chunk length; gain = 10**(0.5*m); the mask is
float64 so mixed tracks are float64; a single-gain mask is all zeros.
"""
import numpy as np
from scipy.signal import savgol_filter

from . import features as _features

STEMS = ('bass', 'drums', 'vocals', 'other')


def scalar_amplitude_to_dB(x):
    return 20 * np.log10(x)


def scalar_dB_to_amplitude(x):
    return np.power(10.0, 0.5 * x)


def interpolate_mask(spec_mask, tgt_len: int) -> np.ndarray:
    """Zero-order hold of K gains over tgt_len samples (inference_utils.py:12-41)."""
    assert len(spec_mask) <= tgt_len, "Target mask should be longer than the initial one"
    sample_mask = np.zeros(tgt_len)
    interp_coef = int(tgt_len / len(spec_mask))
    final_i = -1
    for chunk_i in range(0, len(spec_mask) - 1):
        i_from = chunk_i * interp_coef
        i_to = (chunk_i + 1) * interp_coef
        sample_mask[i_from:i_to] = spec_mask[chunk_i]
        final_i = i_to
    if final_i > -1:
        sample_mask[final_i:] = spec_mask[-1]
    return sample_mask


def savgol_window(num_chunks: int) -> int:
    """inference_utils.py:137-139."""
    cand = int(num_chunks / 4)
    return cand if cand % 2 else cand + 1


def smooth_gains(raw, num_chunks: int) -> np.ndarray:
    """inference_utils.py:140 (scipy default mode='interp', polyorder 2)."""
    return savgol_filter(raw, savgol_window(num_chunks), 2)


def chunk_features(loaded_tracks: dict, i_from: int, i_to: int, window_size=2048,
                   hop_length=1024) -> np.ndarray:
    """inference_utils.py:115-121 with the D4 mono shim: the checked-in code
    passes [2, n] slices, which makes the model input 5-D and raises; the training path it was
    written against down-mixes with a channel mean first (data/dataset.py:181-183,195-196)."""
    feats = []
    for track in STEMS:
        chunk = loaded_tracks[track][:, i_from:i_to]
        mono = np.mean(chunk, axis=0)
        feats.append(_features.compute_features(mono, window_size, hop_length))
    return np.stack(feats)


def mix_song_smooth(model_fn, loaded_tracks: dict, chunk_length=1, sr=44100,
                    window_size=2048, hop_length=1024):
    """inference_utils.py:105-145.  ``model_fn(features[4, F, T]) -> 4 floats`` stands for
    ``model(feature_tensor.unsqueeze(0))[1]`` (raw head outputs m_i)."""
    chunk_samples = chunk_length * sr
    n = len(loaded_tracks['drums'][0])
    num_chunks = int(n / chunk_samples)
    raw_gains = {t: [] for t in STEMS}
    for chunk_i in range(1, num_chunks):
        i_from = (chunk_i - 1) * chunk_samples
        i_to = chunk_i * chunk_samples
        m = model_fn(chunk_features(loaded_tracks, i_from, i_to, window_size, hop_length))
        for i, t in enumerate(STEMS):
            raw_gains[t].append(float(scalar_dB_to_amplitude(np.squeeze(m[i]))))
    smooth = {t: [] for t in STEMS}
    mixed = {}
    for t in raw_gains:
        s = smooth_gains(raw_gains[t], num_chunks)
        smooth[t].extend(s)
        mask = interpolate_mask(s, len(loaded_tracks[t][0]))
        mixed[t] = loaded_tracks[t] * mask
    return mixed, raw_gains, smooth


def apply_gains(stems: np.ndarray, gains: np.ndarray) -> np.ndarray:
    """stems [S, C, N] x per-stem gain sequences [S, K] -> float64 [S, C, N]
    (interpolate_mask + ``track * mask``, inference_utils.py:142-143)."""
    out = np.empty(stems.shape, dtype=np.float64)
    for s in range(stems.shape[0]):
        out[s] = stems[s] * interpolate_mask(gains[s], stems.shape[-1])
    return out


def stem_sum(track_dict_or_array) -> np.ndarray:
    """evaluation.py:59-60: np.sum(np.array(list(d.values())), axis=0)."""
    if isinstance(track_dict_or_array, dict):
        arr = np.array(list(track_dict_or_array.values()))
    else:
        arr = np.asarray(track_dict_or_array)
    return np.sum(arr, axis=0)


def bus_sum(stems: np.ndarray, group, n_bus: int = 4) -> np.ndarray:
    """medleydb_data_surgery.ipynb:321-326: summed_track += track, in stem order, in the
    audio dtype.  stems [S, ...], group[s] in [0, n_bus) or -1 (dropped)."""
    out = np.zeros((n_bus,) + stems.shape[1:], dtype=stems.dtype)
    for s, g in enumerate(group):
        if g >= 0:
            out[g] += stems[s]
    return out


def random_gains(rng_uniform, n=4, lo=0.5, hi=1.5):
    """random_model.py:11: float(np.random.uniform(lo, hi)) per stem."""
    return [float(rng_uniform(lo, hi)) for _ in range(n)]


def normalize_loudness(data: np.ndarray, loudness: float, target: float) -> np.ndarray:
    """pyloudnorm.normalize.loudness: gain = 10**((target - loudness)/20)."""
    gain = np.power(10.0, (target - loudness) / 20.0)
    return gain * data


def peak_normalize(x: np.ndarray, axis=1) -> np.ndarray:
    """librosa.util.normalize(x, axis=1) with norm=inf (inference.ipynb:219): divide by max|x|
    along the axis (rows whose peak is below tiny stay unscaled)."""
    mag = np.max(np.abs(x), axis=axis, keepdims=True)
    tiny = np.finfo(x.dtype).tiny
    scale = np.where(mag < tiny, 1.0, mag)
    return x / scale


def feature_mix(x: np.ndarray, m: np.ndarray) -> np.ndarray:
    """Model tail (model_scalar_1s.py:260-275): masked[b] = sum_i m[b, i] * x[b, i], in the
    accumulation order of the reference (zeros, += m1*x0, += m2*x1, ...)."""
    masked = np.zeros_like(x[:, 0])
    for i in range(x.shape[1]):
        masked += m[:, i].reshape(-1, 1, 1) * x[:, i]
    return masked


def feature_mix_grad(x: np.ndarray, m: np.ndarray, grad_out: np.ndarray):
    """Analytic gradients of feature_mix: dL/dm[b,i] = sum_{f,t} g*x_i ; dL/dx[b,i] = m*g."""
    gm = np.einsum('bft,bift->bi', grad_out.astype(np.float64), x.astype(np.float64))
    gx = m[:, :, None, None] * grad_out[:, None]
    return gm, gx


def db_sum(x: np.ndarray) -> np.ndarray:
    """DummyModel.forward (dummy_model.py:23-34): 20*log10(sum_i 10**(0.5*x_i))."""
    masked = np.zeros_like(x[:, 0])
    for i in range(4):
        masked += np.power(x.dtype.type(10.0), x.dtype.type(0.5) * x[:, i])
    return (x.dtype.type(20) * np.log10(masked)).astype(x.dtype)
