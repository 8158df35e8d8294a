"""Drop-ins for the reference's inference_utils.py: interpolate_mask, mix_song_smooth, mix_song.

What changes under the hood (same results): all chunks of a song are featurised by ONE K1 launch per
stem straight into the [chunks, 4, 1025, T] model input (stereo -> mono fused into the staging), the
CNN runs once on that batch instead of once per chunk with a device->host sync each time, and the
zero-order-hold mask, the per-stem multiply and (optionally) the stem sum are one K2 pass.  The tiny
gain post-processing (10**(0.5 m) and the Savitzky-Golay filter over <= a few hundred values) stays
on the host with numpy/scipy, exactly as the reference computes it.
"""
import numpy as np
import torch
from scipy.signal import savgol_filter

from . import features, mixing

STEMS = ('bass', 'drums', 'vocals', 'other')


def scalar_amplitude_to_dB(x):
    """data/dataset_utils.py:39-43."""
    return 20 * np.log10(x)


def scalar_dB_to_amplitude(x):
    """data/dataset_utils.py:46-50 (10**(0.5 x), sic)."""
    return np.power(10.0, 0.5 * x)


def _device(device=None):
    return torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())


def stack_tracks(tracks: dict, keys=STEMS, device=None):
    """dict of [C, N] arrays/tensors -> ([S, C, N] CUDA tensor, inputs_were_host)."""
    first = tracks[keys[0]]
    host = not (isinstance(first, torch.Tensor) and first.is_cuda)
    dev = _device(device) if host else first.device
    parts = []
    for k in keys:
        t = tracks[k]
        t = torch.from_numpy(np.ascontiguousarray(t)) if isinstance(t, np.ndarray) else t
        parts.append(t.to(dev, non_blocking=True))
    return torch.stack(parts), host


def unstack_tracks(stems: torch.Tensor, keys=STEMS, to_host=True):
    if to_host:
        arr = stems.cpu().numpy()
        return {k: arr[i] for i, k in enumerate(keys)}
    return {k: stems[i] for i, k in enumerate(keys)}


def interpolate_mask(spec_mask, tgt_len: int, device=None):
    """inference_utils.py:12-41: zero-order hold of len(spec_mask) values over tgt_len samples
    (step int(tgt_len / K), last value fills the tail, K == 1 gives zeros).  Computed by K2 on a
    vector of ones; returns a float64 numpy array like the reference."""
    spec_mask = np.asarray(spec_mask, dtype=np.float64).reshape(-1)
    assert len(spec_mask) <= tgt_len, "Target mask should be longer than the initial one"
    ones = torch.ones((1, 1, tgt_len), dtype=torch.float64, device=_device(device))
    out = mixing.gain_mix(ones, spec_mask[None, :], want_mix=True)['mix']
    return out[0].cpu().numpy()


def song_features(stems: torch.Tensor, chunk_samples: int, n_chunks: int, window_size=2048,
                  hop_length=1024, out_dtype=torch.float32, compute_dtype=None) -> torch.Tensor:
    """stems [S, C, N] -> [n_chunks, S, F, T]: features of the first n_chunks chunks of every stem,
    channels averaged (the D4 mono shim), written in place by one K1 launch per stem."""
    S, C, N = stems.shape
    F = window_size // 2 + 1
    T = features.num_frames(chunk_samples, hop_length)
    out = torch.empty((n_chunks, S, F, T), dtype=out_dtype, device=stems.device)
    if compute_dtype is None:
        compute_dtype = torch.float64 if stems.dtype == torch.float64 else torch.float32
    for s in range(S):
        # rows = chunks of stem s: [n_chunks, C, chunk_samples] view with row stride chunk_samples
        view = stems[s].as_strided((n_chunks, C, chunk_samples), (chunk_samples, stems.stride(1), 1),
                                   stems[s].storage_offset())
        if C > 1:
            features.stft_db(view, window_size, hop_length, channel_axis=-2, compute_dtype=compute_dtype,
                             out_dtype=out_dtype, out=out[:, s], out_row_stride=S * F * T)
        else:
            features.stft_db(view[:, 0], window_size, hop_length, compute_dtype=compute_dtype,
                             out_dtype=out_dtype, out=out[:, s], out_row_stride=S * F * T)
    return out


def predict_gains(model, feats: torch.Tensor, batch: int = 64) -> np.ndarray:
    """Run the CNN on [K, 4, F, T] features; returns raw head outputs m [K, 4] (float32, host)."""
    outs = []
    with torch.no_grad():
        for i in range(0, feats.shape[0], batch):
            _, ms = model(feats[i:i + batch].float())
            outs.append(torch.cat([m.reshape(-1, 1) for m in ms], dim=1))
    return torch.cat(outs).float().cpu().numpy()


def smooth_gains(raw, num_chunks: int) -> np.ndarray:
    """inference_utils.py:137-140: Savitzky-Golay, window odd(int(num_chunks/4)), order 2."""
    cand = int(num_chunks / 4)
    win = cand if cand % 2 else cand + 1
    return savgol_filter(raw, win, 2)


def mix_song_smooth(dataset, model, loaded_tracks: dict, chunk_length=1, sr=44100,
                    return_device: bool = False, want_mix: bool = False):
    """inference_utils.py:105-145.  loaded_tracks: {'bass'|'drums'|'vocals'|'other': [2, N]}.
    Returns (mixed_tracks, raw_gains, smooth_gains) like the reference: mixed tracks are float64
    [2, N] (numpy unless return_device), gains are dicts of python-float lists.  With want_mix the
    stem sum (evaluation.py:59-60) is produced by the same pass and returned as a 4th value."""
    keys = [t for t in (dataset.get_tracklist() if dataset is not None else list(STEMS) + ['mix']) if t != 'mix']
    stems, host = stack_tracks(loaded_tracks, keys)
    S, C, N = stems.shape
    chunk_samples = chunk_length * sr
    num_chunks = int(N / chunk_samples)
    n_an = num_chunks - 1                      # the last chunk is never analysed (range(1, num_chunks))
    feats = song_features(stems, chunk_samples, n_an)
    m = predict_gains(model, feats)            # [n_an, 4] float32
    raw_gains, smooth = {}, {}
    g = np.empty((S, n_an), dtype=np.float64)
    for i, t in enumerate(keys):
        raw = [float(scalar_dB_to_amplitude(np.squeeze(m[k, i]))) for k in range(n_an)]
        raw_gains[t] = raw
        sm = smooth_gains(raw, num_chunks)
        smooth[t] = list(sm)
        g[i] = sm
    res = mixing.gain_mix(stems, g, compute_dtype=torch.float64, want_mix=want_mix, want_stems=True)
    mixed = unstack_tracks(res['stems'], keys, to_host=host and not return_device)
    if want_mix:
        mix = res['mix'] if (return_device or not host) else res['mix'].cpu().numpy()
        return mixed, raw_gains, smooth, mix
    return mixed, raw_gains, smooth


def mix_song(dataset, model, loaded_tracks: dict, chunk_length=1, sr=44100):
    """inference_utils.py:44-102 (legacy; the checked-in version unpacks a tuple compute_features no
    longer returns).  Mono tracks [N]; chunk k (0-based, k < num_chunks-1) is multiplied by the RAW
    head output m_k of each stem and the stems are summed; samples past the analysed chunks stay 0.
    Returns (mixed_song float64 [N], mask_history)."""
    keys = [t for t in dataset.get_tracklist() if t != 'mix']
    tracks = {k: np.asarray(loaded_tracks[k])[None, :] if not isinstance(loaded_tracks[k], torch.Tensor)
              else loaded_tracks[k][None, :] for k in keys}
    stems, host = stack_tracks(tracks, keys)
    S, _, N = stems.shape
    chunk_samples = chunk_length * sr
    num_chunks = int(N / chunk_samples)
    n_an = num_chunks - 1
    feats = song_features(stems, chunk_samples, n_an)
    m = predict_gains(model, feats)
    mask_history = {t: [float(m[k, i]) for k in range(n_an)] for i, t in enumerate(keys)}
    covered = n_an * chunk_samples
    mixed = torch.zeros(N, dtype=torch.float64, device=stems.device)
    if n_an > 0:
        res = mixing.gain_mix(stems[:, :, :covered], m.T.astype(np.float64), interp_coef=chunk_samples,
                              hold_single=True, compute_dtype=torch.float64)
        mixed[:covered] = res['mix'][0]
    return (mixed.cpu().numpy() if host else mixed), mask_history
