"""dB-STFT featurisation on the GPU (kernel K1 behind dam_stft_db).

Drop-in for ``MultitrackAudioDataset.compute_features`` (reference data/dataset.py:132-162) plus the
batched device-resident form the training / mixing paths use.
"""
import numpy as np
import torch

from . import _lib

_WINDOWS = {}


def hann_window(n_fft: int, device) -> torch.Tensor:
    """torch.hann_window(n_fft) (float32, periodic), exactly the table the reference passes to
    torch.stft (data/dataset.py:148); cached per device."""
    key = (n_fft, str(device))
    w = _WINDOWS.get(key)
    if w is None:
        w = torch.hann_window(n_fft, dtype=torch.float32).to(device)
        _WINDOWS[key] = w
    return w


def num_frames(n_samples: int, hop: int) -> int:
    return 1 + n_samples // hop


def stft_db(audio: torch.Tensor, n_fft: int = 2048, hop: int = 1024, window: torch.Tensor = None,
            channel_axis=None, amin: float = 1e-5, multiplier: float = 20.0,
            compute_dtype: torch.dtype = None, out_dtype: torch.dtype = None,
            input_scale: float = None, row_gain: torch.Tensor = None,
            out: torch.Tensor = None, out_row_stride: int = 0, stats: torch.Tensor = None,
            frame_pitch: int = 0) -> torch.Tensor:
    """20*log10(max(|STFT(audio)|, amin)) for every row of a CUDA tensor.

    audio        [..., n] (mono rows), or with ``channel_axis`` = -2 ([..., C, n], planar) or -1
                 ([..., n, C], interleaved) the channels are averaged first (np.mean semantics).
                 dtype int16 / float32 / float64.
    compute_dtype  float32 or float64; default follows the input like the reference (float64 for
                 float64 input, float32 otherwise).  out_dtype defaults to compute_dtype.
    out / out_row_stride  optional preallocated destination; plane r is written at
                 out.data_ptr() + r*out_row_stride elements (0 = dense), which lets one call per stem
                 fill a [chunks, S, F, T] model input in place.  Then ``out`` itself is returned.
    frame_pitch  elements between consecutive bins (0 = T, dense).  ``aligned_pitch(T)`` (the next multiple of
                 8) keeps every 8-frame store inside one 32-byte sector -- 12 % faster when T is not a
                 multiple of 8 (T = 173, 87); the result is then the [..., :T] view of a [..., F, pitch]
                 buffer (non-contiguous, like the reference's own compute_features output).
    Returns      [..., n_fft/2+1, T] contiguous (T innermost), T = 1 + n // hop.
    The kernel runs on torch's CURRENT stream (the stream the caching allocator hands the temporaries and the
    output out on); use ``with torch.cuda.stream(s):`` to run it elsewhere.
    """
    if not audio.is_cuda:
        raise ValueError('stft_db needs a CUDA tensor (host arrays go through compute_features)')
    dev = audio.device
    n_inner, outer_stride = 0, 0
    nd = audio.dim()
    if channel_axis is None:
        lead, C, n = audio.shape[:-1], 1, audio.shape[-1]
        n_lead, chan_stride, samp_stride = nd - 1, 0, audio.stride(-1)
    else:
        ax = channel_axis % nd
        lead = audio.shape[:-2]
        n_lead = nd - 2
        if ax == nd - 2:
            C, n = audio.shape[-2], audio.shape[-1]
            chan_stride, samp_stride = audio.stride(-2), audio.stride(-1)
        elif ax == nd - 1:
            C, n = audio.shape[-1], audio.shape[-2]
            chan_stride, samp_stride = audio.stride(-1), audio.stride(-2)
        else:
            raise ValueError('channel_axis must be one of the last two axes')
    a = audio
    if n_lead == 0:
        rows, row_stride = 1, 0
    elif n_lead == 1:
        rows, row_stride = audio.shape[0], audio.stride(0)
    elif n_lead == 2 and not audio.is_contiguous():
        # [outer, inner, ...] strided view (e.g. segments x chunks): two-level row addressing
        rows, row_stride = audio.shape[0] * audio.shape[1], audio.stride(1)
        n_inner, outer_stride = audio.shape[1], audio.stride(0)
    else:
        rows = 1
        for d in lead:
            rows *= d
        a = audio.reshape((rows,) + tuple(audio.shape[n_lead:]))
        row_stride = a.stride(0)
        if channel_axis is None:
            samp_stride = a.stride(-1)
        elif ax == nd - 2:
            chan_stride, samp_stride = a.stride(-2), a.stride(-1)
        else:
            chan_stride, samp_stride = a.stride(-1), a.stride(-2)
    if rows <= 1:
        row_stride = 0 if rows == 0 else row_stride
    if compute_dtype is None:
        compute_dtype = torch.float64 if audio.dtype == torch.float64 else torch.float32
    if out_dtype is None:
        out_dtype = compute_dtype
    if input_scale is None:
        input_scale = 1.0 / 32768.0 if audio.dtype == torch.int16 else 1.0
    if window is None:
        window = hann_window(n_fft, dev)
    if window.dtype != torch.float32 or not window.is_cuda or window.numel() != n_fft:
        raise ValueError('window must be a float32 CUDA tensor of n_fft elements')
    window = window.contiguous()
    T = num_frames(n, hop)
    F = n_fft // 2 + 1
    pitch = int(frame_pitch) if frame_pitch else T
    if pitch < T:
        raise ValueError('frame_pitch %d is smaller than the %d frames of a row' % (pitch, T))
    padded = None
    if out is None:
        padded = torch.empty((rows, F, pitch), dtype=out_dtype, device=dev)
        out = padded
    elif out_row_stride:
        if out.dtype != out_dtype or not out.is_cuda or out_row_stride < F * pitch:
            raise ValueError('strided out must be a CUDA %s tensor with out_row_stride >= F*pitch' % out_dtype)
    elif pitch != T:
        if out.dtype != out_dtype or not out.is_contiguous() or out.numel() != rows * F * pitch:
            raise ValueError('out must be a contiguous %s tensor of %d elements' % (out_dtype, rows * F * pitch))
    else:
        if out.dtype != out_dtype or not out.is_contiguous() or out.numel() != rows * F * T:
            raise ValueError('out must be a contiguous %s tensor of %d elements' % (out_dtype, rows * F * T))
    if row_gain is not None:
        row_gain = row_gain.to(device=dev, dtype=torch.float64).contiguous()
        if row_gain.numel() != rows:
            raise ValueError('row_gain needs one entry per row')
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != 3 or not stats.is_cuda):
        raise ValueError('stats must be a CUDA float64 tensor of 3 elements')
    with torch.cuda.device(dev):
        rc = _lib.lib().dam_stft_db(
            _lib.ptr(a), _lib.dtype_code(audio.dtype), C, rows, n,
            row_stride, chan_stride, samp_stride, n_inner, outer_stride, float(input_scale), _lib.ptr(row_gain),
            _lib.ptr(window), n_fft, hop, float(amin), float(multiplier),
            _lib.dtype_code(compute_dtype), _lib.ptr(out), _lib.dtype_code(out_dtype), int(out_row_stride),
            0 if pitch == T else pitch, _lib.ptr(stats), _lib.stream_ptr())
    _lib.check(rc, 'dam_stft_db')
    if out_row_stride:
        return out
    if pitch != T:
        return out.reshape(tuple(lead) + (F, pitch))[..., :T]
    return out.reshape(tuple(lead) + (F, T))


def aligned_pitch(T: int) -> int:
    """Smallest multiple of 8 frames (32 bytes of float32) that holds T frames."""
    return (int(T) + 7) // 8 * 8


def compute_features(audio, window_size: int = 2048, hop_length: int = 1024, device=None):
    """Drop-in for MultitrackAudioDataset.compute_features (data/dataset.py:132-162).

    numpy array / CPU tensor in -> CPU tensor out (host -> device -> kernel -> host), CUDA tensor
    in -> CUDA tensor out.  dtype follows the input as in the reference: float64 -> float64
    arithmetic and output, float32 -> float32.  Leading batch dimensions are kept.
    """
    if isinstance(audio, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(audio))
        host = True
    else:
        t = audio
        host = not audio.is_cuda
    if t.dtype not in (torch.float32, torch.float64, torch.int16):
        raise TypeError('audio must be float32, float64 or int16')
    if host:
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        t = t.to(dev, non_blocking=True)
    out = stft_db(t, window_size, hop_length)
    return out.cpu() if host else out
