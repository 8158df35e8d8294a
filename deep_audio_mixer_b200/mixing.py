"""Time-domain mix kernels (K2, N2, K5) and feature-domain ops (K3, K6) behind the C ABI."""
import ctypes

import numpy as np
import torch

from . import _lib


def _check_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError('%s must be a CUDA tensor' % name)


def gain_mix(stems: torch.Tensor, gains, interp_coef: int = None, hold_single: bool = False,
             compute_dtype: torch.dtype = torch.float64, out_dtype: torch.dtype = None,
             want_mix: bool = True, want_stems: bool = False, want_stats: bool = False, meter=None):
    """Zero-order-hold gain masks x stems, summed over stems (kernel K2).

    stems [S, C, N] (or a batch [B, S, C, N]) CUDA float32/float64 planar audio; gains [S, K]
    ([B, S, K]) array-like or tensor, amplitude
    gains.  mask_s[i] = gains[s][min(i // interp_coef, K-1)], interp_coef = int(N / K) by default
    (inference_utils.py:26); K == 1 gives an all-zero mask unless hold_single (the reference quirk).
    Returns dict(mix=[C, N], stems=[S, C, N], peak=[C], sumsq=[C]) for the requested outputs.  Runs on torch's
    current stream.  Device gains are used in place (no copy, no synchronisation); host gains go through a small
    ring of pinned staging buffers and an asynchronous copy.

    ``meter`` (a loudness.Meter): the mix is about to be measured -- the kernel also leaves the meter's first pass
    over the mix (K-weighting chunk states) in a workspace, returned as ``kw_workspace``; hand it to
    ``meter.integrated_loudness_device(mix, kw_workspace=...)``, which then reads the mix once instead of twice
    (same bits as measuring the mix from scratch).
    """
    _check_cuda(stems, 'stems')
    batched = stems.dim() == 4
    if stems.dim() not in (3, 4):
        raise ValueError('stems must be [S, C, N] or [B, S, C, N]')
    if stems.stride(-1) != 1:
        stems = stems.contiguous()
    if not batched:
        stems = stems.unsqueeze(0)
    B, S, C, N = stems.shape
    dev = stems.device
    if isinstance(gains, torch.Tensor) and gains.is_cuda:
        g = gains.to(dtype=torch.float64).reshape(B, S, -1).contiguous()         # no-ops when already in that form
    else:
        g = _lib.upload_small(np.asarray(gains.cpu() if isinstance(gains, torch.Tensor) else gains, dtype=np.float64), dev)
        g = g.reshape(B, S, -1)
    K = g.shape[2]
    assert K <= N or N == 0, "Target mask should be longer than the initial one"
    if interp_coef is None:
        interp_coef = int(N / K)
    if out_dtype is None:
        out_dtype = compute_dtype
    mix = torch.empty((B, C, N), dtype=out_dtype, device=dev) if want_mix else None
    st_out = torch.empty((B, S, C, N), dtype=out_dtype, device=dev) if want_stems else None
    stats = torch.zeros((2, B, C), dtype=torch.float64, device=dev) if want_stats else None     # one fill for both
    peak = stats[0] if want_stats else None
    sumsq = stats[1] if want_stats else None
    common = (_lib.ptr(stems), _lib.dtype_code(stems.dtype), B, stems.stride(0) if B > 1 else 0,
              S, C, N, stems.stride(1), stems.stride(2),
              _lib.ptr(g), K, int(interp_coef), int(bool(hold_single)), _lib.dtype_code(compute_dtype),
              _lib.ptr(mix), _lib.ptr(st_out), _lib.dtype_code(out_dtype),
              _lib.ptr(peak), _lib.ptr(sumsq))
    kw_workspace = None
    with torch.cuda.device(dev):
        if meter is not None:
            if mix is None:
                raise ValueError('gain_mix: meter needs want_mix=True')
            kw_workspace = meter.workspace_for(B * C, N, dev)
            coeffs = (ctypes.c_double * 10)(*meter._coeffs)
            rc = _lib.lib().dam_gain_mix_kweight(*common, coeffs, _lib.ptr(kw_workspace[1]), kw_workspace[2],
                                                 _lib.stream_ptr())
        else:
            rc = _lib.lib().dam_gain_mix(*common, _lib.stream_ptr())
    _lib.check(rc, 'dam_gain_mix')
    if not batched:
        mix, st_out, peak, sumsq = [None if t is None else t[0] for t in (mix, st_out, peak, sumsq)]
    if kw_workspace is not None:
        kw_workspace = tuple(kw_workspace) + (B * C, N, mix.data_ptr())          # what the states belong to
    return dict(mix=mix, stems=st_out, peak=peak, sumsq=sumsq, kw_workspace=kw_workspace)


def head_gains(m: torch.Tensor, n_batch: int, n_chunks: int) -> torch.Tensor:
    """CNN head outputs m [n_batch * n_chunks, S] float32 (row = chunk c of song b) -> amplitude gains
    [n_batch, S, n_chunks] float64 for ``gain_mix``: 10 ** (0.5 m) in float32 (data/dataset_utils.py:46-50), widened.
    One launch (dam_head_gains); the same bits as torch.pow(10.0, 0.5 * m).double() transposed."""
    _check_cuda(m, 'm')
    if m.dim() != 2 or m.shape[0] != n_batch * n_chunks:
        raise ValueError('m must be [n_batch * n_chunks, S]')
    m = m.float().contiguous()
    S = m.shape[1]
    out = torch.empty((n_batch, S, n_chunks), dtype=torch.float64, device=m.device)
    with torch.cuda.device(m.device):
        rc = _lib.lib().dam_head_gains(_lib.ptr(m), n_batch, n_chunks, S, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, 'dam_head_gains')
    return out


def stem_sum(stems: torch.Tensor, out_dtype=None, meter=None):
    """np.sum(np.array(list(d.values())), axis=0) (evaluation.py:59-60) for stems [S, C, N].  With ``meter`` (the sum is
    measured next, evaluation.py:63) returns (sum, kw_workspace) for ``meter.integrated_loudness(sum.t(), kw_workspace)``."""
    ones = torch.ones(tuple(stems.shape[:-2]) + (1,), dtype=torch.float64, device=stems.device)
    res = gain_mix(stems, ones, hold_single=True, compute_dtype=stems.dtype, out_dtype=out_dtype or stems.dtype, meter=meter)
    return res['mix'] if meter is None else (res['mix'], res['kw_workspace'])


def bus_sum(stems: torch.Tensor, group, n_bus: int = 4, batch_dims: int = 0,
            input_scale: float = 1.0 / 32768.0, want_mono: bool = False):
    """S raw stems -> n_bus buses, summed in stem order in the audio dtype (kernel N2;
    data/medleydb_data_surgery.ipynb:305-332).  stems [S, ...] (or [B, S, ...] with batch_dims=1);
    group[s] in [0, n_bus) or -1 (dropped).  int16 stems (PCM16) are converted with x * input_scale and
    summed in float32.  ``want_mono`` (stems [.., S, C, N] float32 / int16) also returns the channel mean
    of every bus, [.., n_bus, N] float32 -- the mono rows K1 featurises -- as (buses, mono)."""
    _check_cuda(stems, 'stems')
    if batch_dims:
        B, S = stems.shape[0], stems.shape[1]
        flat = stems.reshape(B, S, -1)
        tail = tuple(stems.shape[2:])
    else:
        B, S = 1, stems.shape[0]
        flat = stems.reshape(1, S, -1)
        tail = tuple(stems.shape[1:])
    if flat.stride(2) != 1:
        flat = flat.contiguous()
    out_dtype = torch.float32 if stems.dtype == torch.int16 else stems.dtype
    out = torch.empty(((B,) if batch_dims else ()) + (n_bus,) + tail, dtype=out_dtype, device=stems.device)
    grp = (ctypes.c_int32 * S)(*[int(v) for v in group])
    mono, C = None, 1
    if want_mono:
        if len(tail) != 2 or stems.dtype not in (torch.float32, torch.int16):
            raise ValueError('want_mono needs [.., S, C, N] float32 or int16 stems')
        C = tail[0]
        mono = torch.empty(((B,) if batch_dims else ()) + (n_bus, tail[1]), dtype=torch.float32, device=stems.device)
    with torch.cuda.device(stems.device):
        rc = _lib.lib().dam_bus_sum(_lib.ptr(flat), _lib.dtype_code(stems.dtype), B, flat.stride(0) if B > 1 else 0,
                                    S, flat.shape[2], flat.stride(1), grp, n_bus, float(input_scale),
                                    _lib.ptr(out), _lib.ptr(mono), C, _lib.stream_ptr())
    _lib.check(rc, 'dam_bus_sum')
    return (out, mono) if want_mono else out


def pcm_unpack(raw: torch.Tensor, bits: int, out_dtype=torch.float64, shape=None) -> torch.Tensor:
    """Packed little-endian PCM bytes (uint8 CUDA tensor) -> float samples scaled by 2**-(bits-1), the values
    soundfile.read returns (N1: 24-bit WAV data decoded on the device, 3 bytes per sample over PCIe)."""
    _check_cuda(raw, 'raw')
    if raw.dtype != torch.uint8 or not raw.is_contiguous():
        raise ValueError('raw must be a contiguous uint8 tensor')
    n = raw.numel() // (bits // 8)
    out = torch.empty(n, dtype=out_dtype, device=raw.device)
    with torch.cuda.device(raw.device):
        rc = _lib.lib().dam_pcm_unpack(_lib.ptr(raw), int(bits), n, 1.0 / float(1 << (bits - 1)), _lib.ptr(out),
                                       _lib.dtype_code(out_dtype), _lib.stream_ptr())
    _lib.check(rc, 'dam_pcm_unpack')
    return out.reshape(shape) if shape is not None else out


def row_reduce(x: torch.Tensor, want_absmax=True, want_sumsq=True):
    """Per-row max|x| and sum x^2 (float64) of a [R, n] CUDA tensor."""
    _check_cuda(x, 'x')
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    R, n = x2.shape
    absmax = torch.zeros(R, dtype=torch.float64, device=x.device) if want_absmax else None
    sumsq = torch.zeros(R, dtype=torch.float64, device=x.device) if want_sumsq else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().dam_row_reduce(_lib.ptr(x2), _lib.dtype_code(x.dtype), R, n, x2.stride(0),
                                       _lib.ptr(absmax), _lib.ptr(sumsq), _lib.stream_ptr())
    _lib.check(rc, 'dam_row_reduce')
    return absmax, sumsq


def scale_rows(x: torch.Tensor, factor, divide: bool = False, out_dtype=None) -> torch.Tensor:
    """x[r, :] * factor[r] (or / factor[r]) in float64, rounded to out_dtype (kernel K5)."""
    _check_cuda(x, 'x')
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    R, n = x2.shape
    f = torch.as_tensor(factor, dtype=torch.float64).reshape(-1).to(x.device)
    if f.numel() == 1 and R > 1:
        f = f.expand(R)
    f = f.contiguous()
    out_dtype = out_dtype or x.dtype
    out = torch.empty((R, n), dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dam_scale_rows(_lib.ptr(x2), _lib.dtype_code(x.dtype), R, n, x2.stride(0),
                                       _lib.ptr(f), int(bool(divide)), _lib.ptr(out),
                                       _lib.dtype_code(out_dtype), _lib.stream_ptr())
    _lib.check(rc, 'dam_scale_rows')
    return out.reshape(x.shape)


def scale(x: torch.Tensor, gain: float, out_dtype=None) -> torch.Tensor:
    """x * gain for a host scalar gain (C entry point dam_scale): float64 product rounded to out_dtype.
    No upload of the factor, so the call never touches the host allocator."""
    _check_cuda(x, 'x')
    xc = x if x.is_contiguous() else x.contiguous()
    out_dtype = out_dtype or x.dtype
    out = torch.empty(xc.shape, dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dam_scale(_lib.ptr(xc), _lib.dtype_code(x.dtype), xc.numel(), float(gain), _lib.ptr(out),
                                  _lib.dtype_code(out_dtype), _lib.stream_ptr())
    _lib.check(rc, 'dam_scale')
    return out


def normalize_loudness(data: torch.Tensor, input_loudness: float, target_loudness: float,
                       out_dtype=torch.float64) -> torch.Tensor:
    """pyloudnorm.normalize.loudness: ``np.power(10.0, (target - input) / 20.0) * data`` (evaluation.py:64,
    mean_loudness_model.py:15-17, listening_test_data_preparation.py:23).  The gain is a numpy float64 scalar, so
    under numpy >= 2 (the version the goldens were generated with) the product -- and the result -- is float64
    whatever the dtype of ``data``; every drop-in of that expression here (sum_and_normalize, produce_mixture,
    MeanLoudnessModel) follows this one rule."""
    gain = float(np.power(10.0, (target_loudness - input_loudness) / 20.0))
    return scale(data, gain, out_dtype=out_dtype)


def peak_normalize(x: torch.Tensor) -> torch.Tensor:
    """librosa.util.normalize(x, axis=1) for [R, n]: each row divided by its max |x|
    (inference.ipynb:219); rows with a peak below the dtype's tiny are left unscaled."""
    absmax, _ = row_reduce(x, True, False)
    tiny = torch.finfo(x.dtype).tiny
    absmax = torch.where(absmax < tiny, torch.ones_like(absmax), absmax)
    return scale_rows(x, absmax, divide=True)


class _FeatureMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, m):
        B, S = x.shape[0], x.shape[1]
        P = x[0, 0].numel()
        xc = x.contiguous()
        mc = m.contiguous()
        out = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().dam_feature_mix_fwd(_lib.ptr(xc), _lib.ptr(mc), B, S, P, _lib.ptr(out),
                                                _lib.stream_ptr())
        _lib.check(rc, 'dam_feature_mix_fwd')
        ctx.save_for_backward(xc, mc)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xc, mc = ctx.saved_tensors
        B, S = xc.shape[0], xc.shape[1]
        P = xc[0, 0].numel()
        go = grad_out.contiguous()
        grad_m = torch.zeros_like(mc)
        grad_x = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(xc.device):
            rc = _lib.lib().dam_feature_mix_bwd(_lib.ptr(xc), _lib.ptr(mc), _lib.ptr(go), B, S, P,
                                                _lib.ptr(grad_m), _lib.ptr(grad_x), _lib.stream_ptr())
        _lib.check(rc, 'dam_feature_mix_bwd')
        return grad_x, grad_m


def feature_mix(x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """masked[b] = sum_i m[b, i] * x[b, i] with autograd (kernel K3; model_scalar_1s.py:260-275).
    x [B, S, F, T] float32 CUDA, m [B, S] float32."""
    _check_cuda(x, 'x')
    if x.dtype != torch.float32 or m.dtype != torch.float32:
        raise TypeError('feature_mix is float32 (the reference raises on float64 input too)')
    return _FeatureMix.apply(x, m)


def db_sum(x: torch.Tensor) -> torch.Tensor:
    """DummyModel.forward: 20*log10(sum_i 10**(0.5*x_i)) (kernel K6; dummy_model.py:23-34)."""
    _check_cuda(x, 'x')
    xc = x.contiguous().float()
    B, S = xc.shape[0], xc.shape[1]
    P = xc[0, 0].numel()
    out = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().dam_db_sum(_lib.ptr(xc), B, S, P, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, 'dam_db_sum')
    return out
