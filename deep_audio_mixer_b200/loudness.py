"""BS.1770 integrated loudness on the GPU (kernel K4) behind a pyloudnorm-shaped API.

``Meter(rate).integrated_loudness(data)`` mirrors ``pyloudnorm.Meter`` as the reference uses it
(evaluation.py:32,40,63; data/dataset.py:118,126): data is [N, C] (or [N]).  The heavy part -- the two
biquads over every sample and the block mean squares -- runs in dam_kweight_segments; gating and
the logarithms run on the host over the tiny z[c, j] array, in float64, following
ITU-R BS.1770-4 / pyloudnorm 0.1.0 (absolute gate -70 LUFS, relative gate -10 LU).
"""
import ctypes

import numpy as np
import torch

from . import _lib

G_CHANNEL = (1.0, 1.0, 1.0, 1.41, 1.41)


def _biquad(G, Q, fc, rate, kind):
    """RBJ cookbook biquad normalised by a0 (pyloudnorm.IIRfilter.generate_coefficients)."""
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    cw = np.cos(w0)
    if kind == 'high_shelf':
        b = [A * ((A + 1) + (A - 1) * cw + 2 * np.sqrt(A) * alpha),
             -2 * A * ((A - 1) + (A + 1) * cw),
             A * ((A + 1) + (A - 1) * cw - 2 * np.sqrt(A) * alpha)]
        a = [(A + 1) - (A - 1) * cw + 2 * np.sqrt(A) * alpha,
             2 * ((A - 1) - (A + 1) * cw),
             (A + 1) - (A - 1) * cw - 2 * np.sqrt(A) * alpha]
    else:
        b = [(1 + cw) / 2, -(1 + cw), (1 + cw) / 2]
        a = [1 + alpha, -2 * cw, 1 - alpha]
    b, a = np.array(b) / a[0], np.array(a) / a[0]
    return [b[0], b[1], b[2], a[1], a[2]]


class Meter:
    def __init__(self, rate, filter_class='K-weighting', block_size=0.400, device=None):
        if filter_class != 'K-weighting':
            raise ValueError('only the K-weighting filter class is implemented')
        self.rate = rate
        self.block_size = block_size
        self.device = device
        self._coeffs = (_biquad(4.0, 1 / np.sqrt(2), 1500.0, rate, 'high_shelf') +
                        _biquad(0.0, 0.5, 38.0, rate, 'high_pass'))
        self._bounds_cache = {}
        self._ws_cache = {}          # (R, n, device, stream) -> (segment sums, filter-state workspace), reused per shape

    # -- gating-block geometry, evaluated in floating point exactly like pyloudnorm does
    def _geometry(self, n, dev):
        key = (n, str(dev))
        hit = self._bounds_cache.get(key)
        if hit is not None:
            return hit
        T_g, step = self.block_size, 0.25
        T = n / self.rate
        num_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
        lo = np.array([int(T_g * (j * step) * self.rate) for j in range(num_blocks)], dtype=np.int64)
        hi = np.array([int(T_g * (j * step + 1) * self.rate) for j in range(num_blocks)], dtype=np.int64)
        hi = np.minimum(hi, n)
        bounds = np.unique(np.concatenate([lo, hi]))
        first = np.searchsorted(bounds, lo)
        last = np.searchsorted(bounds, hi)           # block j = segments [first, last)
        geo = (torch.from_numpy(bounds).to(dev), len(bounds) - 1, first, last, num_blocks,
               torch.from_numpy(first.astype(np.int64)).to(dev), torch.from_numpy(last.astype(np.int64)).to(dev))
        if len(self._bounds_cache) > 64:
            self._bounds_cache.clear()
        self._bounds_cache[key] = geo
        return geo

    def block_energies(self, rows: torch.Tensor, kw_workspace=None) -> np.ndarray:
        """rows [R, N] CUDA (float32/float64) -> z [R, num_blocks] float64 (host): mean square of the
        K-weighted signal over each 400 ms gating block.  ``kw_workspace``: see _segment_sums."""
        if not rows.is_cuda:
            raise ValueError('block_energies needs a CUDA tensor')
        R, n = rows.shape
        if n < self.block_size * self.rate:
            raise ValueError('Audio must have length greater than the block size.')
        dev = rows.device
        seg, first, last = self._segment_sums(rows, kw_workspace)
        csum = np.concatenate([np.zeros((R, 1)), np.cumsum(seg.cpu().numpy(), axis=1)], axis=1)
        return (csum[:, last] - csum[:, first]) / (self.block_size * self.rate)

    def block_energies_device(self, rows: torch.Tensor) -> torch.Tensor:
        """rows [R, N] CUDA -> z [R, num_blocks] float64 on the device through the one-call entry point
        dam_kweight_blocks (the library evaluates pyloudnorm's block geometry itself); asynchronous."""
        if not rows.is_cuda:
            raise ValueError('block_energies_device needs a CUDA tensor')
        R, n = rows.shape
        if n < self.block_size * self.rate:
            raise ValueError('Audio must have length greater than the block size.')
        dev = rows.device
        lib = _lib.lib()
        nb = lib.dam_kweight_num_blocks(n, float(self.rate), float(self.block_size), 0.75)
        ws_bytes = lib.dam_kweight_blocks_workspace_bytes(R, n, float(self.rate), float(self.block_size), 0.75)
        ws = torch.empty(ws_bytes // 8 + 2, dtype=torch.float64, device=dev)
        z = torch.empty((R, nb), dtype=torch.float64, device=dev)
        coeffs = (ctypes.c_double * 10)(*self._coeffs)
        with torch.cuda.device(dev):
            rc = lib.dam_kweight_blocks(_lib.ptr(rows), _lib.dtype_code(rows.dtype), R, n, rows.stride(0), rows.stride(1),
                                        float(self.rate), coeffs, float(self.block_size), 0.75, _lib.ptr(z), _lib.ptr(ws),
                                        ws_bytes, _lib.stream_ptr())
        _lib.check(rc, 'dam_kweight_blocks')
        return z

    def _workspace(self, R, n, n_seg, dev):
        """Scratch of one call, kept per (shape, device, stream): small songs otherwise spend more time in two
        allocator calls and a memset launch than in the kernels.  Calls are ordered on the stream, and every caller
        consumes the segment sums before the next call of the same shape is enqueued."""
        key = (R, n, n_seg, str(dev), torch.cuda.current_stream(dev).cuda_stream)
        hit = self._ws_cache.get(key)
        if hit is None:
            ws_bytes = _lib.lib().dam_kweight_workspace_bytes(R, n)
            hit = (torch.empty((R, n_seg), dtype=torch.float64, device=dev),
                   torch.empty(ws_bytes // 8 + 2, dtype=torch.float64, device=dev), ws_bytes)
            if len(self._ws_cache) > 32:
                self._ws_cache.clear()
            self._ws_cache[key] = hit
        return hit

    def workspace_for(self, R, n, dev):
        """(segment sums, filter-state workspace, its size in bytes) of one call over R rows of n samples: cached per
        shape and stream, or fresh from the graph's private pool inside a CUDA-graph capture."""
        if n < self.block_size * self.rate:
            raise ValueError('Audio must have length greater than the block size.')
        n_seg = self._geometry(n, dev)[1]
        if torch.cuda.is_current_stream_capturing():
            ws_bytes = _lib.lib().dam_kweight_workspace_bytes(R, n)
            return (torch.empty((R, n_seg), dtype=torch.float64, device=dev),
                    torch.empty(ws_bytes // 8 + 2, dtype=torch.float64, device=dev), ws_bytes)
        return self._workspace(R, n, n_seg, dev)

    def _segment_sums(self, rows: torch.Tensor, kw_workspace=None):
        """kw_workspace: what mixing.gain_mix(..., meter=self) returned -- the chunk states of ``rows`` are already
        in it, so the call resumes at the scan (dam_kweight_segments_resume) and reads ``rows`` once."""
        R, n = rows.shape
        dev = rows.device
        bounds, n_seg, first, last, num_blocks = self._geometry(n, dev)[:5]
        lib = _lib.lib()
        resume = kw_workspace is not None
        if resume:
            seg, ws, ws_bytes, ws_rows, ws_n, ws_ptr = kw_workspace
            if (ws_rows, ws_n, ws_ptr) != (R, n, rows.data_ptr()) or rows.stride(1) != 1 or rows.stride(0) != n:
                raise ValueError('kw_workspace belongs to another mix (gain_mix(..., meter=) returns the pair)')
        else:
            seg, ws, ws_bytes = self.workspace_for(R, n, dev)
            seg.zero_()                      # dam_kweight_segments ADDS; the resumed form clears its output itself
        coeffs = (ctypes.c_double * 10)(*self._coeffs)
        entry = lib.dam_kweight_segments_resume if resume else lib.dam_kweight_segments
        with torch.cuda.device(dev):
            rc = entry(_lib.ptr(rows), _lib.dtype_code(rows.dtype), R, n,
                       rows.stride(0), rows.stride(1), coeffs, _lib.ptr(bounds), n_seg,
                       _lib.ptr(seg), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, 'dam_kweight_segments')
        return seg, first, last

    def integrated_loudness_device(self, signals: torch.Tensor, kw_workspace=None, gain_target: float = None):
        """signals [R, C, N] CUDA -> float64 CUDA tensor [R] of LUFS values; K-weighting, block sums
        and gating all run on the device with no host synchronisation.  ``kw_workspace``: see _segment_sums.
        With ``gain_target`` (LUFS) returns (lufs, gain): gain = 10**((gain_target - lufs) / 20), the factor of
        pyln.normalize.loudness / evaluation.py:64, from the same launch."""
        R, C, n = signals.shape
        if n < self.block_size * self.rate:
            raise ValueError('Audio must have length greater than the block size.')
        rows = signals.reshape(R * C, n)
        if rows.stride(1) != 1:
            if kw_workspace is not None:
                raise ValueError('kw_workspace needs the contiguous mix gain_mix returned')
            rows = rows.contiguous()
        seg, _, _ = self._segment_sums(rows, kw_workspace)
        geo = self._geometry(n, rows.device)
        out = torch.empty((2, R) if gain_target is not None else (R,), dtype=torch.float64, device=rows.device)
        lib = _lib.lib()
        with torch.cuda.device(rows.device):
            if gain_target is None:
                rc = lib.dam_gated_loudness(_lib.ptr(seg), R, C, geo[1], _lib.ptr(geo[5]), _lib.ptr(geo[6]),
                                            geo[4], 1.0 / (self.block_size * self.rate), _lib.ptr(out), _lib.stream_ptr())
            else:
                rc = lib.dam_gated_loudness_gain(_lib.ptr(seg), R, C, geo[1], _lib.ptr(geo[5]), _lib.ptr(geo[6]),
                                                 geo[4], 1.0 / (self.block_size * self.rate), float(gain_target),
                                                 _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.stream_ptr())
        _lib.check(rc, 'dam_gated_loudness')
        return out if gain_target is None else (out[0], out[1])

    @staticmethod
    def gated_loudness(z: np.ndarray) -> float:
        """z [C, J] block energies -> LUFS with absolute and relative gating."""
        C = z.shape[0]
        G = np.asarray(G_CHANNEL[:C])[:, None]
        with np.errstate(divide='ignore', invalid='ignore'):
            l = -0.691 + 10.0 * np.log10(np.sum(G * z, axis=0))
            keep = l >= -70.0
            z_avg = z[:, keep].mean(axis=1) if keep.any() else np.full(C, np.nan)
            gamma_r = -0.691 + 10.0 * np.log10(np.sum(G[:, 0] * z_avg)) - 10.0
            keep = (l > gamma_r) & (l > -70.0)
            z_avg = np.nan_to_num(z[:, keep].mean(axis=1) if keep.any() else np.full(C, np.nan))
            return float(-0.691 + 10.0 * np.log10(np.sum(G[:, 0] * z_avg)))

    def _rows(self, data):
        """[N, C] / [N] host or device data -> ([C, N] CUDA rows with unit sample stride)."""
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(np.ascontiguousarray(data.T if data.ndim == 2 else data[None]))
            dev = self.device or torch.device('cuda', torch.cuda.current_device())
            rows = data.to(dev)
        else:
            rows = data.t() if data.dim() == 2 else data[None]
            if not rows.is_cuda:
                rows = rows.to(self.device or torch.device('cuda', torch.cuda.current_device()))
        if rows.dtype not in (torch.float32, torch.float64):
            rows = rows.double()
        if rows.stride(1) != 1:
            rows = rows.contiguous()
        return rows

    def integrated_loudness(self, data, kw_workspace=None) -> float:
        """pyln.Meter.integrated_loudness(data) for data [N, C] or [N].  ``kw_workspace``: ``data`` is the transpose of
        a mix that mixing.gain_mix / stem_sum produced with ``meter=self`` (its first pass is already done)."""
        return self.gated_loudness(self.block_energies(self._rows(data), kw_workspace))

    def integrated_loudness_batch(self, signals: torch.Tensor):
        """signals [R, C, N] CUDA -> list of R loudness values, one kernel pass for all of them."""
        R, C, n = signals.shape
        rows = signals.reshape(R * C, n)
        if rows.stride(1) != 1:
            rows = rows.contiguous()
        z = self.block_energies(rows).reshape(R, C, -1)
        return [self.gated_loudness(z[r]) for r in range(R)]
