"""ctypes binding of libdam_b200.so (include/dam_b200.h).  There is no CPU fallback: if the
library is missing or a call fails, this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DAM_LIB_VARIANT=<name> loads libdam_b200_<name>.so (built by `build.py --variant`): kernel A/B runs only
LIB_PATH = os.path.join(_HERE, 'libdam_b200%s.so' % ('_' + os.environ['DAM_LIB_VARIANT'] if os.environ.get('DAM_LIB_VARIANT') else ''))

DAM_I16, DAM_F32, DAM_F64 = 0, 1, 2

_lib = None

_c = ctypes
_vp, _i, _i64, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double

# name -> (restype, argtypes); must list every symbol include/dam_b200.h declares
SIGNATURES = {
    'dam_version': (_c.c_char_p, []),
    'dam_last_error': (_c.c_char_p, []),
    'dam_init': (_i, []),
    'dam_launch_count': (_i64, []),
    'dam_last_stft_kernel': (_c.c_char_p, []),
    'dam_kweight_init': (_i, [_c.POINTER(_d)]),
    'dam_stft_num_frames': (_i64, [_i64, _i]),
    'dam_stft_db': (_i, [_vp, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _d, _vp, _vp, _i, _i, _d, _d, _i,
                         _vp, _i, _i64, _i64, _vp, _vp]),
    'dam_gain_mix': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp, _i,
                          _vp, _vp, _vp]),
    'dam_gain_mix_kweight': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp, _i,
                                  _vp, _vp, _c.POINTER(_d), _vp, _c.c_size_t, _vp]),
    'dam_bus_sum': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _i64, _c.POINTER(_c.c_int32), _i, _d, _vp, _vp, _i, _vp]),
    'dam_feature_mix_fwd': (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    'dam_feature_mix_bwd': (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    'dam_db_sum': (_i, [_vp, _i64, _i64, _i64, _vp, _vp]),
    'dam_scale_rows': (_i, [_vp, _i, _i64, _i64, _i64, _vp, _i, _vp, _i, _vp]),
    'dam_row_reduce': (_i, [_vp, _i, _i64, _i64, _i64, _vp, _vp, _vp]),
    'dam_kweight_workspace_bytes': (_c.c_size_t, [_i64, _i64]),
    'dam_gated_loudness': (_i, [_vp, _i64, _i, _i64, _vp, _vp, _i64, _d, _vp, _vp]),
    'dam_gated_loudness_gain': (_i, [_vp, _i64, _i, _i64, _vp, _vp, _i64, _d, _d, _vp, _vp, _vp]),
    'dam_head_gains': (_i, [_vp, _i64, _i64, _i64, _vp, _vp]),
    'dam_kweight_segments': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _c.POINTER(_d), _vp, _i64, _vp, _vp,
                                  _c.c_size_t, _vp]),
    'dam_kweight_segments_resume': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _c.POINTER(_d), _vp, _i64, _vp, _vp,
                                         _c.c_size_t, _vp]),
    'dam_scale': (_i, [_vp, _i, _i64, _d, _vp, _i, _vp]),
    'dam_pcm_unpack': (_i, [_vp, _i, _i64, _d, _vp, _i, _vp]),
    'dam_kweight_num_blocks': (_i64, [_i64, _d, _d, _d]),
    'dam_kweight_blocks_workspace_bytes': (_c.c_size_t, [_i64, _i64, _d, _d, _d]),
    'dam_kweight_blocks': (_i, [_vp, _i, _i64, _i64, _i64, _i64, _d, _c.POINTER(_d), _d, _d, _vp, _vp, _c.c_size_t, _vp]),
}


class DamError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DamError('%s not found: build it with `python -m deep_audio_mixer_b200.build` '
                           '(there is no CPU fallback)' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().dam_last_error().decode('utf-8', 'replace')
        raise DamError('%s failed with status %d: %s' % (what, status, msg))


def dtype_code(torch_dtype):
    import torch
    table = {torch.int16: DAM_I16, torch.float32: DAM_F32, torch.float64: DAM_F64}
    if torch_dtype not in table:
        raise TypeError('unsupported dtype %s (int16, float32, float64)' % torch_dtype)
    return table[torch_dtype]


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# ---- small host -> device uploads (gain tables, factors): a ring of pinned staging buffers, asynchronous copies.
# A pageable source would make every cudaMemcpyAsync synchronous (and each call would stall the stream it is on).
_RING = {}
_RING_SLOTS, _RING_BYTES = 8, 1 << 16


def upload_small(arr, device):
    """numpy array -> device tensor of the same dtype / shape, on torch's current stream, without a pageable copy.
    Arrays above 64 KB fall back to a one-off pinned buffer."""
    import numpy as np
    import torch
    arr = np.ascontiguousarray(arr)
    nbytes = arr.nbytes
    dev = torch.device(device)
    if nbytes == 0:
        return torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype, device=dev)
    if nbytes > _RING_BYTES:
        return torch.from_numpy(arr).pin_memory().to(dev, non_blocking=True)
    ring = _RING.get(dev)
    if ring is None:
        ring = {'bufs': [torch.empty(_RING_BYTES, dtype=torch.uint8).pin_memory() for _ in range(_RING_SLOTS)],
                'events': [None] * _RING_SLOTS, 'next': 0}
        _RING[dev] = ring
    i = ring['next']
    ring['next'] = (i + 1) % _RING_SLOTS
    if ring['events'][i] is not None:
        ring['events'][i].synchronize()          # the copy that last used this slot has finished (8 uploads ago)
    stage = ring['bufs'][i][:nbytes]
    stage.numpy()[:] = arr.reshape(-1).view(np.uint8)
    out = stage.to(dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    ring['events'][i] = ev
    return out.view(torch.from_numpy(arr).dtype).reshape(arr.shape)
