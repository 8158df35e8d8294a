"""Import the UNMODIFIED reference from /root/reference (build container only).

The reference imports ``soundfile``, ``pyloudnorm``, ``librosa`` (and ``openpyxl`` in
evaluation.py) at module top; none is installed and there is no network.  None of them is
touched by the functions we call (compute_features, the models, interpolate_mask,
mix_song_smooth), so empty stub modules are enough.  Nothing here is used at test/bench time on
the GPU box: it only feeds ``tools/gen_golden.py``, whose outputs are committed.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('DAM_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'data'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference modules."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)

    class _Meter:                       # never called on the paths we exercise
        def __init__(self, *a, **k):
            pass

    _stub('soundfile')
    pyln = _stub('pyloudnorm', Meter=_Meter)
    pyln.normalize = _stub('pyloudnorm.normalize')
    _stub('librosa')
    _stub('openpyxl', Workbook=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import data.dataset as ref_dataset
    import data.dataset_utils as ref_dataset_utils
    import inference_utils as ref_inference
    import models.model_scalar_1s as ref_m1
    import models.model_scalar_2s as ref_m2
    import models.model_resnet as ref_resnet
    import models.baselines.dummy_model as ref_dummy
    ns = types.SimpleNamespace(dataset=ref_dataset, dataset_utils=ref_dataset_utils,
                               inference=ref_inference, m1=ref_m1, m2=ref_m2,
                               resnet=ref_resnet, dummy=ref_dummy)
    return ns
