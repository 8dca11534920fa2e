"""Drive the UNMODIFIED reference hot path from ``/root/reference`` (build container only).

Only used to (a) generate ``tests/golden/*.npz`` (``tests/golden/make_golden.py``)
and (b) pin the C oracle against the live reference in ``-m "not gpu"`` tests.
The GPU box has no ``/root/reference``: everything here is skipped there.

The single stub: ``smplpytorch.native.webuser.serialization.ready_arguments``
(needs chumpy + licensed .pkl, `serialization.py:1-39`) is replaced by the
synthetic model provider; `SMPL_Layer.__init__` reads only the attributes the
stub supplies (`smpl_layer.py:40-61`).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get('POSERISK_REFERENCE', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'lib', 'utils', 'reba.py'))


class _R:
    def __init__(self, a):
        self.r = np.asarray(a, dtype=np.float64)


_loaded = None


def load_reference():
    """Returns a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError('reference tree not present')
    import scipy.sparse as sp
    for p in ('lib/utils', 'lib/smplpytorch'):
        full = os.path.join(REF_ROOT, p)
        if full not in sys.path:
            sys.path.insert(0, full)
    import smplpytorch.pytorch.smpl_layer as ref_smpl_layer  # noqa
    import reba as ref_reba  # noqa
    import rula as ref_rula  # noqa
    import coord_utils as ref_coord  # noqa

    from poserisk_release_b200.model_provider import synthetic_smpl, GENDER_FILE

    def fake_ready_arguments(path):
        gender = {v: k for k, v in GENDER_FILE.items()}[os.path.basename(path)]
        m = synthetic_smpl(gender)
        return {
            'betas': _R(m.betas), 'shapedirs': _R(m.shapedirs), 'posedirs': _R(m.posedirs),
            'v_template': _R(m.v_template), 'weights': _R(m.weights),
            'J_regressor': sp.csc_matrix(m.J_regressor.astype(np.float64)),
            'f': m.faces.astype(np.uint32), 'kintree_table': m.kintree_table,
        }

    ref_smpl_layer.ready_arguments = fake_ready_arguments
    ns = types.SimpleNamespace(
        SMPL_Layer=ref_smpl_layer.SMPL_Layer, REBA=ref_reba.REBA, RULA=ref_rula.RULA,
        coord_utils=ref_coord, root=REF_ROOT)
    _loaded = ns
    return ns


def ref_log_to_parts(results, which):
    """Flatten reference ``log_score`` lists to a uint8 (N, 9|11) array + int scores."""
    n_parts = 9 if which == 'REBA' else 11
    parts = np.zeros((len(results), n_parts), np.int64)
    scores = np.zeros(len(results), np.int64)
    for i, r in enumerate(results):
        flat = []
        for item in r['log_score']:
            if isinstance(item, str):
                flat.extend(int(x) for x in item.split(','))
            else:
                flat.append(int(item))
        parts[i] = flat
        scores[i] = int(r['score'])
    return scores, parts
