"""SMPL body-model constants: synthetic SMPL-shaped provider + real ``.pkl`` loader.

The reference loads ``SMPL_{NEUTRAL,FEMALE,MALE}.pkl`` through chumpy
(`lib/smplpytorch/smplpytorch/native/webuser/serialization.py:1-39`) and
``SMPL_Layer.__init__`` keeps seven arrays from it
(`lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:40-63`).  The licensed
``.pkl`` files are not redistributable, so benchmarks/tests use a synthetic
model with the same tensor shapes (SURVEY.md §8d recipe).  A real pickle, when
present, is read without chumpy by :func:`load_smpl_pkl`.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEATS = 207  # 23 joints x 9 rotation entries (tensutils.py:41-48)

# kintree_table[0] of every SMPL model (smpl_layer.py:60-63); parents[0] is the
# uint32 wrap of -1 and is never read.
SMPL_PARENTS = (4294967295, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14,
                16, 17, 18, 19, 20, 21)

GENDER_FILE = {'neutral': 'SMPL_NEUTRAL.pkl', 'female': 'SMPL_FEMALE.pkl',
               'male': 'SMPL_MALE.pkl'}  # smpl_layer.py:30-35
GENDER_SEED = {'neutral': 0, 'female': 1, 'male': 2}


@dataclass
class SMPLModelData:
    """Host-side (numpy, float32 unless noted) constants of one body model."""
    v_template: np.ndarray   # (6890, 3)
    shapedirs: np.ndarray    # (6890, 3, 10)
    posedirs: np.ndarray     # (6890, 3, 207)
    J_regressor: np.ndarray  # (24, 6890) dense
    weights: np.ndarray      # (6890, 24)
    betas: np.ndarray        # (10,) model default betas (zeros in SMPL)
    faces: np.ndarray        # (F, 3) int64
    kintree_table: np.ndarray  # (2, 24) uint32
    gender: str = 'neutral'
    synthetic: bool = True

    @property
    def parents(self):
        return list(self.kintree_table[0].tolist())


def synthetic_smpl(gender: str = 'neutral', seed: int | None = None) -> SMPLModelData:
    """Deterministic SMPL-shaped random model.

    Skinning weights have exactly 4 random non-zeros per vertex (the worst case
    for joint locality), the joint regressor 40 non-zeros per joint.
    """
    if seed is None:
        seed = GENDER_SEED[gender]
    rng = np.random.default_rng(seed)
    v_template = rng.normal(0.0, 0.3, (NUM_VERTS, 3))
    shapedirs = rng.normal(0.0, 1e-2, (NUM_VERTS, 3, NUM_BETAS))
    posedirs = rng.normal(0.0, 1e-3, (NUM_VERTS, 3, NUM_POSE_FEATS))
    weights = np.zeros((NUM_VERTS, NUM_JOINTS))
    for v in range(NUM_VERTS):
        idx = rng.choice(NUM_JOINTS, 4, replace=False)
        w = rng.uniform(0.05, 1.0, 4)
        weights[v, idx] = w / w.sum()
    J_regressor = np.zeros((NUM_JOINTS, NUM_VERTS))
    for j in range(NUM_JOINTS):
        idx = rng.choice(NUM_VERTS, 40, replace=False)
        w = rng.uniform(0.05, 1.0, 40)
        J_regressor[j, idx] = w / w.sum()
    faces = rng.integers(0, NUM_VERTS, (13776, 3)).astype(np.int64)
    kintree = np.stack([np.array(SMPL_PARENTS, dtype=np.uint32),
                        np.arange(NUM_JOINTS, dtype=np.uint32)])
    f32 = np.float32
    return SMPLModelData(
        v_template=v_template.astype(f32), shapedirs=shapedirs.astype(f32),
        posedirs=posedirs.astype(f32), J_regressor=J_regressor.astype(f32),
        weights=weights.astype(f32), betas=np.zeros(NUM_BETAS, f32), faces=faces,
        kintree_table=kintree, gender=gender, synthetic=True)


class _ChStub:
    """Stand-in for chumpy classes while un-pickling a licensed SMPL file.

    chumpy's ``Ch`` pickles as an object whose state dict carries the array
    under ``x``; we only need that array (serialization.py:24-26 wraps the
    same fields with ``ch.array``)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})

    @property
    def r(self):
        return np.asarray(self.__dict__.get('x'))


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split('.')[0] == 'chumpy':
            return _ChStub
        return super().find_class(module, name)


def _arr(x):
    if isinstance(x, _ChStub):
        return x.r
    if hasattr(x, 'toarray'):
        return np.asarray(x.toarray())
    if hasattr(x, 'r'):
        return np.asarray(x.r)
    return np.asarray(x)


def load_smpl_pkl(path: str, gender: str = 'neutral') -> SMPLModelData:
    """Read a licensed SMPL pickle without chumpy (serialization.py:8-26)."""
    with open(path, 'rb') as f:
        dd = _Unpickler(f, encoding='latin1').load()
    f32 = np.float32
    shapedirs = _arr(dd['shapedirs']).astype(f32)
    betas = _arr(dd['betas']).astype(f32) if 'betas' in dd else np.zeros(shapedirs.shape[-1], f32)
    return SMPLModelData(
        v_template=_arr(dd['v_template']).astype(f32), shapedirs=shapedirs,
        posedirs=_arr(dd['posedirs']).astype(f32),
        J_regressor=_arr(dd['J_regressor']).astype(f32),
        weights=_arr(dd['weights']).astype(f32), betas=betas,
        faces=_arr(dd['f']).astype(np.int64),
        kintree_table=np.asarray(dd['kintree_table']).astype(np.uint32),
        gender=gender, synthetic=False)


def synthetic_allowed() -> bool:
    return os.environ.get('PRK_SYNTHETIC_SMPL', '') not in ('', '0')


def get_model_data(gender: str = 'neutral', model_root: str | None = None,
                   allow_synthetic: bool | None = None) -> SMPLModelData:
    """``model_root/SMPL_<GENDER>.pkl`` (read without chumpy).

    A missing file raises ``FileNotFoundError`` exactly like the reference's ``open()``
    (serialization.py:10).  The synthetic SMPL-shaped model is only handed out on request:
    ``allow_synthetic=True`` or the environment variable ``PRK_SYNTHETIC_SMPL=1`` (tests, bench) --
    and says so with a warning, because its vertices and joints mean nothing anatomically."""
    if gender not in GENDER_FILE:
        raise KeyError(gender)
    path = None
    if model_root is not None:
        path = os.path.join(model_root, GENDER_FILE[gender])
        if os.path.isfile(path):
            return load_smpl_pkl(path, gender)
    if allow_synthetic is None:
        allow_synthetic = synthetic_allowed()
    if not allow_synthetic:
        raise FileNotFoundError(
            f'SMPL model file {path or GENDER_FILE[gender]!r} not found (the licensed .pkl files are not shipped); '
            'pass model_data=synthetic_smpl(...) / allow_synthetic=True or set PRK_SYNTHETIC_SMPL=1 to run on the '
            'synthetic SMPL-shaped model instead')
    import warnings
    warnings.warn(f'poserisk_release_b200: using the SYNTHETIC SMPL-shaped {gender} model '
                  f'({path or "no model_root"} not found); vertices and joints are not anatomical', stacklevel=2)
    return synthetic_smpl(gender)
