"""poserisk_release_b200 -- B200-native body-model -> risk-score path of PoseRisk.

Drop-in surface (same names and call contracts as the reference):
    SMPL_Layer                         lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py
    SMPL                               lib/utils/smpl.py
    REBA, RULA                         lib/utils/reba.py, lib/utils/rula.py
    get_joint_cam, rot_to_angle, axis_angle_to_euler_angle   lib/utils/coord_utils.py
plus the batched :class:`PoseRiskEngine`.  Everything computes in
csrc/libposerisk_b200.so (hand-written sm_100a CUDA behind a C ABI); importing the
compute classes without that library raises ImportError.
"""
from .model_provider import SMPLModelData, get_model_data, synthetic_smpl  # noqa: F401

__all__ = ['SMPL_Layer', 'SMPL', 'REBA', 'RULA', 'get_joint_cam', 'rot_to_angle', 'axis_angle_to_euler_angle',
           'PoseRiskEngine', 'synthetic_smpl', 'get_model_data', 'SMPLModelData']


def __getattr__(name):
    # lazy: keeps `import poserisk_release_b200.model_provider` usable without torch/CUDA
    if name == 'SMPL_Layer':
        from .smpl_layer import SMPL_Layer
        return SMPL_Layer
    if name == 'SMPL':
        from .smpl import SMPL
        return SMPL
    if name == 'REBA':
        from .reba import REBA
        return REBA
    if name == 'RULA':
        from .rula import RULA
        return RULA
    if name in ('get_joint_cam', 'rot_to_angle', 'axis_angle_to_euler_angle'):
        from . import coord_utils
        return getattr(coord_utils, name)
    if name == 'PoseRiskEngine':
        from .pipeline import PoseRiskEngine
        return PoseRiskEngine
    raise AttributeError(name)
