"""Drop-in for the reference's ``REBA`` scorer (lib/utils/reba.py:7-392).

``REBA(debug)(poses, joint_cams, add_info)`` returns the same list of
``{'score', 'log_score'}`` dicts; the rule ladders and table look-ups run in the
integer CUDA kernel of libposerisk_b200.so (csrc/prk_score.cu), quirks included.
"""
from __future__ import annotations

import numpy as np

from . import _lib, _scorer
from ._scorer import JOINT_NAME, score_euler_records


class REBA:
    def __init__(self, debug=False):
        self.joint_name = JOINT_NAME
        # tables kept as attributes for API parity (reba.py:13-43); the kernel holds its own copy
        self.table_a = np.array(
            [[[1, 2, 3, 4], [1, 2, 3, 4], [3, 3, 5, 6]],
             [[2, 3, 4, 5], [3, 4, 5, 6], [4, 5, 6, 7]],
             [[2, 4, 5, 6], [4, 5, 6, 7], [5, 6, 7, 8]],
             [[3, 5, 6, 7], [5, 6, 7, 8], [6, 7, 8, 9]],
             [[4, 6, 7, 8], [6, 7, 8, 9], [7, 8, 9, 9]]])
        self.table_b = np.array(
            [[[1, 2, 2], [1, 2, 3]], [[1, 2, 3], [2, 3, 4]], [[3, 4, 5], [4, 5, 5]],
             [[4, 5, 5], [5, 6, 7]], [[6, 7, 8], [7, 8, 8]], [[7, 8, 8], [8, 9, 9]]])
        self.table_c = np.array([
            [1, 1, 1, 2, 3, 3, 4, 5, 6, 7, 7, 7], [1, 2, 2, 3, 4, 4, 5, 6, 6, 7, 7, 8],
            [2, 3, 3, 3, 4, 5, 6, 7, 7, 8, 8, 8], [3, 4, 4, 4, 5, 6, 7, 8, 8, 9, 9, 9],
            [4, 4, 4, 5, 6, 7, 8, 8, 9, 9, 9, 9], [6, 6, 6, 7, 8, 8, 9, 9, 10, 10, 10, 10],
            [7, 7, 7, 8, 9, 9, 9, 10, 10, 11, 11, 11], [8, 8, 8, 9, 10, 10, 10, 10, 10, 11, 11, 11],
            [9, 9, 9, 10, 10, 10, 11, 11, 11, 12, 12, 12], [10, 10, 10, 11, 11, 11, 11, 12, 12, 12, 12, 12],
            [11, 11, 11, 11, 12, 12, 12, 12, 12, 12, 12, 12], [12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12]])
        self.eval_items = ['Trunk', 'Neck', 'Leg', 'Upper_arm (L,R)', 'Lower_arm (L,R)', 'Wrist (L,R)']
        self.debugging = debug
        self.angle_log = {}
        self.log = []

    def score_records(self, poses, add_info, track_of_frame=None):
        """(extension) structured array of prk_score_rec for all frames, no per-frame dicts."""
        return score_euler_records(poses, add_info, _lib.PRK_SCORE_REBA, track_of_frame)

    def __call__(self, poses, joint_cams, add_info):
        n = len(poses)
        if n > 0:
            joint_cams[n - 1]          # the reference indexes joint_cams[ii] (reba.py:55)
        rec = self.score_records(poses, add_info)
        score = rec['reba_score'].astype(np.int64)
        p = rec['reba_parts']
        results = []
        for ii in range(n):
            q = p[ii]
            results.append({
                'score': score[ii],
                'log_score': [int(q[0]), int(q[1]), int(q[2]), f'{q[3]},{q[4]}', f'{q[5]},{q[6]}', f'{q[7]},{q[8]}'],
            })
            if self.debugging:
                self.log.append(self._angle_log(poses[ii]))
                self.angle_log = {}
        return results

    def _angle_log(self, pose):
        """Debug strings in the reference's insertion order (reba.py:142-390)."""
        j = self.joint_name.index
        P = lambda name, c: pose[j(name)][c]
        log = {}
        log['trunk_bending'] = f"{P('Torso', 0):.1f}"
        log['trunk_twist'] = f"{P('Torso', 1):.1f}"
        log['trunk_side_bending'] = f"{P('Torso', 2):.1f}"
        log['neck_bending'] = f"{P('Neck', 0):.1f}"
        log['neck_twist'] = f"{P('Neck', 2):.1f},{P('Neck', 1):.1f}"
        log['leg_bending'] = f"L {P('L_Knee', 0):.1f} R {P('R_Knee', 0):.1f}"
        log['upper_arm_bending'] = (f"L {P('L_Shoulder', 2):.1f},{P('L_Shoulder', 1):.1f} "
                                    f"R {P('R_Shoulder', 2):.1f},{P('R_Shoulder', 1):.1f}")
        log['shoulder_rise'] = f"L {P('L_Thorax', 2):.1f} R {P('R_Thorax', 2):.1f}"
        # reba.py:334 prints angle3/angle4 = L_Shoulder[1], R_Shoulder[2] on the R side
        log['upper_arm_abducted_rotated'] = (f"L {P('L_Shoulder', 2):.1f},{P('L_Shoulder', 0):.1f} "
                                             f"R {P('L_Shoulder', 1):.1f},{P('R_Shoulder', 2):.1f}")
        log['lower_arm_bending'] = (f"L {max(P('L_Elbow', 1), P('L_Elbow', 2)):.1f} "
                                    f"R {max(P('R_Elbow', 1), P('R_Elbow', 2)):.1f}")
        log['wrist_bending'] = f"L {P('L_Wrist', 2):.1f} R {P('R_Wrist', 2):.1f}"
        log['wrist_side_bending_or_twisted'] = (f"L {P('L_Wrist', 1):.1f},{P('L_Wrist', 0):.1f} "
                                                f"R {P('R_Wrist', 1):.1f},{P('R_Wrist', 0):.1f}")
        return log

    # (highest score of the band, level, text): reba.py:83-104; the last band is open-ended
    ACTION_BANDS = ((1, 1, "Negligible risk"),
                    (3, 2, "Low risk. Change may be needed."),
                    (7, 3, "Medium risk. Further Investigate. Change Soon."),
                    (10, 4, "High risk. Investigate and implement change"),
                    (None, 5, "Very high risk. Implement change"))

    def action_level(self, score):
        return _scorer.action_band(self.ACTION_BANDS, score)
