"""Drop-in for the reference's ``RULA`` scorer (lib/utils/rula.py:7-422).

Same call contract as :class:`poserisk_release_b200.reba.REBA`; the arithmetic runs
in csrc/prk_score.cu.
"""
from __future__ import annotations

import numpy as np

from . import _lib, _scorer
from ._scorer import JOINT_NAME, score_euler_records


class RULA:
    def __init__(self, debug=False):
        self.joint_name = JOINT_NAME
        # rula.py:13-58, kept for API parity
        self.table_a = np.array([
            [[[1, 2], [2, 2], [2, 3], [3, 3]], [[2, 2], [2, 2], [3, 3], [3, 3]], [[2, 3], [3, 3], [3, 3], [4, 4]]],
            [[[2, 3], [3, 3], [3, 4], [4, 4]], [[3, 3], [3, 3], [3, 4], [4, 4]], [[3, 4], [4, 4], [4, 4], [5, 5]]],
            [[[3, 3], [4, 4], [4, 4], [5, 5]], [[3, 4], [4, 4], [4, 4], [5, 5]], [[4, 4], [4, 4], [4, 5], [5, 5]]],
            [[[4, 4], [4, 4], [4, 5], [5, 5]], [[4, 4], [4, 4], [4, 5], [5, 5]], [[4, 4], [4, 5], [5, 5], [6, 6]]],
            [[[5, 5], [5, 5], [5, 6], [6, 7]], [[5, 6], [6, 6], [6, 7], [7, 7]], [[6, 6], [6, 7], [7, 7], [7, 8]]],
            [[[7, 7], [7, 7], [7, 8], [8, 9]], [[8, 8], [8, 8], [8, 9], [9, 9]], [[9, 9], [9, 9], [9, 9], [9, 9]]]])
        self.table_b = np.array([
            [[1, 3], [2, 3], [3, 4], [5, 5], [6, 6], [7, 7]], [[2, 3], [2, 3], [4, 5], [5, 5], [6, 7], [7, 7]],
            [[3, 3], [3, 4], [4, 5], [5, 5], [6, 7], [7, 7]], [[5, 5], [5, 6], [6, 7], [7, 7], [7, 7], [8, 8]],
            [[7, 7], [7, 7], [7, 8], [8, 8], [8, 8], [8, 8]], [[8, 8], [8, 8], [8, 8], [8, 9], [9, 9], [9, 9]]])
        self.table_c = np.array([
            [1, 2, 3, 3, 4, 5, 5], [2, 2, 3, 4, 4, 5, 5], [3, 3, 3, 4, 4, 5, 6], [3, 3, 3, 4, 5, 6, 6],
            [4, 4, 4, 5, 6, 7, 7], [5, 5, 6, 6, 7, 7, 7], [5, 5, 6, 7, 7, 7, 7]])
        self.eval_items = ['Upper_arm (L,R)', 'Lower_arm (L,R)', 'Wrist (L,R)', 'Wrist_twist (L,R)', 'Neck', 'Trunk', 'Leg']
        self.debugging = debug
        self.angle_log = {}
        self.log = []

    def score_records(self, poses, add_info, track_of_frame=None):
        """(extension) structured array of prk_score_rec for all frames, no per-frame dicts."""
        return score_euler_records(poses, add_info, _lib.PRK_SCORE_RULA, track_of_frame)

    def __call__(self, poses, joint_cams, add_info):
        n = len(poses)
        if n > 0:
            joint_cams[n - 1]          # the reference indexes joint_cams[ii] (rula.py:71)
        rec = self.score_records(poses, add_info)
        score = rec['rula_score'].astype(np.int64)
        p = rec['rula_parts']
        results = []
        for ii in range(n):
            q = p[ii]
            results.append({
                'score': score[ii],
                'log_score': [f'{q[0]},{q[1]}', f'{q[2]},{q[3]}', f'{q[4]},{q[5]}', f'{q[6]},{q[7]}',
                              int(q[8]), int(q[9]), int(q[10])],
            })
            if self.debugging:
                self.log.append(self._angle_log(poses[ii]))
                self.angle_log = {}
        return results

    def _angle_log(self, pose):
        """Debug strings in the reference's insertion order (rula.py:120-156 call order)."""
        j = self.joint_name.index
        P = lambda name, c: pose[j(name)][c]
        log = {}
        a3, a4 = P('R_Shoulder', 2), P('R_Shoulder', 1)
        if a3 > -70 and a3 < 110 and abs(a4) < 20:
            a4 = 1                       # rula.py:183 overwrites angle4 before it is printed (:198)
        log['upper_arm_bending'] = (f"L {P('L_Shoulder', 2):.1f},{P('L_Shoulder', 1):.1f} "
                                    f"R {a3:.1f},{a4:.1f}")
        log['shoulder_rise'] = f"L {P('L_Thorax', 2):.1f} R {P('R_Thorax', 2):.1f}"
        # rula.py:284 prints angle1/angle2 = L_Shoulder[2], L_Shoulder[1]
        log['upper_arm_abducted'] = f"L {P('L_Shoulder', 2):.1f} R {P('L_Shoulder', 1):.1f}"
        log['lower_arm_bending'] = (f"L {max(P('L_Elbow', 1), P('L_Elbow', 2)):.1f} "
                                    f"R {max(P('R_Elbow', 1), P('R_Elbow', 2)):.1f}")
        log['bent_from_midline_or_out_to_side'] = f"L {P('L_Thorax', 0):.1f} R {P('R_Thorax', 0):.1f}"
        log['wrist_bending'] = f"L {P('L_Wrist', 2):.1f} R {P('R_Wrist', 2):.1f}"
        log['wrist_side_bending'] = f"L {P('L_Wrist', 1):.1f} R {P('R_Wrist', 1):.1f}"
        log['wrist_twist'] = f"L {P('L_Wrist', 0):.1f} R {P('R_Wrist', 0):.1f}"
        log['neck_bending'] = f"{P('Neck', 0):.1f}"
        log['neck_side_bending_twisted'] = f"{P('Neck', 2):.1f}, {P('Neck', 1):.1f}"
        log['trunk_bending'] = f"{P('Torso', 0):.1f}"
        log['trunk_twisted'] = f"{P('Torso', 1):.1f}"
        log['trunk_side_bending'] = f"{P('Torso', 2):.1f}"
        return log

    # (highest score of the band, level, text): rula.py:100-118; the last band is open-ended
    ACTION_BANDS = ((2, 1, "Acceptable posture"),
                    (4, 2, "Further investigation, change may be needed"),
                    (6, 3, "Further investigation, change soon"),
                    (None, 4, "Investigate and implement change"))

    def action_level(self, score):
        return _scorer.action_band(self.ACTION_BANDS, score)
