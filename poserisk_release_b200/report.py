"""Host-side report writers of the reference's per-run output: the same text
files ``lib/core/base.py`` produces, written from the batched results of this package instead of
from per-frame Python lists.

  pose_to_str            lib/utils/vis_utils.py:9-16        "(x, y, z)" strings, 3 decimals
  post_processing        lib/core/base.py:242-271           sorted-score statistics (the score plot is
                                                            matplotlib work and not produced here)
  save_csv_pose_log      lib/core/base.py:329-350           debug/pose_log.csv
  save_csv               lib/core/base.py:352-397           debug/<TITLE>_score_log.csv, <TITLE>_eval_pose_log.csv
  write_result_txt       lib/core/base.py:158-165,176-183   reba_result.txt / rula_result.txt
  save_obj               lib/utils/vis_utils.py:238-245     Wavefront .obj of one frame's mesh (the debug
                                                            smpl_model.obj of base.py:273-282)

``timestamp`` is the reference's ``(0, frames, img_num)`` tuple (base.py:112): the CSVs get one row per
frame index in ``range(timestamp[0], timestamp[-1])`` and only the frames listed in ``timestamp[1]`` carry
values.  Pure formatting: no GPU work happens here.
"""
from __future__ import annotations

import csv
import os.path as osp

import numpy as np


def pose_to_str(poses):
    """(N, J, 3) Euler degrees -> N lists of J strings "(x, y, z)" with three decimals each."""
    poses = np.asarray(poses)
    return [['(%.3f, %.3f, %.3f)' % (j[0], j[1], j[2]) for j in frame] for frame in poses]


def post_processing(results):
    """results: the list of {'score', 'log_score'} dicts a scorer returns.  Returns the reference's
    ((mean, top-50 % mean, top-10 % mean, max, mode), scores in frame order, log_score array)."""
    from scipy.stats import mode
    scores = np.array([r['score'] for r in results])
    logs = np.array([r['log_score'] for r in results])
    in_frame_order = scores.copy()
    ranked = np.sort(scores)[::-1]
    n = len(ranked)
    final = (round(ranked.mean(), 3), round(ranked[:n // 2].mean(), 3), round(ranked[:n // 10].mean(), 3),
             round(ranked.max(), 3), mode(ranked).mode.item())
    return final, in_frame_order, logs


def _frame_rows(timestamp, cells_of_frame):
    """One CSV row per frame index of the clip; frames the tracker kept get `cells_of_frame(idx)` appended."""
    first, kept, last = timestamp[0], np.asarray(timestamp[1]), timestamp[-1]
    position = {int(f): i for i, f in enumerate(kept)}
    for frame in range(first, last):
        row = [frame]
        if frame in position:
            row.extend(cells_of_frame(position[frame]))
        yield row


def _write(path, header, rows):
    with open(path, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(header)
        for r in rows:
            w.writerow(r)


def save_csv_pose_log(pose_str, timestamp, output_path, debug_joints, joints_name_upper):
    """pose_log.csv: the Euler strings of the --debug_joints joints for every tracked frame."""
    cols = [joints_name_upper.index(name.upper()) for name in debug_joints]
    _write(osp.join(output_path, 'pose_log.csv'), ['Frame', 'Joint Pose'] + list(debug_joints),
           _frame_rows(timestamp, lambda i: [''] + [str(pose_str[i][c]) for c in cols]))


def save_csv(pose_str, timestamp, scores, joint_names, logs, pose_logs, output_path, title='REBA'):
    """<title>_score_log.csv (final score + per-part scores) and <title>_eval_pose_log.csv (the scorer's
    debug angle log, one column per key of pose_logs[0])."""
    n_parts = len(joint_names)
    _write(osp.join(output_path, title + '_score_log.csv'), ['Frame', 'Final_score', 'Joint Score'] + list(joint_names),
           _frame_rows(timestamp, lambda i: [str(scores[i]), ''] + [str(logs[i][j]) for j in range(n_parts)]))
    keys = list(pose_logs[0].keys())
    _write(osp.join(output_path, title + '_eval_pose_log.csv'), ['Frame', ''] + keys,
           _frame_rows(timestamp, lambda i: [''] + [str(pose_logs[i][k]) for k in keys]))


def result_text(final_score, action_level, action_name, title='REBA'):
    """The text of reba_result.txt / rula_result.txt, whitespace included (the reference builds it with a
    backslash-continued f-string, so the MAX line is preceded by a run of blanks; the REBA text ends
    with a blank, the RULA text does not)."""
    avg, top50, top10, mx, md = final_score
    pad = ' ' * 20
    tail = ' ' if title.upper() == 'REBA' else ''
    return (f'AVG Score: {avg} \n%50 Score: {top50} \n%10 Score: {top10} {pad}\nMAX Score: {mx} \n'
            f'MODE Score: {md} \nAction level: {action_level} \nAction: {action_name}{tail}')


def write_result_txt(final_score, action_level, action_name, output_path, title='REBA'):
    with open(osp.join(output_path, title.lower() + '_result.txt'), 'w') as f:
        f.write(result_text(final_score, action_level, action_name, title))


def save_obj(v, f=None, file_name=''):
    """Wavefront .obj of one mesh: a `v x y z` line per vertex with Python's shortest-repr number text (what
    ``str()`` of the element gives, so float32 input prints as float32), then -- when faces are given --
    one `f a/a b/b c/c` line per triangle with 1-based indices.  Accepts numpy arrays or torch tensors (a CUDA
    vertex tensor is copied to the host once)."""
    if hasattr(v, 'detach'):
        v = v.detach().cpu().numpy()
    if f is not None and hasattr(f, 'detach'):
        f = f.detach().cpu().numpy()
    lines = ['v %s %s %s\n' % (str(p[0]), str(p[1]), str(p[2])) for p in v]
    if f is not None:
        for tri in f:
            a, b, c = tri[0] + 1, tri[1] + 1, tri[2] + 1
            lines.append('f %s/%s %s/%s %s/%s\n' % (a, a, b, b, c, c))
    with open(file_name, 'w') as out:
        out.writelines(lines)


def save_debug_mesh(verts_m, faces, output_path):
    """The debug mesh of base.py:273-282: one frame's vertices (metres, (6890, 3)) scaled to millimetres as
    float32 and written to ``output_path/smpl_model.obj``."""
    if hasattr(verts_m, 'detach'):
        verts_m = verts_m.detach().cpu().numpy()
    v = np.asarray(verts_m).astype(np.float32).reshape(-1, 3) * 1000
    save_obj(v, faces, osp.join(output_path, 'smpl_model.obj'))
