"""Expert data-generation record: RecordPlanner.record_traj_plan / save_training_data
(record_planner.py:75-185) for a batch -- batch_plan on the device, then rows in the reference's 34-column
`train.csv` schema (record_planner.py:95-129) + normalised depth PNGs, readable by the reference's trainer
(nn_trainer.py:71-94: `row[1:-9]` motion inputs, `row[-9:]` targets, image `<id without 't'>.png`)."""
from __future__ import annotations

import datetime
import os

import numpy as np
import pandas as pd

from . import frames

TABLE_HEADER = (['id', 'drone_vel_x', 'drone_vel_y', 'drone_vel_z'] + [f'R{i}{j}' for i in (1, 2, 3) for j in (1, 2, 3)] +
                [f'{a}_{b}_{c}' for a in ('init', 'target') for b in ('pos', 'vel') for c in 'xyz'] +
                [f'wpts{i}_{c}' for i in (1, 2) for c in 'xyz'] + ['ts1', 'ts2', 'ts3'])
assert len(TABLE_HEADER) == 34


def training_rows(ids, motion_info, int_wpts_local, ts):
    """One DataFrame row per sample: id, 24 motion inputs, 6 body-frame waypoint coordinates, 3 durations."""
    data = np.concatenate([np.asarray(motion_info), np.asarray(int_wpts_local), np.asarray(ts)], axis=1)
    if data.shape[1] != 33:
        raise ValueError('the reference schema is fixed to M = 3 (2 waypoints, 3 durations)')
    df = pd.DataFrame(data, columns=TABLE_HEADER[1:])
    df.insert(0, 'id', list(ids))
    return df


def append_csv(csv_path, df):
    """record_planner.py:131-134 / :173: create with header once, then append without header."""
    os.makedirs(os.path.dirname(os.path.abspath(csv_path)), exist_ok=True)
    if not os.path.isfile(csv_path):
        pd.DataFrame(columns=TABLE_HEADER).to_csv(csv_path, index=False)
    df.to_csv(csv_path, mode='a', header=False, index=False)


def make_ids(n, now=None):
    """'t' + %Y%m%d%H%M%S + milliseconds (record_planner.py:156, :170); made unique within a batch."""
    now = now or datetime.datetime.now()
    base = int(now.strftime('%Y%m%d%H%M%S%f')[:-3])
    return ['t' + str(base + i) for i in range(n)]


class BatchRecorder:
    """record_traj_plan for B samples per call: `planner` is a planner.BatchPlanner with its maps set."""

    def __init__(self, planner, des_pos_z, out_dir):
        self.planner, self.des_pos_z = planner, des_pos_z
        self.csv_path = os.path.join(out_dir, 'train.csv')
        self.img_path = os.path.join(out_dir, 'depth_img')
        os.makedirs(self.img_path, exist_ok=True)

    def record(self, depth_img, local_vel, attitude, global_pos, global_vel, init_pos, init_vel, target_state,
               map_ids=None, rng=None, save_images=True):
        head = np.stack([np.asarray(init_pos)[:, :2], np.asarray(init_vel)[:, :2]], axis=1)      # record_planner.py:140-142
        res = self.planner.batch_plan(head, target_state, map_ids, rng=rng)                       # EP:142-168
        depth_norm, motion = frames.form_nn_input(depth_img, local_vel, attitude, global_pos, global_vel, self.des_pos_z,
                                                  init_pos, init_vel, target_state)
        keep = np.nonzero(res['ok'] == 1)[0]          # the reference raises (and records nothing) when planning fails
        nq = 4
        wl = frames.form_nn_output(np.asarray(attitude)[keep], np.asarray(global_pos)[keep], self.des_pos_z,
                                   res['x'][keep, :nq].reshape(-1, 2, 2))
        ids = make_ids(len(keep))
        df = training_rows(ids, motion[keep], wl, res['ts'][keep])
        append_csv(self.csv_path, df)
        if save_images:
            from PIL import Image
            for k, i in zip(keep, ids):
                Image.fromarray(depth_norm[k]).save(os.path.join(self.img_path, i[1:] + '.png'))
        return df, res
