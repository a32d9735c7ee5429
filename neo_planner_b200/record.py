"""Expert data-generation record: RecordPlanner.record_traj_plan / save_training_data
(record_planner.py:75-185) for a batch -- batch_plan on the device, then rows in the reference's 34-column
`train.csv` schema (record_planner.py:95-129) + normalised depth PNGs, readable by the reference's trainer
(nn_trainer.py:71-94: `row[1:-9]` motion inputs, `row[-9:]` targets, image `<id without 't'>.png`)."""
from __future__ import annotations

import datetime
import os
import threading

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


_last_id = 0
_id_lock = threading.Lock()


def make_ids(n, now=None, rank=0, world_size=1):
    """'t' + %Y%m%d%H%M%S + milliseconds (record_planner.py:156, :170). The reference writes one sample per wall-clock
    millisecond, so its ids are unique; a batch takes n consecutive numbers, and the next call starts after the last
    number any earlier call of this process handed out (never before `now`), so two batches written within the same
    few milliseconds cannot collide and overwrite each other's depth PNGs. Sharded writers (one process per GPU,
    rank / world_size) take interleaved numbers: id = base + i * world_size + rank."""
    global _last_id
    now = now or datetime.datetime.now()
    stamp = int(now.strftime('%Y%m%d%H%M%S%f')[:-3])
    with _id_lock:
        base = max(stamp, _last_id + 1)
        base += (-base) % world_size                     # aligned, so that ranks starting in the same millisecond interleave
        _last_id = base + n * world_size
    return ['t' + str(base + i * world_size + rank) for i in range(n)]


class BatchRecorder:
    """record_traj_plan for B samples per call: `planner` is a planner.BatchPlanner with its maps set."""

    def __init__(self, planner, des_pos_z, out_dir, rank=0, world_size=1):
        self.planner, self.des_pos_z = planner, des_pos_z
        self.rank, self.world_size = rank, world_size
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
        ids = make_ids(len(keep), rank=self.rank, world_size=self.world_size)
        df = training_rows(ids, motion[keep], wl, res['ts'][keep])
        append_csv(self.csv_path, df)
        if save_images:
            from PIL import Image
            for k, i in zip(keep, ids):
                path = os.path.join(self.img_path, i[1:] + '.png')
                if os.path.exists(path):
                    raise FileExistsError(f'{path}: refusing to overwrite the depth image of an earlier sample')
                Image.fromarray(depth_norm[k]).save(path)
        return df, res


class RecordPlanner:
    """Drop-in for the reference's RecordPlanner (record_planner.py:75-185; `selected_planner:=record`): one sample
    per call -- batch_plan on the device (EP:142-168), then one row of train.csv + one normalised depth PNG.
    The reference derives its output directory from its own file location; here it is an argument."""

    def __init__(self, planner_config, out_dir=None, device: int = 0):
        from .planner import MinJerkPlanner
        self._planner = MinJerkPlanner(planner_config, device)
        self.des_pos_z = planner_config.des_pos_z
        out_dir = out_dir or os.path.join(os.getcwd(), 'training_data')
        self.csv_path = os.path.join(out_dir, 'train.csv')
        self.img_path = os.path.join(out_dir, 'depth_img')
        os.makedirs(self.img_path, exist_ok=True)
        self.table_header = list(TABLE_HEADER)
        if not os.path.isfile(self.csv_path):
            pd.DataFrame(columns=TABLE_HEADER).to_csv(self.csv_path, index=False)

    def __getattr__(self, name):          # int_wpts, ts, coeffs, get_pos, ...: everything else is the planner's
        return getattr(self._planner, name)

    def record_traj_plan(self, map, depth_img, drone_state, plan_init_state, target_state):
        drone_state_2d = np.array([plan_init_state.global_pos[:2], plan_init_state.global_vel[:2]])
        self._planner.batch_plan(map, drone_state_2d, target_state)
        self.save_training_data(depth_img, drone_state, plan_init_state, target_state, self._planner.int_wpts, self._planner.ts)

    def save_training_data(self, depth_img, drone_state, plan_init_state, target_state, int_wpts, ts):
        q = frames.quat_array(drone_state.attitude)[None]
        depth_norm, motion = frames.form_nn_input(np.asarray(depth_img)[None], np.asarray(drone_state.local_vel)[None], q,
                                                  np.asarray(drone_state.global_pos)[None], np.asarray(drone_state.global_vel)[None],
                                                  self.des_pos_z, np.asarray(plan_init_state.global_pos)[None],
                                                  np.asarray(plan_init_state.global_vel)[None],
                                                  np.asarray(target_state, dtype=np.float64)[None, :2, :2])
        local = frames.form_nn_output(q, np.asarray(drone_state.global_pos)[None], self.des_pos_z, np.asarray(int_wpts)[None])
        ids = make_ids(1)
        append_csv(self.csv_path, training_rows(ids, motion, local, np.asarray(ts)[None]))
        from PIL import Image
        Image.fromarray(depth_norm[0]).save(os.path.join(self.img_path, ids[0][1:] + '.png'))
        print("Training data (ID: %s) saved!" % ids[0][1:])
