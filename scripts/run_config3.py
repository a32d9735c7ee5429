"""BASELINE.json configs[2]: learned-init path. 4,096 synthetic depth images + motion states -> initializer CNN (PyTorch,
bf16 autocast, cuDNN) -> batched optimizer; iterations-to-converge vs the expert straight-line init on the same problems.

The reference's trained weights are not in its repository, so the network has seeded random weights: its guesses are
not informative and the iteration comparison is reported as measured, not as a reproduction of the paper. To show the
mechanism the script also runs an "oracle initializer" (the optimum perturbed by 5 cm / 5 % noise), which is what a
trained network approximates.

    python scripts/run_config3.py [B]      -> one JSON line (also written to gpurun_out/config3.json)
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neo_planner_b200 import frames
from neo_planner_b200.initializer import NeoBatchPlanner
from neo_planner_b200.planner import BatchPlanner
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
cfg = YamlConfig()
w = make_world(0)
head, tail = make_problems(w, B)
yaw = np.arctan2(head[:, 1, 1], head[:, 1, 0])
att = np.stack([np.cos(yaw / 2), np.zeros(B), np.zeros(B), np.sin(yaw / 2)], axis=1)
gp = np.concatenate([head[:, 0] - 0.5 * head[:, 1], np.full((B, 1), 2.0)], axis=1)
gv = np.concatenate([head[:, 1], np.zeros((B, 1))], axis=1)
lv = frames.rotate_inverse(att, gv)
init_pos = np.concatenate([head[:, 0], np.full((B, 1), 2.0)], axis=1)
depth = rng.integers(1, 256, size=(B, 480, 640), dtype=np.uint8)          # synthetic depth frames (uint8, as stored)

bp = BatchPlanner(cfg); bp.set_map(w)
neo = NeoBatchPlanner(bp, des_pos_z=2.0, device='cuda', dtype=torch.bfloat16)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps

motion = frames.motion_info(lv, att, gp, gv, 2.0, init_pos, gv, tail)
dn, t_prep = timed(lambda: neo.normalize_depth(depth))             # H2D of the uint8 frames + per-image normalisation on the device
assert np.array_equal(dn[:16].cpu().numpy(), frames.normalize_depth(depth[:16]))
pred, t_net = timed(lambda: neo.predict(dn, motion))
nn_w, nn_ts = frames.wpts_world(att, gp, pred)
res_nn, t_opt_nn = timed(lambda: bp.warm_start_plan(head, tail, nn_w, nn_ts, rng=np.random.default_rng(1)))
res_ex, t_opt_ex = timed(lambda: bp.plan(head, tail, rng=np.random.default_rng(1)))
good = (res_ex['ok'] == 1)
noisy_w = res_ex['x'][:, :4].reshape(B, 2, 2) + rng.normal(0, 0.05, (B, 2, 2))
noisy_t = np.clip(res_ex['ts'] * (1 + rng.normal(0, 0.05, (B, 3))), 0.6, 4.9)
res_or, t_opt_or = timed(lambda: bp.warm_start_plan(head, tail, noisy_w, noisy_t, rng=np.random.default_rng(1)))
line = {'config': 'BASELINE.json configs[2]: learned-init path', 'samples': B, 'network': 'PlannerNetConv (ResNet-18 trunk), bf16 autocast, seeded random weights',
        'h2d_and_depth_normalisation_s': t_prep, 'cnn_s': t_net, 'cnn_images_per_s': B / t_net,
        'optimizer_s': {'nn_init': t_opt_nn, 'expert_init': t_opt_ex, 'near_optimum_init': t_opt_or},
        'pipeline_traj_per_s_nn_init': B / (t_prep + t_net + t_opt_nn),
        'mean_iterations': {'nn_init_random_weights': float(res_nn['nit'].mean()), 'expert_init': float(res_ex['nit'].mean()),
                            'near_optimum_init': float(res_or['nit'][good].mean())},
        'mean_evaluations': {'nn_init_random_weights': float(res_nn['nfev'].mean()), 'expert_init': float(res_ex['nfev'].mean()),
                             'near_optimum_init': float(res_or['nfev'][good].mean())},
        'ok_fraction': {'nn_init_random_weights': float(res_nn['ok'].mean()), 'expert_init': float(res_ex['ok'].mean()),
                        'near_optimum_init': float(res_or['ok'].mean())},
        'nn_ts_outside_bounds_fraction': float(((nn_ts <= cfg.T_min) | (nn_ts >= cfg.T_max)).any(axis=1).mean())}
print(json.dumps(line))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(line, open('gpurun_out/config3.json', 'w'), indent=1)
