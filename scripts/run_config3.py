"""BASELINE.json configs[2]: learned-init path. 4,096 synthetic depth images + motion states -> initializer CNN (PyTorch,
bf16 autocast, cuDNN) -> batched optimizer; iterations-to-converge vs the expert straight-line init on the same problems.

The reference's trained weights are not in its repository. Three initializers are therefore measured against the expert
straight-line guess on the same 4,096 problems:
  * the network with seeded random weights (plumbing / throughput; predicted durations are clamped into (T_min, T_max)
    as SURVEY.md §8d prescribes, so the predicted waypoints do warm-start the optimizer instead of dying in map_T2tau);
  * the same network after a short training run HERE, on expert data this library generates (the reference's own
    pipeline: record -> train -> neo; nn_trainer_conv.py:162-230: MSE loss, Adam 1e-3) -- the synthetic depth frames carry
    no information, so it learns from the 24 motion inputs only;
  * a near-optimum guess (the optimum perturbed by 5 cm / 5 % noise): what a well-trained network approximates.

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
neo = NeoBatchPlanner(bp, des_pos_z=2.0, device='cuda', dtype=torch.bfloat16, clamp_ts=True)

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
nn_ts_c, touched = frames.clamp_durations(nn_ts, cfg.T_min, cfg.T_max)
res_nn, t_opt_nn = timed(lambda: bp.warm_start_plan(head, tail, nn_w, nn_ts_c, rng=np.random.default_rng(1)))
res_ex, t_opt_ex = timed(lambda: bp.plan(head, tail, rng=np.random.default_rng(1)))

# ---- a short training run on expert data generated here (other problems of the same world) ---------------------------
TRAIN_N, STEPS, BS = 8192, 300, 64
th, tt = make_problems(w, B + TRAIN_N)
th, tt = th[B:], tt[B:]
tyaw = np.arctan2(th[:, 1, 1], th[:, 1, 0])
tatt = np.stack([np.cos(tyaw / 2), np.zeros(TRAIN_N), np.zeros(TRAIN_N), np.sin(tyaw / 2)], axis=1)
tgp = np.concatenate([th[:, 0] - 0.5 * th[:, 1], np.full((TRAIN_N, 1), 2.0)], axis=1)
tgv = np.concatenate([th[:, 1], np.zeros((TRAIN_N, 1))], axis=1)
tip = np.concatenate([th[:, 0], np.full((TRAIN_N, 1), 2.0)], axis=1)
texp = bp.plan(th, tt, rng=np.random.default_rng(2))
keep = np.nonzero(texp['ok'] == 1)[0]
t_motion = frames.motion_info(frames.rotate_inverse(tatt, tgv), tatt, tgp, tgv, 2.0, tip, tgv, tt)[keep]
t_target = np.concatenate([frames.form_nn_output(tatt[keep], tgp[keep], 2.0, texp['x'][keep, :4].reshape(-1, 2, 2)), texp['ts'][keep]], axis=1)
net = neo.net.float().train()
for p_ in net.img_backbone.parameters():
    p_.requires_grad = False                                        # nn_trainer_conv.py:115-117: the image trunk stays frozen
net.img_backbone.conv1.weight.requires_grad = True; net.img_backbone.fc.weight.requires_grad = True; net.img_backbone.fc.bias.requires_grad = True
opt = torch.optim.Adam([p_ for p_ in net.parameters() if p_.requires_grad], lr=1e-3)
xm = torch.from_numpy(t_motion.astype(np.float32)).cuda(); yt = torch.from_numpy(t_target.astype(np.float32)).cuda()
gen = torch.Generator(device='cuda').manual_seed(0)
t0 = time.perf_counter()
for step in range(STEPS):
    idx = torch.randint(0, len(keep), (BS,), device='cuda', generator=gen)
    img = torch.randint(1, 256, (BS, 480 * 640), device='cuda', generator=gen).float()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out = net(torch.cat([img, xm[idx]], dim=1)).float().reshape(BS, 9)
    loss = torch.nn.functional.mse_loss(out, yt[idx])
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
torch.cuda.synchronize()
t_train = time.perf_counter() - t0
neo.net = net.eval().to(memory_format=torch.channels_last)
pred_t = neo.predict(dn, motion)
tr_w, tr_ts = frames.wpts_world(att, gp, pred_t)
tr_ts_c, tr_touched = frames.clamp_durations(tr_ts, cfg.T_min, cfg.T_max)
res_tr, t_opt_tr = timed(lambda: bp.warm_start_plan(head, tail, tr_w, tr_ts_c, rng=np.random.default_rng(1)))
good = (res_ex['ok'] == 1)
noisy_w = res_ex['x'][:, :4].reshape(B, 2, 2) + rng.normal(0, 0.05, (B, 2, 2))
noisy_t = np.clip(res_ex['ts'] * (1 + rng.normal(0, 0.05, (B, 3))), 0.6, 4.9)
res_or, t_opt_or = timed(lambda: bp.warm_start_plan(head, tail, noisy_w, noisy_t, rng=np.random.default_rng(1)))
line = {'config': 'BASELINE.json configs[2]: learned-init path', 'samples': B, 'network': 'PlannerNetConv (ResNet-18 trunk), bf16 autocast, seeded random weights',
        'h2d_and_depth_normalisation_s': t_prep, 'cnn_s': t_net, 'cnn_images_per_s': B / t_net,
        'optimizer_s': {'nn_init_random_weights': t_opt_nn, 'nn_init_trained_here': t_opt_tr, 'expert_init': t_opt_ex, 'near_optimum_init': t_opt_or},
        'training_here': {'samples': int(len(keep)), 'steps': STEPS, 'batch': BS, 'seconds': t_train, 'final_mse': float(loss.item())},
        'pipeline_traj_per_s_nn_init': B / (t_prep + t_net + t_opt_nn),
        'mean_iterations': {'nn_init_random_weights': float(res_nn['nit'].mean()), 'nn_init_trained_here': float(res_tr['nit'].mean()), 'expert_init': float(res_ex['nit'].mean()),
                            'near_optimum_init': float(res_or['nit'][good].mean())},
        'mean_evaluations': {'nn_init_random_weights': float(res_nn['nfev'].mean()), 'nn_init_trained_here': float(res_tr['nfev'].mean()), 'expert_init': float(res_ex['nfev'].mean()),
                             'near_optimum_init': float(res_or['nfev'][good].mean())},
        'ok_fraction': {'nn_init_random_weights': float(res_nn['ok'].mean()), 'nn_init_trained_here': float(res_tr['ok'].mean()), 'expert_init': float(res_ex['ok'].mean()),
                        'near_optimum_init': float(res_or['ok'].mean())},
        'first_attempt_accepted_fraction': {'nn_init_random_weights': float((res_nn['attempt'] == 0).mean()), 'nn_init_trained_here': float((res_tr['attempt'] == 0).mean()),
                                            'expert_init': float((res_ex['attempt'] == 0).mean())},
        'nn_ts_clamped_fraction': {'random_weights': float(touched.mean()), 'trained_here': float(tr_touched.mean())}}
print(json.dumps(line))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(line, open('gpurun_out/r2_config3.json', 'w'), indent=1)
