"""Learned initializer of NEO-Planner: depth image + 24 motion floats -> 2 body-frame
waypoints + 3 durations, which warm-start the MINCO optimizer (neo_planner.py:42-51).

The network is the one dense contraction on the path, so it stays PyTorch (cuDNN/cuBLAS convs, bf16 autocast on the
GPU) as north_star prescribes; no custom kernel. Architecture restated from nn_trainer.py:109-155 (`PlannerNet`,
MLP heads) and nn_trainer_conv.py:108-160 (`PlannerNetConv`, Conv1d heads -- the variant that is saved as
saved_net/planner_net.*, which nn_planner.py:25 loads). The reference's trained weights are not in the repository
(.MISSING_LARGE_BLOBS) and torchvision's ImageNet weights cannot be downloaded here, so weights are seeded random:
throughput and plumbing are real, the paper's iteration savings are not reproducible from this repo alone."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import frames

IMG_WIDTH, IMG_HEIGHT, MOTION_INPUT_SIZE, OUTPUT_SIZE = 640, 480, 24, 9
IMG_FEATURE_SIZE = MOTION_FEATURE_SIZE = 24


def _backbone():
    from torchvision import models
    net = models.resnet18(weights=None)
    net.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)      # 1-channel depth input
    net.fc = nn.Linear(net.fc.in_features, IMG_FEATURE_SIZE)
    return net


class PlannerNet(nn.Module):
    """nn_trainer.py:109-155."""

    def __init__(self):
        super().__init__()
        self.img_backbone = _backbone()
        self.motion_backbone = nn.Sequential(nn.Linear(MOTION_INPUT_SIZE, 48), nn.LeakyReLU(), nn.Linear(48, 24),
                                             nn.LeakyReLU(), nn.Linear(24, 24), nn.LeakyReLU(),
                                             nn.Linear(24, MOTION_FEATURE_SIZE))
        self.mlp = nn.Sequential(nn.Linear(IMG_FEATURE_SIZE + MOTION_FEATURE_SIZE, 48), nn.LeakyReLU(), nn.Linear(48, 96),
                                 nn.LeakyReLU(), nn.Linear(96, 96), nn.LeakyReLU(), nn.Linear(96, OUTPUT_SIZE))

    def forward(self, x):
        img = x[:, :IMG_WIDTH * IMG_HEIGHT].reshape(-1, 1, IMG_HEIGHT, IMG_WIDTH)
        vec = x[:, IMG_WIDTH * IMG_HEIGHT:]
        return self.mlp(torch.cat([self.img_backbone(img), self.motion_backbone(vec)], dim=1))


class PlannerNetConv(nn.Module):
    """nn_trainer_conv.py:108-160."""

    def __init__(self):
        super().__init__()
        self.img_backbone = _backbone()

        def head(width, out):
            return nn.Sequential(nn.Conv1d(1, 16, 3, 1, 1), nn.LeakyReLU(), nn.Conv1d(16, 32, 3, 1, 1), nn.LeakyReLU(),
                                 nn.Conv1d(32, 64, 3, 1, 1), nn.LeakyReLU(), nn.Flatten(), nn.Linear(64 * width, out))
        self.motion_backbone = head(MOTION_INPUT_SIZE, MOTION_FEATURE_SIZE)
        self.mlp = head(IMG_FEATURE_SIZE + MOTION_FEATURE_SIZE, OUTPUT_SIZE)

    def forward(self, x):
        img = x[:, :IMG_WIDTH * IMG_HEIGHT].reshape(-1, 1, IMG_HEIGHT, IMG_WIDTH)
        vec = x[:, IMG_WIDTH * IMG_HEIGHT:].unsqueeze(1)
        feat = torch.cat([self.img_backbone(img), self.motion_backbone(vec)], dim=1)
        return self.mlp(feat.unsqueeze(1))


def process_input(depth_norm, motion_info):
    """nn_trainer.py:51-58 for a batch: (B,H,W) uint8 + (B,24) -> (B, 307224) float32."""
    B = depth_norm.shape[0]
    return np.concatenate([depth_norm.reshape(B, -1).astype(np.float32), np.asarray(motion_info).astype(np.float32)], axis=1)


class NeoBatchPlanner:
    """NeoPlanner.enhanced_traj_plan (neo_planner.py:42-51) for B samples: network guess -> warm_start_plan.

    `planner` is a planner.BatchPlanner. Network guesses with durations outside (T_min, T_max) lose their first
    attempt exactly as in the reference (map_T2tau raises, EP:209 -> re-seeded straight line, EP:197-200)."""

    def __init__(self, planner, des_pos_z=2.0, net=None, device='cuda', dtype=torch.bfloat16, seed=42):
        self.planner, self.des_pos_z, self.device, self.dtype = planner, des_pos_z, torch.device(device), dtype
        if net is None:
            torch.manual_seed(seed)
            net = PlannerNetConv()
        self.net = net.to(self.device).eval()
        if self.device.type == 'cuda':
            self.net = self.net.to(memory_format=torch.channels_last)
            torch.backends.cudnn.benchmark = True        # fixed 480x640 frames: let cuDNN pick its fastest kernels once

    @torch.no_grad()
    def normalize_depth(self, depth_img, chunk=512):
        """record_planner.py:15 on the device: (depth / max(depth) * 255).astype(uint8), evaluated in fp64 like NumPy
        does, so the truncation to uint8 is the reference's. depth_img: (B,H,W) array (any real dtype). Returns a uint8
        CUDA tensor (B,H,W)."""
        depth_img = np.asarray(depth_img)
        out = torch.empty(depth_img.shape, dtype=torch.uint8, device=self.device)
        for i in range(0, len(depth_img), chunk):
            d = torch.from_numpy(np.ascontiguousarray(depth_img[i:i + chunk])).to(self.device, non_blocking=True).double()
            mx = d.amax(dim=(1, 2), keepdim=True)
            out[i:i + chunk] = (d / mx * 255).to(torch.uint8)
        return out

    @torch.no_grad()
    def predict(self, depth_norm, motion_info, chunk=512):
        """Network forward for B samples. depth_norm: (B,H,W) uint8, NumPy or CUDA tensor (images travel as uint8 and
        are widened on the device; the flattened float vector of nn_trainer.py:51-58 is formed there). Returns (B,9)."""
        if not torch.is_tensor(depth_norm):
            depth_norm = torch.from_numpy(np.ascontiguousarray(depth_norm))
        motion = torch.from_numpy(np.asarray(motion_info, dtype=np.float32)).to(self.device)
        B = depth_norm.shape[0]
        outs = []
        for i in range(0, B, chunk):
            img = depth_norm[i:i + chunk].to(self.device, non_blocking=True).reshape(-1, IMG_WIDTH * IMG_HEIGHT)
            with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.device.type == 'cuda'):
                lowp = self.dtype if self.device.type == 'cuda' else torch.float32      # 0..255 is exact in bf16
                x = torch.cat([img.to(lowp), motion[i:i + chunk].to(lowp)], dim=1)
                outs.append(self.net(x).float())
        return torch.cat(outs).double().cpu().numpy()

    def enhanced_traj_plan(self, depth_img, local_vel, attitude, global_pos, global_vel, init_pos, init_vel, target_state,
                           map_ids=None, rng=None):
        motion = frames.motion_info(local_vel, attitude, global_pos, global_vel, self.des_pos_z, init_pos, init_vel,
                                    target_state)
        out = self.predict(self.normalize_depth(depth_img), motion)
        int_wpts, ts = frames.wpts_world(attitude, global_pos, out, M=3)
        head = np.stack([np.asarray(init_pos)[:, :2], np.asarray(init_vel)[:, :2]], axis=1)
        res = self.planner.warm_start_plan(head, target_state, int_wpts, ts, map_ids, rng)
        res['nn_int_wpts'], res['nn_ts'] = int_wpts, ts
        return res
