"""net_weights.py -- TEST INFRASTRUCTURE ONLY. Deterministic stand-in weights for the initializer networks.

The reference's trained weights are not in its repository (.MISSING_LARGE_BLOBS) and torchvision's ImageNet weights
cannot be downloaded here, so structural parity of neo_planner_b200/initializer.py with the reference's PlannerNet
classes (nn_trainer.py:109-155, nn_trainer_conv.py:108-160) is pinned like this: both sides fill their state_dict from
the SAME function of (parameter name, shape) below; oracle/gen_golden.py runs the reference classes and stores their
outputs (tests/golden/nets.npz); tests/test_host_logic.py runs ours and compares."""
import zlib

import numpy as np
import torch


def fill_deterministic(net, seed=2024):
    sd = net.state_dict()
    for name in sorted(sd):
        p = sd[name]
        if not torch.is_floating_point(p):
            continue
        g = torch.Generator().manual_seed(seed + zlib.crc32(name.encode()))
        v = torch.randn(p.shape, generator=g, dtype=torch.float32) * 0.05
        if name.endswith('running_var'):
            v = v.abs() + 1.0
        elif name.endswith('bn1.weight') or name.endswith('bn2.weight') or name.endswith('downsample.1.weight'):
            v = v + 1.0
        p.copy_(v)
    return net


def sample_input(n=2, seed=7):
    """(n, 307224) float32: a seeded uint8 depth frame (480 x 640) + 24 motion floats, as process_input_np lays it out."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (n, 480 * 640)).astype(np.float32)
    motion = rng.normal(0, 1.5, (n, 24)).astype(np.float32)
    return np.concatenate([img, motion], axis=1)


def signature(net):
    return [f'{k}:{tuple(v.shape)}' for k, v in net.state_dict().items()]
