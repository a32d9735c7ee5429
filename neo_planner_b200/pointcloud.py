"""Point cloud -> planner map: the step the reference leaves to the external octomap_server
(`map_server_global.launch:17-31`: octomap of the cloud at `resolution`, projected to 2-D over the slab
[occupancy_min_z, occupancy_max_z]) before ESDF.occupancy_map_cb (ESDF:11-33) ever runs. Here the voxel-centre list
goes straight to the device: k_points_to_occ -> exact EDT -> gradient (csrc/map_kernels.cuh).

Projection semantics are restated from the launch parameters only (octomap_server is not in the reference repo):
parity of this step is pinned to a NumPy statement of the same rule (oracle/pointcloud_ref.py) and to the survey's cell count for
`src/simulator/worlds/poles.pcd`, not to octomap_server itself."""
from __future__ import annotations

import numpy as np


def read_pcd_ascii(path):
    """Minimal reader for the ascii .pcd files the reference ships/writes (FIELDS x y z, TYPE F;
    plugin_build_octomap.cpp:104-131). Returns (n, 3) float32."""
    with open(path) as f:
        n = None
        for line in f:
            if line.startswith('POINTS'):
                n = int(line.split()[1])
            if line.startswith('DATA'):
                if 'ascii' not in line:
                    raise ValueError('only ascii .pcd is supported')
                break
        pts = np.loadtxt(f, dtype=np.float32, ndmin=2)
    if n is not None and len(pts) != n:
        raise ValueError(f'{path}: header says {n} points, found {len(pts)}')
    return pts[:, :3]


def grid_for(points, res, margin=0.0):
    """Axis-aligned grid covering the cloud: origin on the voxel edge below the minimum centre (ox, oy, H, W)."""
    pts = np.asarray(points, dtype=np.float64)
    lo = np.floor(pts[:, :2].min(axis=0) / res) * res - margin
    hi = pts[:, :2].max(axis=0) + margin
    W = int(np.floor((hi[0] - lo[0]) / res)) + 1
    H = int(np.floor((hi[1] - lo[1]) / res)) + 1
    return float(lo[0]), float(lo[1]), H, W
