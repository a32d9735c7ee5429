"""pointcloud_ref.py -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).

NumPy statement of the point cloud -> 2-D occupancy projection that k_points_to_occ (csrc/map_kernels.cuh) performs on the
device. The reference leaves this step to the external octomap_server
(map_server_global.launch:17-31), which is not in the reference repo: the rule is restated from the launch parameters --
a cell is occupied iff a point with z_min <= z <= z_max falls into its column -- so this pin is "parity unpinned" with
respect to octomap_server itself."""
import numpy as np


def project_numpy(points, z_min, z_max, H, W, res, ox, oy):
    """NumPy statement of the projection rule (test oracle for k_points_to_occ)."""
    p = np.asarray(points, dtype=np.float32).astype(np.float64)
    m = (p[:, 2] >= z_min) & (p[:, 2] <= z_max)
    c = np.floor((p[m, 0] - ox) / res); r = np.floor((p[m, 1] - oy) / res)
    ok = (r >= 0) & (r < H) & (c >= 0) & (c < W)
    occ = np.zeros((H, W), np.int8)
    occ[r[ok].astype(int), c[ok].astype(int)] = 100
    return occ
