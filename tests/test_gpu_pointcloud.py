"""§8f-1: voxel-centre list -> occupancy -> ESDF on the device, on the reference's own map fixture
(src/simulator/worlds/poles.pcd, regenerated from tests/golden/poles_columns.npz: full ground plane + 1,454 pillar
columns = the file's 190,732 voxel centres at 0.1 m)."""
import os

import numpy as np
import pytest

from neo_planner_b200 import lib, pointcloud
from neo_planner_b200.worlds import YamlConfig
from oracle import c_oracle, pointcloud_ref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def poles_points():
    g = np.load(os.path.join(GOLDEN, 'poles_columns.npz'))
    leaf = float(g['leaf'])
    gx = np.arange(g['ground_x'][0], g['ground_x'][1] + 1); gy = np.arange(g['ground_y'][0], g['ground_y'][1] + 1)
    X, Y = np.meshgrid(gx, gy, indexing='ij')
    vox = [np.stack([X.ravel(), Y.ravel(), np.zeros(X.size, np.int64)], 1)]
    for x, y, z0, z1 in g['columns']:
        z = np.arange(z0, z1 + 1)
        vox.append(np.stack([np.full(len(z), x), np.full(len(z), y), z], 1))
    vox = np.concatenate(vox)
    assert len(vox) == int(g['n_points']) == 190732
    return (vox * leaf + leaf / 2).astype(np.float32)


def test_poles_pointcloud_to_esdf():
    pts = poles_points()
    # map_server_global.launch:26-31: resolution 0.1, slab z in [0.5, 10]
    res, z_min, z_max = 0.1, 0.5, 10.0
    ox, oy, H, W = pointcloud.grid_for(pts, res)
    assert (H, W) == (300, 400) and abs(ox + 5.0) < 1e-12 and abs(oy + 15.0) < 1e-12
    h = lib.Handle(YamlConfig())
    h.set_map_points(0, pts, z_min, z_max, H, W, res, ox, oy)
    occ = h.get_occupancy(0, H, W)
    assert int((occ == 100).sum()) == 1454                      # SURVEY.md §8c probe of the same file
    assert np.array_equal(occ, pointcloud_ref.project_numpy(pts, z_min, z_max, H, W, res, ox, oy))
    e, gx, gy = h.get_map(0, H, W)
    m = c_oracle.OracleMap(occ, H, W, res, ox, oy)              # checker pinned bit-exact to scipy/numpy
    assert np.array_equal(e, m.esdf) and np.array_equal(gx, m.gx) and np.array_equal(gy, m.gy)
    # a slab that only sees the ground plane, and an empty slab (all-free map convention)
    h.set_map_points(0, pts, 0.0, 0.1, H, W, res, ox, oy)
    assert (h.get_occupancy(0, H, W) == 100).all()
    h.set_map_points(0, pts, 20.0, 30.0, H, W, res, ox, oy)
    assert (h.get_occupancy(0, H, W) == 0).all()


def test_pcd_reader_roundtrip(tmp_path):
    pts = poles_points()[::97]
    p = tmp_path / 'a.pcd'
    with open(p, 'w') as f:
        f.write('# .PCD v0.7\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n'
                f'WIDTH {len(pts)}\nHEIGHT 1\nVIEWPOINT 0 0 0 0 0 0 1\nPOINTS {len(pts)}\nDATA ascii\n')
        for x, y, z in pts:
            f.write(f'{x:.2f} {y:.2f} {z:.2f}\n')
    back = pointcloud.read_pcd_ascii(str(p))
    assert back.shape == pts.shape and np.allclose(back, pts, atol=1e-6)
