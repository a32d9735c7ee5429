import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200.geo import BatchGeoPlanner
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig
w = make_world(0)
bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
bp.set_map(w)
head, tail = make_problems(w, 1024, M=3)
for _ in range(3):
    r = bp.handle.astar(head[:, 0], tail[:, 0])
print(bp.handle.last_kernel_ms())
