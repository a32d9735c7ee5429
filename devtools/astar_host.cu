// astar_host.cu -- development aid: runs csrc/astar_warp.cuh with ONE lane on the host so that its logic can be checked
// against oracle/astar_ref.py in the GPU-less build container (devtools/README.md). Not part of the library and never
// loaded by it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -o devtools/_astar_host.so devtools/astar_host.cu
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../neo_planner_b200/csrc/astar_warp.cuh"

extern "C" int sim_astar(int H, int W, double res, double ox, double oy, const double *esdf, int B, const double *start,
                         const double *target, int max_closed, int max_path, double *path, int32_t *path_len,
                         double *pruned, int32_t *status, int32_t *closed)
{
    std::vector<neo::Cell> cells((size_t)H * W);
    for (size_t i = 0; i < cells.size(); i++) { cells[i].gx = 0; cells[i].gy = 0; cells[i].d = esdf[i]; cells[i].pad = 0; }
    neo::MapView map;
    map.cells = cells.data(); map.H = H; map.W = W; map.res = res; map.ox = ox; map.oy = oy; map.inv_res = 1.0 / res;
    const size_t cap = neo::astar_grid_cells(H, W, res);
    std::vector<neo::AstarNode> nodes(cap);
    memset(nodes.data(), 0, sizeof(neo::AstarNode) * cap);
    std::vector<int> open(cap), order(cap);
    for (int b = 0; b < B; b++) {
        neo::astar_problem(map, nodes.data(), open.data(), order.data(), start + 2 * b, target + 2 * b, max_closed, max_path,
                           path ? path + (size_t)b * max_path * 2 : nullptr, path_len + b, pruned + 8 * b, status + b,
                           closed + b);
    }
    // the scratch must come back clean
    for (size_t i = 0; i < cap; i++)
        if (nodes[i].tag != 0 || nodes[i].g != 0.0 || nodes[i].parent != 0) return 1;
    return 0;
}
