// astar_host.cu -- development aid: runs csrc/astar_warp.cuh with ONE lane on the host so that its logic can be checked
// against oracle/astar_ref.py in the GPU-less build container (devtools/README.md). Not part of the library and never
// loaded by it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -o devtools/_astar_host.so devtools/astar_host.cu
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../neo_planner_b200/csrc/astar_warp.cuh"

// open_fast: positions of the open list kept in the "shared memory" arrays (small values exercise the spill path)
extern "C" int sim_astar(int H, int W, double res, double ox, double oy, const double *esdf, int B, const double *start,
                         const double *target, int max_closed, int max_path, int open_fast, double *path,
                         int32_t *path_len, double *pruned, int32_t *status, int32_t *closed)
{
    std::vector<neo::Cell> cells((size_t)H * W);
    for (size_t i = 0; i < cells.size(); i++) { cells[i].gx = 0; cells[i].gy = 0; cells[i].d = esdf[i]; cells[i].pad = 0; }
    neo::MapView map;
    map.cells = cells.data(); map.H = H; map.W = W; map.res = res; map.ox = ox; map.oy = oy; map.inv_res = 1.0 / res;
    const size_t cap = neo::astar_grid_cells(H, W, res);
    map.blocked = nullptr;
    const neo::AstarGrid grid = neo::astar_grid(map);
    std::vector<unsigned char> blocked(cap);
    for (int iy = 0; iy < grid.H; iy++)
        for (int ix = 0; ix < grid.W; ix++) blocked[(size_t)iy * grid.W + ix] = neo::astar_blocked_at(map, grid, ix, iy) ? 1 : 0;
    map.blocked = blocked.data();
    std::vector<neo::AstarNode> nodes(cap);
    memset(nodes.data(), 0, sizeof(neo::AstarNode) * cap);
    std::vector<int> order(cap + neo::ASTAR_ORDER_SLACK);
    std::vector<neo::OpenRec> spill(cap);
    // like k_astar: a first pass with room for `icap` inserted nodes (env NEO_ASTAR_ICAP; default: the whole grid), and a
    // second one with full-size lists for the searches that outgrow it
    size_t icap = cap;
    if (const char *e = getenv("NEO_ASTAR_ICAP")) { const long v = atol(e); if (v > 0 && (size_t)v < cap) icap = (size_t)v; }
    std::vector<double> f(open_fast), g(open_fast);
    std::vector<int> xy(open_fast), tag(open_fast);
    neo::OpenList ol;
    ol.f = f.data(); ol.g = g.data(); ol.xy = xy.data(); ol.tag = tag.data(); ol.cap = open_fast; ol.spill = spill.data();
    for (int b = 0; b < B; b++) {
        for (int pass = 0; pass < 2; pass++) {
            const bool done = neo::astar_problem(map, nodes.data(), order.data(), ol, (int)(pass ? cap : icap), start + 2 * b,
                                                 target + 2 * b, max_closed, max_path,
                                                 path ? path + (size_t)b * max_path * 2 : nullptr, path_len + b,
                                                 pruned + 8 * b, status + b, closed + b);
            if (done) break;
            if (pass == 1) return 2;
        }
    }
    // the scratch must come back clean
    for (size_t i = 0; i < cap; i++)
        if (nodes[i] != 0u) return 1;
    return 0;
}
