// map_kernels.cuh -- ESDF.occupancy_map_cb (ESDF:11-33) and ESDF.get_edt_dis/get_edt_grad (ESDF:53-82) on the
// device. Integer-exact Euclidean distance transform (separable: row scan, then pruned exact column minimum),
// IEEE sqrt * res, numpy.gradient central differences; results are bit-identical to
// scipy.ndimage.distance_transform_edt(1-occ)*res and numpy.gradient (verified in tests/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "minco_tile.cuh"

namespace neo {

constexpr int EDT_INF = 1 << 28;

// Point cloud / voxel-centre list -> 2-D occupancy: a cell is occupied (100) iff at least one point with
// z_min <= z <= z_max falls into its column (what octomap_server's projected_map delivers to ESDF.occupancy_map_cb for
// the slab [occupancy_min_z, occupancy_max_z], map_server_global.launch:26-31). xyz: (n,3) float32 as stored in .pcd.
__global__ void k_points_to_occ(const float *__restrict__ xyz, int n, double z_min, double z_max, double ox, double oy,
                                double res, int H, int W, int8_t *__restrict__ occ)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
    if (!(z >= z_min && z <= z_max)) return;
    const double fc = floor(__ddiv_rn(__dsub_rn(x, ox), res)), fr = floor(__ddiv_rn(__dsub_rn(y, oy), res));
    if (fr >= 0.0 && fr < (double)H && fc >= 0.0 && fc < (double)W) occ[(size_t)(int)fr * W + (int)fc] = 100;
}

// pass 1: one warp per row; g[r][c] = squared distance to the nearest occupied cell of row r (EDT_INF if none).
// The row is walked in 32-cell chunks, forwards (nearest occupied cell at or left of c) and backwards (at or right of
// c); inside a chunk every lane finds its neighbour in the ballot word of the chunk, the carry is one integer.
__global__ void k_edt_rows(const int8_t *__restrict__ occ, int H, int W, int *__restrict__ g, int *__restrict__ any)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= H) return;
    const int8_t *o = occ + (size_t)r * W;
    int *row = g + (size_t)r * W;
    int last = -1;
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int c = c0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, c < W && o[c] == 100);   // ESDF:23: only 100 is occupied, unknown(-1) is free
        const unsigned mine = m & (0xffffffffu >> (31 - lane));                // occupied cells at or left of this lane
        const int left = mine ? c0 + 31 - __clz(mine) : last;
        if (c < W) row[c] = left < 0 ? EDT_INF : (c - left) * (c - left);
        if (m) last = c0 + 31 - __clz(m);
    }
    if (lane == 0 && last >= 0) atomicOr(any, 1);
    last = -1;
    for (int c0 = ((W - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        const int c = c0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, c < W && o[c] == 100);
        const unsigned mine = m & (0xffffffffu << lane);                       // occupied cells at or right of this lane
        const int right = mine ? c0 + __ffs(mine) - 1 : last;
        if (c < W && right >= 0) { const int d = (right - c) * (right - c); if (d < row[c]) row[c] = d; }
        if (m) last = c0 + __ffs(m) - 1;
    }
}

// pass 2: one thread per cell; exact min over rows rr of (r-rr)^2 + g[rr][c], scanning outward from r and
// stopping once the vertical offset alone exceeds the best value (exact pruning).
__global__ void k_edt_cols(const int *__restrict__ g, int H, int W, const int *__restrict__ any, double res,
                           double *__restrict__ esdf)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= W) return;
    long long best;
    if (!*any) {
        // scipy's answer for a map without features: one virtual feature at (row -1, col 0)
        best = (long long)(r + 1) * (r + 1) + (long long)c * c;
    } else {
        int b = g[(size_t)r * W + c];
        for (int k = 1; k < H; k++) {
            const int k2 = k * k;
            if (k2 >= b) break;
            if (r - k >= 0) { const int v = g[(size_t)(r - k) * W + c] + k2; if (v < b) b = v; }
            if (r + k < H) { const int v = g[(size_t)(r + k) * W + c] + k2; if (v < b) b = v; }
        }
        best = b;
    }
    esdf[(size_t)r * W + c] = __dmul_rn(__dsqrt_rn((double)best), res);       // ESDF:29
}

// np.gradient (ESDF:33): interior (f[i+1]-f[i-1])/2, borders one-sided; packs {gx, gy, d, 0} per cell
__global__ void k_pack_cells(const double *__restrict__ esdf, const double *__restrict__ gx_in,
                             const double *__restrict__ gy_in, int H, int W, Cell *__restrict__ cells,
                             double *__restrict__ gx_out, double *__restrict__ gy_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= W) return;
    const size_t i = (size_t)r * W + c;
    double gx, gy;
    if (gx_in) { gx = gx_in[i]; gy = gy_in[i]; }
    else {
        if (c == 0) gx = __dsub_rn(esdf[i + 1], esdf[i]);
        else if (c == W - 1) gx = __dsub_rn(esdf[i], esdf[i - 1]);
        else gx = __ddiv_rn(__dsub_rn(esdf[i + 1], esdf[i - 1]), 2.0);
        if (r == 0) gy = __dsub_rn(esdf[i + W], esdf[i]);
        else if (r == H - 1) gy = __dsub_rn(esdf[i], esdf[i - W]);
        else gy = __ddiv_rn(__dsub_rn(esdf[i + W], esdf[i - W]), 2.0);
    }
    Cell cell;
    cell.gx = gx; cell.gy = gy; cell.d = esdf[i]; cell.pad = 0.0;
    cells[i] = cell;
    if (gx_out) { gx_out[i] = gx; gy_out[i] = gy; }
}

__global__ void k_unpack_cells(const Cell *__restrict__ cells, size_t n, double *__restrict__ esdf,
                               double *__restrict__ gx, double *__restrict__ gy)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cell c = cells[i];
    esdf[i] = c.d; gx[i] = c.gx; gy[i] = c.gy;
}

// ESDF:53-82 for a batch of points
__global__ void k_query(MapView map, int n, const double *__restrict__ xy, int32_t *__restrict__ idx,
                        double *__restrict__ dis, double *__restrict__ grad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = xy[2 * i], y = xy[2 * i + 1];
    const double fr = __ddiv_rn(__dsub_rn(y, map.oy), map.res);     // ESDF:61
    const double fc = __ddiv_rn(__dsub_rn(x, map.ox), map.res);     // ESDF:62
    const double tr = trunc(fr), tc = trunc(fc);                    // Python int(): toward zero
    const bool inside = tr >= 0.0 && tr < (double)map.H && tc >= 0.0 && tc < (double)map.W;
    if (inside) {
        const Cell c = map.cells[(size_t)(int)tr * map.W + (int)tc];
        idx[2 * i] = (int)tr; idx[2 * i + 1] = (int)tc;
        dis[i] = c.d; grad[2 * i] = c.gx; grad[2 * i + 1] = c.gy;
    } else {
        idx[2 * i] = -1; idx[2 * i + 1] = -1;
        dis[i] = 10000.0; grad[2 * i] = 0.0; grad[2 * i + 1] = 0.0;
    }
}

}  // namespace neo
