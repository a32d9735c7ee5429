// astar_warp.cuh -- the reference's geometric initializer on a warp:
//   AP  = src/planner/scripts/traj_planner/astar_planner.py  (AstarPlanner.plan, AP:22-103)
//   GEO = src/planner/scripts/traj_planner/geo_planner.py    (seg_feasible_check GEO:41-55, prune_path_nodes GEO:57-101)
// One warp per start/target pair. The search is sequential by nature (one node is closed at a time and the choice
// depends on everything before it), so the warp parallelises what the reference does inside one step: the scan of the
// open set for the first minimum of g + h in insertion order (AP:62, a Python min() over a dict), the eight neighbour
// tests (AP:76-95) and the sampled line-of-sight checks of the pruning pass. Results are identical to the reference:
// same closed order, same parents, same path, same four key nodes.
//
// The open list {f, g, cell, insertion number} lives in shared memory (512 entries per warp, structure of arrays; longer
// lists spill to HBM), so the scan of every step reads on-chip data only and f is computed once per insertion/update.
// Per-warp scratch in HBM: ONE 4-byte node record per cell of the enlarged grid {move that led here, state/open position}
// (dense: the search indexes it by cell), plus an insertion list (used to wipe exactly the touched records afterwards
// and, later, as the key-node list) and the spill area of the open list, both sized for `icap` inserted nodes (28 B
// each), not for the grid: one search touches a few thousand records. With 400 x 400 search cells that is 0.64 + 0.46 MB
// per warp (round 1: 5.8 MB). A search that inserts more than icap nodes stops, reports ASTAR_OVERFLOW to the kernel and
// is re-run by the second pass with full-size lists on a few warps (k_astar, pass 1).
//
// The same source runs with a single lane on the host (devtools/astar_host.cu) -- a development aid used to check the
// logic against oracle/astar_ref.py in the GPU-less build container; the library never calls it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "minco_tile.cuh"

namespace neo {

#define NEO_HD_DECL __host__ __device__ __forceinline__

// Node record: bits 0..3 = the move that led to this node (0..7, index into the move table; 8 = start node, AP:16: the
// parent is this cell minus that move), bits 4..31 = state: 0 untouched, 1 closed (AP:72-73), k + 2: open, stored at
// position k of the open list.
typedef unsigned int AstarNode;
constexpr unsigned ASTAR_CLOSED = 1u;
NEO_HD_DECL unsigned an_state(AstarNode n) { return n >> 4; }
NEO_HD_DECL unsigned an_move(AstarNode n) { return n & 15u; }
NEO_HD_DECL AstarNode an_make(unsigned state, unsigned move) { return (state << 4) | move; }
// moves (1,0) (0,1) (-1,0) (0,-1) (-1,-1) (-1,1) (1,-1) (1,1): (d + 1) packed two bits per move
NEO_HD_DECL int an_dx(unsigned mv) { return (int)((0xA046u >> (2 * mv)) & 3u) - 1; }
NEO_HD_DECL int an_dy(unsigned mv) { return (int)((0x8819u >> (2 * mv)) & 3u) - 1; }
// grid index of the parent of cell `idx` (row length W), -1 for the start node
NEO_HD_DECL int an_parent(AstarNode n, int idx, int W) { const unsigned mv = an_move(n); return mv >= 8u ? -1 : idx - an_dx(mv) - an_dy(mv) * W; }

// One open node (AP:55 `open_set`). f = g + hypot is kept up to date with g, so the scan for the minimum reads f and, on
// ties, the insertion number only.
struct OpenRec {
    double f, g;
    int xy, tag;    // node x | y << 16; tag = n-th node ever inserted (the dict order the reference's min() resolves ties by)
};

// The open list: positions [0, cap) live in shared memory (structure of arrays, conflict-free scans), positions beyond
// that spill to this warp's block in HBM. Removal swaps the last entry into the hole.
struct OpenList {
    double *f, *g;
    int *xy, *tag;
    int cap;
    OpenRec *spill;
};

enum { ASTAR_FOUND = 0, ASTAR_EXHAUSTED = 1, ASTAR_START_OUTSIDE = 2, ASTAR_LIMIT = 3, ASTAR_OVERFLOW = 100 /* internal */ };

#define NEO_HD __host__ __device__ __forceinline__

// ---- lane primitives: 32 lanes on the device, 1 lane in the host build ---------------------------------------------
NEO_HD int a_lanes()
{
#ifdef __CUDA_ARCH__
    return 32;
#else
    return 1;
#endif
}
NEO_HD int a_lane()
{
#ifdef __CUDA_ARCH__
    return threadIdx.x & 31;
#else
    return 0;
#endif
}
NEO_HD void a_sync()
{
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}
NEO_HD unsigned a_ballot(bool p)
{
#ifdef __CUDA_ARCH__
    return __ballot_sync(0xffffffffu, p);
#else
    return p ? 1u : 0u;
#endif
}
NEO_HD int a_popc(unsigned v)
{
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
NEO_HD int a_from_lane0(int v)
{
#ifdef __CUDA_ARCH__
    return __shfl_sync(0xffffffffu, v, 0);
#else
    return v;
#endif
}
// IEEE operations without contraction (the reference is Python float arithmetic)
NEO_HD double a_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
NEO_HD double a_sub(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
NEO_HD double a_mul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
NEO_HD double a_div(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
NEO_HD double a_sqrt(double a)
{
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}
// lexicographic minimum of (f, tag) over the warp; f >= 0, so its bit pattern orders like an unsigned integer and three
// 32-bit warp reductions (high word, low word, tag) replace five rounds of 64-bit shuffles. Tags are unique: one winner.
NEO_HD void a_argmin(double &f, int &tag, int &k)
{
#ifdef __CUDA_ARCH__
    const unsigned FULL = 0xffffffffu;
    const unsigned hi = (unsigned)__double2hiint(f), lo = (unsigned)__double2loint(f);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    const bool cand = hi == mh && lo == ml;
    const unsigned mt = __reduce_min_sync(FULL, cand ? (unsigned)tag : 0xffffffffu);
    const int src = __ffs(__ballot_sync(FULL, cand && (unsigned)tag == mt)) - 1;
    f = __hiloint2double((int)mh, (int)ml);
    tag = (int)mt;
    k = __shfl_sync(FULL, k, src);
#endif
}

// ESDF.get_edt_dis (ESDF:53-67): nearest-cell read, 10000 outside the map, Python int() truncation
NEO_HD double a_dist(const MapView &map, double x, double y)
{
    const double tr = trunc(a_div(a_sub(y, map.oy), map.res));
    const double tc = trunc(a_div(a_sub(x, map.ox), map.res));
    if (!(tr >= 0.0 && tr < (double)map.H && tc >= 0.0 && tc < (double)map.W)) return 10000.0;
    return map.cells[(size_t)(int)tr * map.W + (int)tc].d;
}

// numpy.linspace(a, b, n)[i] (float64, endpoint=True): i * step + a with two roundings, the last sample is b itself
NEO_HD double a_linspace(double a, double b, int n, int i)
{
    if (n > 1 && i == n - 1) return b;
    if (n <= 1) return a;
    const double delta = a_sub(b, a), div = (double)(n - 1);
    const double step = a_div(delta, div);
    if (step == 0.0) return a_add(a_mul(a_div((double)i, div), delta), a);
    return a_add(a_mul((double)i, step), a);
}

// search-grid geometry of one map (AP:31-41)
struct AstarGrid {
    int W, H;
    double ox, oy;
};
NEO_HD AstarGrid astar_grid(const MapView &map)
{
    AstarGrid g;
    const int pad = (int)a_div(10.0, map.res);      // int(map_expand_radius / resolution)
    g.W = map.W + pad; g.H = map.H + pad;
    g.ox = a_sub(map.ox, 5.0); g.oy = a_sub(map.oy, 5.0);
    return g;
}
inline size_t astar_grid_cells(int H, int W, double res)
{
    const int pad = (int)(10.0 / res);
    return (size_t)(W + pad) * (size_t)(H + pad);
}

// AP:130-132 for one node of the enlarged grid: map.has_collision(calc_real_pos(ix, iy)). Evaluated once per map into
// MapView::blocked (k_astar_blocked), so the search reads one byte per neighbour instead of two divisions and a cell.
NEO_HD bool astar_blocked_at(const MapView &map, const AstarGrid &g, int ix, int iy)
{
    const double px = a_add(g.ox, a_mul((double)ix, map.res));       // AP:116-117
    const double py = a_add(g.oy, a_mul((double)iy, map.res));
    return a_dist(map, px, py) < 0.5;                                // ESDF:50-51
}

// GEO:41-55: every sample of the segment keeps 0.4 m clearance
NEO_HD bool a_segment_clear(const MapView &map, double x0, double y0, double x1, double y1)
{
    const int lane = a_lane(), NL = a_lanes();
    const double m = fmax(fabs(a_sub(x1, x0)), fabs(a_sub(y1, y0)));
    const int n = (int)ceil(a_div(m, 0.1)) + 1;
    bool bad = false;
    for (int i = lane; i < n; i += NL)
        if (a_dist(map, a_linspace(x0, x1, n, i), a_linspace(y0, y1, n, i)) < 0.4) bad = true;
    return a_ballot(bad) == 0u;
}

// open-list accessors (position k: shared memory below cap, HBM above)
NEO_HD OpenRec ol_get(const OpenList &o, int k)
{
    if (k >= o.cap) return o.spill[k - o.cap];
    OpenRec r; r.f = o.f[k]; r.g = o.g[k]; r.xy = o.xy[k]; r.tag = o.tag[k];
    return r;
}
NEO_HD void ol_put(const OpenList &o, int k, const OpenRec &r)
{
    if (k >= o.cap) { o.spill[k - o.cap] = r; return; }
    o.f[k] = r.f; o.g[k] = r.g; o.xy[k] = r.xy; o.tag[k] = r.tag;
}

// One start/target pair, executed by one warp (all lanes call with the same arguments).
//   nodes/order/ol.spill: this warp's scratch in HBM; nodes all-zero on entry and restored to all-zero on exit;
//   order and ol.spill hold icap entries (+8 ints of slack in order).
//   path_out (max_path, 2) may be NULL; path_len is the full length even when it exceeds max_path.
// Returns false (outputs untouched, scratch clean) when the search inserted more than icap nodes.
NEO_HD bool astar_problem(const MapView map, AstarNode *nodes, int *order, const OpenList ol, int icap, const double *start,
                          const double *target, int max_closed, int max_path, double *path_out, int32_t *path_len_out,
                          double *pruned_out, int32_t *status_out, int32_t *closed_out)
{
    const int lane = a_lane(), NL = a_lanes();
    const unsigned lt = (1u << lane) - 1u;
    const AstarGrid g = astar_grid(map);
    const double res = map.res;
    const double SQRT2 = 1.4142135623730951;                     // math.sqrt(2) (AP:111-114)

    // AP:119-120 (and Node.__init__'s int()): truncation toward zero
    const double fsx = trunc(a_div(a_sub(start[0], g.ox), res)), fsy = trunc(a_div(a_sub(start[1], g.oy), res));
    const double ftx = trunc(a_div(a_sub(target[0], g.ox), res)), fty = trunc(a_div(a_sub(target[1], g.oy), res));
    const double BIG = 1.0e9;
    int status = ASTAR_FOUND;
    if (!(fsx >= 0.0 && fsx < (double)g.W && fsy >= 0.0 && fsy < (double)g.H)) status = ASTAR_START_OUTSIDE;
    // a target far outside the grid can never be reached; clamp only so that index differences stay in range
    const int tx = (int)fmin(fmax(ftx, -BIG), BIG), ty = (int)fmin(fmax(fty, -BIG), BIG);
    int n_open = 0, n_seen = 0, n_closed = 0, t_parent = -1;

#define NEO_HYPOT(X, Y) a_sqrt((double)(((long long)(X) - tx) * ((long long)(X) - tx) + ((long long)(Y) - ty) * ((long long)(Y) - ty)))

    if (status == ASTAR_FOUND) {
        const int sx = (int)fsx, sy = (int)fsy;
        const int s_idx = sx + sy * g.W;                          // AP:122-124
        if (lane == 0) {
            OpenRec r; r.g = 0.0; r.f = a_add(0.0, NEO_HYPOT(sx, sy)); r.xy = sx | (sy << 16); r.tag = 1;
            ol_put(ol, 0, r);
            nodes[s_idx] = an_make(2u, 8u); order[0] = s_idx;
        }
        n_open = 1; n_seen = 1;
        a_sync();
        for (;;) {
            if (n_open == 0) { status = ASTAR_EXHAUSTED; break; }            // AP:58-60
            // AP:62: first minimum of cost + hypot in dict (insertion) order
            double bf = INFINITY; int bt = 0x7fffffff, bk = -1;
            const int n_fast = n_open < ol.cap ? n_open : ol.cap;
            for (int k = lane; k < n_fast; k += NL) {
                const double f = ol.f[k];
                const int t = ol.tag[k];
                if (f < bf || (f == bf && t < bt)) { bf = f; bt = t; bk = k; }
            }
            for (int k = ol.cap + lane; k < n_open; k += NL) {
                const OpenRec r = ol.spill[k - ol.cap];
                if (r.f < bf || (r.f == bf && r.tag < bt)) { bf = r.f; bt = r.tag; bk = k; }
            }
            a_argmin(bf, bt, bk);
            const OpenRec cur_rec = ol_get(ol, bk);
            const int cx = cur_rec.xy & 0xffff, cy = cur_rec.xy >> 16;
            const int cur = cx + cy * g.W;
            const double cg = cur_rec.g;
            if (cx == tx && cy == ty) { t_parent = an_parent(nodes[cur], cur, g.W); break; }  // AP:66-69
            a_sync();
            if (lane == 0) {                                                     // AP:72-73
                const int last = n_open - 1;
                if (bk != last) {
                    const OpenRec mv = ol_get(ol, last);
                    ol_put(ol, bk, mv);
                    AstarNode &moved = nodes[(mv.xy & 0xffff) + (mv.xy >> 16) * g.W];
                    moved = an_make((unsigned)bk + 2u, an_move(moved));
                }
                nodes[cur] = an_make(ASTAR_CLOSED, an_move(nodes[cur]));
            }
            n_open--; n_closed++;
            a_sync();
            if (max_closed > 0 && n_closed > max_closed) { status = ASTAR_LIMIT; break; }
            // AP:76-95: the eight moves, insertion order = move order
            for (int base = 0; base < 8; base += NL) {
                const int mv = base + lane;
                bool fresh = false; int nidx = 0, nxy = 0; double ng = 0.0, nf = 0.0;
                if (mv < 8) {
                    const int nx = cx + an_dx((unsigned)mv), ny = cy + an_dy((unsigned)mv);
                    if (nx >= 0 && nx < g.W && ny >= 0 && ny < g.H) {
                        nidx = nx + ny * g.W; nxy = nx | (ny << 16);
                        const AstarNode c = nodes[nidx];
                        const bool blk = map.blocked[nidx] != 0;                      // has_collision at this node
                        if (an_state(c) != ASTAR_CLOSED && !blk) {                    // AP:83, AP:86
                            ng = a_add(cg, mv < 4 ? 1.0 : SQRT2);
                            nf = a_add(ng, NEO_HYPOT(nx, ny));
                            if (an_state(c) == 0u) fresh = true;                      // AP:91-92
                            else {                                                    // AP:94-95
                                const int pos = (int)an_state(c) - 2;
                                OpenRec r = ol_get(ol, pos);
                                if (r.g > ng) { r.g = ng; r.f = nf; ol_put(ol, pos, r); nodes[nidx] = an_make(an_state(c), (unsigned)mv); }
                            }
                        }
                    }
                }
                const unsigned b = a_ballot(fresh);
                if (n_seen + a_popc(b) > icap) { status = ASTAR_OVERFLOW; break; }     // the lists are full: second pass
                if (fresh) {
                    const int r = a_popc(b & lt);
                    OpenRec nr; nr.f = nf; nr.g = ng; nr.xy = nxy; nr.tag = n_seen + r + 1;
                    ol_put(ol, n_open + r, nr);
                    nodes[nidx] = an_make((unsigned)(n_open + r) + 2u, (unsigned)mv); order[n_seen + r] = nidx;
                }
                n_open += a_popc(b); n_seen += a_popc(b);
            }
            a_sync();
            if (status == ASTAR_OVERFLOW) break;
        }
    }
#undef NEO_HYPOT

    // AP:143-151: [target cell] + closed parents, reversed. chain[] (in the open list's HBM part) holds the parents.
    int *chain = (int *)ol.spill;
    int n_chain = 0;
    if (status == ASTAR_FOUND && lane == 0) {
        int p = t_parent;
        while (p != -1) { chain[n_chain++] = p; p = an_parent(nodes[p], p, g.W); }
    }
    n_chain = a_from_lane0(n_chain);
    a_sync();
    for (int k = lane; k < n_seen; k += NL) nodes[order[k]] = 0u;
    if (status == ASTAR_OVERFLOW) { a_sync(); return false; }
    const int L = (status == ASTAR_FOUND || status == ASTAR_EXHAUSTED) ? n_chain + 1 : 0;

#define NEO_PATH_XY(i, X, Y)                                                                              \
    do {                                                                                                  \
        int ix_, iy_;                                                                                     \
        if ((i) < n_chain) { const int c_ = chain[n_chain - 1 - (i)]; ix_ = c_ % g.W; iy_ = c_ / g.W; }   \
        else { ix_ = tx; iy_ = ty; }                                                                      \
        X = a_add(g.ox, a_mul((double)ix_, res)); Y = a_add(g.oy, a_mul((double)iy_, res));               \
    } while (0)

    if (path_out)
        for (int i = lane; i < L && i < max_path; i += NL) {
            double x, y;
            NEO_PATH_XY(i, x, y);
            path_out[2 * i] = x; path_out[2 * i + 1] = y;
        }

    // GEO:61-75: greedy shortcutting; key indices go to keys[] (the insertion list's storage, free again)
    int *keys = order;
    int nk = 0;
    a_sync();
    if (L > 0) {
        if (lane == 0) keys[0] = 0;
        nk = 1;
        int head = 0, tail = 1;
        while (tail < L) {
            double hx, hy;
            NEO_PATH_XY(head, hx, hy);
            for (;;) {
                double qx, qy;
                NEO_PATH_XY(tail, qx, qy);
                if (!(a_segment_clear(map, hx, hy, qx, qy) || tail - head == 1)) break;
                tail++;
                if (tail == L) break;
            }
            if (lane == 0) keys[nk] = tail - 1;
            nk++;
            head = tail - 1;
        }
    }
    a_sync();

    // GEO:78-95: exactly four key nodes
    int pick[4] = {0, 0, 0, 0};
    if (nk == 2) {
        const double k0 = (double)keys[0], k1 = (double)keys[1];
        pick[0] = (int)a_linspace(k0, k1, 4, 0); pick[1] = (int)a_linspace(k0, k1, 4, 1);
        pick[2] = (int)a_linspace(k0, k1, 4, 2); pick[3] = (int)a_linspace(k0, k1, 4, 3);
    } else if (nk == 3) {
        const int k0 = keys[0], k1 = keys[1], k2 = keys[2];
        if (k1 - k0 > k2 - k1) { pick[0] = k0; pick[1] = (k0 + k1) / 2; pick[2] = k1; pick[3] = k2; }
        else { pick[0] = k0; pick[1] = k1; pick[2] = (k1 + k2) / 2; pick[3] = k2; }
    } else if (nk == 4) {
        for (int i = 0; i < 4; i++) pick[i] = keys[i];
    } else if (nk > 0) {
        const int last = keys[nk - 1];
        const double left = a_mul(a_div(1.0, 3.0), (double)last), right = a_mul(a_div(2.0, 3.0), (double)last);
        int il = keys[0], ir = keys[0];
        double dl = fabs(a_sub((double)il, left)), dr = fabs(a_sub((double)ir, right));
        for (int i = 1; i < nk; i++) {
            const int v = keys[i];
            const double el = fabs(a_sub((double)v, left)), er = fabs(a_sub((double)v, right));
            if (el < dl) { dl = el; il = v; }
            if (er < dr) { dr = er; ir = v; }
        }
        pick[0] = keys[0]; pick[1] = il; pick[2] = ir; pick[3] = last;
    }
    if (lane == 0) {
        for (int i = 0; i < 4; i++) {
            double x = 0.0, y = 0.0;
            if (L > 0) NEO_PATH_XY(pick[i], x, y);
            pruned_out[2 * i] = x; pruned_out[2 * i + 1] = y;
        }
        *path_len_out = L; *status_out = status; *closed_out = n_closed;
    }
#undef NEO_PATH_XY
    a_sync();
    return true;
}

#ifdef __CUDACC__
// persistent warps; warp w owns scratch block w
constexpr int ASTAR_OPEN_FAST = 512;     // open-list positions per warp held in shared memory (24 B each)
constexpr int ASTAR_WARPS_PER_CTA = 4;
constexpr size_t ASTAR_SMEM_BYTES = (size_t)ASTAR_WARPS_PER_CTA * ASTAR_OPEN_FAST * 24;
constexpr size_t ASTAR_BYTES_PER_INSERT = sizeof(int) + sizeof(OpenRec);      // insertion list + open-list spill
constexpr int ASTAR_ORDER_SLACK = 8;

struct AstarArgs {
    const MapView *maps;
    const int32_t *map_ids;      // (B) or NULL
    const double *start, *target;   // (B,2)
    int B, max_closed, max_path;
    double *path;                // (B,max_path,2) or NULL
    int32_t *path_len, *status, *closed;
    double *pruned;              // (B,4,2)
    AstarNode *nodes;            // per-warp blocks of `cap` node records
    int *order; OpenRec *spill;  // per-warp blocks of icap (+ slack) entries
    size_t cap;                  // cells of the largest enlarged grid
    int icap;                    // inserted nodes a search may reach in this pass
    unsigned int *counter;       // work queue
    int32_t *overflow;           // pass 0: problems whose search outgrew icap (count in overflow_count); pass 1: the work list
    unsigned int *overflow_count;
    int pass;
};

// one thread per node of the enlarged grid (map.blocked itself is not read)
__global__ void k_astar_blocked(const MapView map, unsigned char *__restrict__ out)
{
    const AstarGrid g = astar_grid(map);
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y;
    if (ix < g.W && iy < g.H) out[(size_t)iy * g.W + ix] = astar_blocked_at(map, g, ix, iy) ? 1 : 0;
}

__global__ void __launch_bounds__(ASTAR_WARPS_PER_CTA * 32) k_astar(const AstarArgs a)
{
    extern __shared__ __align__(16) unsigned char astar_smem[];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wc = threadIdx.x >> 5;
    OpenList ol;
    ol.cap = ASTAR_OPEN_FAST;
    ol.f = (double *)astar_smem + (size_t)wc * 2 * ASTAR_OPEN_FAST;
    ol.g = ol.f + ASTAR_OPEN_FAST;
    ol.xy = (int *)((double *)astar_smem + (size_t)ASTAR_WARPS_PER_CTA * 2 * ASTAR_OPEN_FAST) + (size_t)wc * 2 * ASTAR_OPEN_FAST;
    ol.tag = ol.xy + ASTAR_OPEN_FAST;
    ol.spill = a.spill + (size_t)warp * a.icap;
    AstarNode *nodes = a.nodes + (size_t)warp * a.cap;
    int *order = a.order + (size_t)warp * ((size_t)a.icap + ASTAR_ORDER_SLACK);
    // pass 0: every problem, lists of icap entries; pass 1: the problems pass 0 could not finish, full-size lists
    const unsigned total = a.pass ? *a.overflow_count : (unsigned)a.B;
    for (;;) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(a.counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= total) break;
        const unsigned b = a.pass ? (unsigned)a.overflow[q] : q;
        const MapView map = a.maps[a.map_ids ? a.map_ids[b] : 0];
        const bool done = astar_problem(map, nodes, order, ol, a.icap, a.start + 2 * (size_t)b, a.target + 2 * (size_t)b,
                                        a.max_closed, a.max_path, a.path ? a.path + (size_t)b * a.max_path * 2 : nullptr,
                                        a.path_len + b, a.pruned + 8 * (size_t)b, a.status + b, a.closed + b);
        if (!done && lane == 0) {
            if (a.pass == 0) a.overflow[atomicAdd(a.overflow_count, 1u)] = (int32_t)b;
            else { a.path_len[b] = 0; a.status[b] = ASTAR_LIMIT; a.closed[b] = 0; }      // cannot happen: icap = cap in pass 1
        }
    }
}
#endif

}  // namespace neo
