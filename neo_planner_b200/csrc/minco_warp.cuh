// minco_warp.cuh -- one warp = one planning problem: fused get_cost + get_grad of the reference
// optimizer (EP:539-585) in fp64 on sm_100a.
//
//   tau -> T (EP:477-483, double-double exp)             lanes 0..M-1, one piece each
//   MINCO system A c = b (EP:261-336)                    banded LU, kl = ku = 6, no pivoting, in shared memory
//   energy / time cost and gradient (EP:345-390)         lanes 0..M-1
//   sampled feasibility + collision penalties (EP:392-466)  lanes over samples, ESDF cells gathered from L2,
//                                                        16-value butterfly ("transposed") warp reduction
//   adjoint A^T G = dW/dc, grad_q, grad_T, grad_tau (EP:494-537, EP:485-492)
//
// Row order: the reference orders the six rows of interior waypoint i as
//   [pos=q, pos-cont, vel-cont, acc-cont, jerk-cont, snap-cont]   (rows 6i+3 .. 6i+8, EP:284-316)
// which puts zeros on the diagonal (numpy pivots). We permute them to
//   [jerk-cont, snap-cont, pos=q, pos-cont, vel-cont, acc-cont]   (rows 6i+3 .. 6i+8 of P A)
// so that P A is banded with nonzero pivots and factors without pivoting; G is un-permuted on the fly
// (reference row 6i+3+a  <->  permuted row 6i+3+PERM[a], PERM = {2,3,4,5,0,1}).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dd_exp.h"

namespace neo {

constexpr unsigned FULL = 0xffffffffu;
constexpr int BW = 13;          // band storage width: offsets -6..+6
constexpr int HIST = 10;        // L-BFGS memory (maxcor, EP:220)

struct DevParams {
    double v_max2, T_min, T_max, safe_dis, dt;
    double w0, w1, w2, w3;
    double collision_cost_tol;
};

// One ESDF cell = one 32-byte L2 sector: {grad_x, grad_y, dist, pad}
struct __align__(32) Cell {
    double gx, gy, d, pad;
};

struct MapView {
    const Cell *cells;
    int H, W;
    double res, ox, oy;
};

// Per-warp shared-memory slice (all doubles). n = 3M-2 decision variables, N = 6M rows.
struct WarpMem {
    double *Ab;    // [N][BW] band of P A, overwritten by its LU factors
    double *rinv;  // [N] reciprocals of the U diagonal
    double *c;     // [N][2] rhs -> polynomial coefficients
    double *gC;    // [N][2] dW/dc -> adjoint variable (permuted rows)
    double *ts;    // [M]
    double *ex;    // [M] exp(-tau)
    double *gT;    // [M]
    double *ht;    // [12] head (3,2), tail (3,2)
    double *S;     // [HIST][n]
    double *Y;     // [HIST][n]
    double *rho;   // [HIST]
};

__host__ __device__ inline int warp_mem_doubles(int M)
{
    int N = 6 * M, n = 3 * M - 2;
    int tot = N * BW + N + 2 * N + 2 * N + 3 * M + 12 + 2 * HIST * n + HIST;
    return (tot + 1) & ~1;
}

__device__ inline WarpMem carve(double *base, int M)
{
    int N = 6 * M, n = 3 * M - 2;
    WarpMem m;
    m.Ab = base; base += N * BW;
    m.rinv = base; base += N;
    m.c = base; base += 2 * N;
    m.gC = base; base += 2 * N;
    m.ts = base; base += M;
    m.ex = base; base += M;
    m.gT = base; base += M;
    m.ht = base; base += 12;
    m.S = base; base += HIST * n;
    m.Y = base; base += HIST * n;
    m.rho = base;
    return m;
}

#define AB(r, j) Ab[(r) * BW + ((j) - (r) + 6)]

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// Reduce 16 per-lane values over the warp with 16 (not 80) double shuffles: after the call lane l holds
// in v[0] the warp-wide total of slot (l >> 1).
__device__ __forceinline__ void warp_reduce16(double (&v)[16], int lane)
{
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double send = up ? v[i] : v[i + 8];
            double keep = up ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double send = up ? v[i] : v[i + 4];
            double keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            double send = up ? v[i] : v[i + 2];
            double keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        double send = up ? v[0] : v[1];
        double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
}

// ---------------------------------------------------------------------------------------------------------
// MINCO system: build P A and b, factor, solve. (EP:261-336)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_system(const WarpMem &m, int M, int lane, double xl)
{
    const int N = 6 * M, nq = 2 * (M - 1);
    double *Ab = m.Ab;
    for (int i = lane; i < N * BW; i += 32) Ab[i] = 0.0;
    for (int i = lane; i < 2 * N; i += 32) m.c[i] = 0.0;
    __syncwarp();
    if (lane < nq) {   // b rows "p_i(T_i) = q_i": x = [q_x(0..M-2), q_y(0..M-2), ...] (EP:211)
        int d = lane / (M - 1), i = lane - d * (M - 1);
        m.c[(6 * i + 5) * 2 + d] = xl;
    }
    if (lane < 6) {    // head rows 0..2, tail rows N-3..N-1 (EP:274-275)
        m.c[lane] = m.ht[lane];
        m.c[(N - 3) * 2 + lane] = m.ht[6 + lane];
    }
    if (lane == 31) { AB(0, 0) = 1.0; AB(1, 1) = 1.0; AB(2, 2) = 2.0; }   // EP:277-279
    if (lane < M) {
        const double T = m.ts[lane];
        const double T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
        if (lane < M - 1) {
            const int r = 6 * lane + 3, cb = 6 * lane;
            // jerk continuity (EP:306-309)
            AB(r, cb + 3) = 6.0; AB(r, cb + 4) = 24.0 * T; AB(r, cb + 5) = 60.0 * T2; AB(r, cb + 9) = -6.0;
            // snap continuity (EP:310-312)
            AB(r + 1, cb + 4) = 24.0; AB(r + 1, cb + 5) = 120.0 * T; AB(r + 1, cb + 10) = -24.0;
            // position at the waypoint (EP:285-290)
            AB(r + 2, cb) = 1.0; AB(r + 2, cb + 1) = T; AB(r + 2, cb + 2) = T2; AB(r + 2, cb + 3) = T3;
            AB(r + 2, cb + 4) = T4; AB(r + 2, cb + 5) = T5;
            // position continuity (EP:291-297)
            AB(r + 3, cb) = 1.0; AB(r + 3, cb + 1) = T; AB(r + 3, cb + 2) = T2; AB(r + 3, cb + 3) = T3;
            AB(r + 3, cb + 4) = T4; AB(r + 3, cb + 5) = T5; AB(r + 3, cb + 6) = -1.0;
            // velocity continuity (EP:298-303)
            AB(r + 4, cb + 1) = 1.0; AB(r + 4, cb + 2) = 2.0 * T; AB(r + 4, cb + 3) = 3.0 * T2;
            AB(r + 4, cb + 4) = 4.0 * T3; AB(r + 4, cb + 5) = 5.0 * T4; AB(r + 4, cb + 7) = -1.0;
            // acceleration continuity (EP:304-308)
            AB(r + 5, cb + 2) = 2.0; AB(r + 5, cb + 3) = 6.0 * T; AB(r + 5, cb + 4) = 12.0 * T2;
            AB(r + 5, cb + 5) = 20.0 * T3; AB(r + 5, cb + 8) = -2.0;
        } else {       // tail rows (EP:318-332)
            const int r = N - 3, cb = N - 6;
            AB(r, cb) = 1.0; AB(r, cb + 1) = T; AB(r, cb + 2) = T2; AB(r, cb + 3) = T3; AB(r, cb + 4) = T4;
            AB(r, cb + 5) = T5;
            AB(r + 1, cb + 1) = 1.0; AB(r + 1, cb + 2) = 2.0 * T; AB(r + 1, cb + 3) = 3.0 * T2;
            AB(r + 1, cb + 4) = 4.0 * T3; AB(r + 1, cb + 5) = 5.0 * T4;
            AB(r + 2, cb + 2) = 2.0; AB(r + 2, cb + 3) = 6.0 * T; AB(r + 2, cb + 4) = 12.0 * T2;
            AB(r + 2, cb + 5) = 20.0 * T3;
        }
    }
    __syncwarp();
}

// In-place banded LU (no pivoting) fused with the forward elimination of the two right-hand sides.
// Lanes 0..23: row offset ii = lane/4 (6 rows below the pivot), column set {jq, jq+4}, jq = lane%4, out of
// 8 columns = 6 band columns right of the pivot + the 2 rhs columns.
__device__ __forceinline__ void factor_and_forward(const WarpMem &m, int M, int lane)
{
    const int N = 6 * M;
    double *Ab = m.Ab;
    const int ii = lane >> 2, jq = lane & 3;
    for (int k = 0; k < N; k++) {
        const double rk = 1.0 / AB(k, k);
        if (lane == 31) m.rinv[k] = rk;
        const int i = k + 1 + ii;
        double l = 0.0, v0 = 0.0, v1 = 0.0;
        const bool act = lane < 24 && i < N;
        const int j0 = k + 1 + jq;                 // band column (jq < 4 < 6)
        const bool c0 = act && j0 < N;
        const bool c1band = act && jq < 2 && (j0 + 4) < N;   // second column is a band column
        const bool c1rhs = act && jq >= 2;                     // second column is rhs dim jq-2
        if (act) {
            l = AB(i, k) * rk;
            if (c0) v0 = AB(i, j0) - l * AB(k, j0);
            if (c1band) v1 = AB(i, j0 + 4) - l * AB(k, j0 + 4);
            if (c1rhs) v1 = m.c[i * 2 + (jq - 2)] - l * m.c[k * 2 + (jq - 2)];
        }
        __syncwarp();
        if (act) {
            if (c0) AB(i, j0) = v0;
            if (c1band) AB(i, j0 + 4) = v1;
            if (c1rhs) m.c[i * 2 + (jq - 2)] = v1;
            if (jq == 0) AB(i, k) = l;
        }
        __syncwarp();
    }
}

// U x = y (column sweep, 12 lanes: 6 rows above x 2 dims); v: [N][2]
__device__ __forceinline__ void solve_U(const WarpMem &m, int M, int lane, double *v)
{
    const int N = 6 * M;
    const double *Ab = m.Ab;
    const int jj = lane >> 1, d = lane & 1;
    for (int k = N - 1; k >= 0; k--) {
        const double xk = v[k * 2 + d] * m.rinv[k];
        const int i = k - 1 - jj;
        double nv = 0.0;
        const bool act = lane < 12 && i >= 0;
        if (act) nv = v[i * 2 + d] - AB(i, k) * xk;
        __syncwarp();
        if (lane < 2) v[k * 2 + d] = xk;
        if (act) v[i * 2 + d] = nv;
        __syncwarp();
    }
}

// U^T w = g (forward column sweep); v: [N][2]
__device__ __forceinline__ void solve_UT(const WarpMem &m, int M, int lane, double *v)
{
    const int N = 6 * M;
    const double *Ab = m.Ab;
    const int jj = lane >> 1, d = lane & 1;
    for (int k = 0; k < N; k++) {
        const double wk = v[k * 2 + d] * m.rinv[k];
        const int i = k + 1 + jj;
        double nv = 0.0;
        const bool act = lane < 12 && i < N;
        if (act) nv = v[i * 2 + d] - AB(k, i) * wk;
        __syncwarp();
        if (lane < 2) v[k * 2 + d] = wk;
        if (act) v[i * 2 + d] = nv;
        __syncwarp();
    }
}

// L^T z = w (backward column sweep, unit diagonal); v: [N][2]
__device__ __forceinline__ void solve_LT(const WarpMem &m, int M, int lane, double *v)
{
    const int N = 6 * M;
    const double *Ab = m.Ab;
    const int jj = lane >> 1, d = lane & 1;
    for (int k = N - 1; k > 0; k--) {
        const double zk = v[k * 2 + d];
        const int i = k - 1 - jj;
        double nv = 0.0;
        const bool act = lane < 12 && i >= 0;
        if (act) nv = v[i * 2 + d] - AB(k, i) * zk;
        __syncwarp();
        if (act) v[i * 2 + d] = nv;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// fused cost + gradient
// ---------------------------------------------------------------------------------------------------------
struct EvalOut {
    double f;          // dot(costs, weights) (EP:558)
    double costs[4];   // unweighted (EP:339)
    double g;          // lane l < n: grad[l] (EP:581)
    int status;        // 0, NEO_ST_OVERFLOW (4) or NEO_ST_NAN (6)
    unsigned ns, nv, nc;   // samples, velocity-violating samples, colliding samples (work accounting)
};

// permuted row of reference row 6i+3+a
__device__ __forceinline__ int perm_row(int i, int a) { return 6 * i + 3 + (a < 4 ? a + 2 : a - 4); }

__device__ __forceinline__ void eval_fg(const DevParams &P, const MapView &map, const WarpMem &m, int M, int lane,
                                        double xl, bool want_grad, EvalOut &out)
{
    const int N = 6 * M, nq = 2 * (M - 1);
    out.status = 0; out.ns = out.nv = out.nc = 0;
    out.g = 0.0;

    // ---- tau -> T (EP:477-483) --------------------------------------------------------------------
    const double tau = __shfl_sync(FULL, xl, (nq + lane) & 31);
    int bad = 0;
    double e = 0.0, T = 1.0;
    if (lane < M) {
        bool ovf;
        e = exp_dd(-tau, &ovf);
        const double den = (1.0 + e) * (1.0 + e);          // (1+exp(-tau))**2 raises OverflowError (EP:490)
        if (ovf || den == INFINITY) bad = 4;
        T = (P.T_max - P.T_min) / (1.0 + e) + P.T_min;
        if (T != T) bad = 6;                               // int(nan) raises ValueError (EP:401)
        m.ts[lane] = T;
        m.ex[lane] = e;
    }
    bad = __reduce_max_sync(FULL, bad);
    if (bad) { out.status = bad; out.f = 0.0; out.costs[0] = out.costs[1] = out.costs[2] = out.costs[3] = 0.0; return; }
    __syncwarp();

    // ---- coefficients (EP:261-336) ----------------------------------------------------------------
    build_system(m, M, lane, xl);
    factor_and_forward(m, M, lane);
    solve_U(m, M, lane, m.c);

    // ---- energy + time (EP:345-390): lane i < M owns piece i ----------------------------------------
    double cost0 = 0.0, cost1 = 0.0;
    if (lane < M) {
        const double T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
        const double *ci = m.c + 12 * lane;
        double gt = P.w1;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const double c3 = ci[6 + d], c4 = ci[8 + d], c5 = ci[10 + d];
            const double r3 = 36.0 * T * c3 + 72.0 * T2 * c4 + 120.0 * T3 * c5;     // rows of beta3_mat @ c
            const double r4 = 72.0 * T2 * c3 + 192.0 * T3 * c4 + 360.0 * T4 * c5;
            const double r5 = 120.0 * T3 * c3 + 360.0 * T4 * c4 + 720.0 * T5 * c5;
            cost0 += c3 * r3 + c4 * r4 + c5 * r5;
            m.gC[12 * lane + 0 + d] = 0.0; m.gC[12 * lane + 2 + d] = 0.0; m.gC[12 * lane + 4 + d] = 0.0;
            m.gC[12 * lane + 6 + d] = (P.w0 * 2.0) * r3;
            m.gC[12 * lane + 8 + d] = (P.w0 * 2.0) * r4;
            m.gC[12 * lane + 10 + d] = (P.w0 * 2.0) * r5;
            const double jend = 6.0 * c3 + 24.0 * T * c4 + 60.0 * T2 * c5;         // jerk at the piece end
            gt += P.w0 * (jend * jend);
        }
        m.gT[lane] = gt;
        cost1 = T;
    }
    __syncwarp();
    double costs0 = 0.0, costs1 = 0.0;
    for (int i = 0; i < M; i++) {          // sequential, like the reference's loops / np.sum
        costs0 += __shfl_sync(FULL, cost0, i);
        costs1 += __shfl_sync(FULL, cost1, i);
    }

    // ---- sampled penalties (EP:392-466): lanes over the samples of one piece at a time ---------------
    double costs2 = 0.0, costs3 = 0.0;
    for (int i = 0; i < M; i++) {
        const double Ti = m.ts[i];
        const int ns = (int)(Ti / P.dt);                    // int(T/delta_t) (EP:401)
        const double *ci = m.c + 12 * i;
        double cx[6], cy[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { cx[k] = ci[2 * k]; cy[k] = ci[2 * k + 1]; }
        double acc[16];
#pragma unroll
        for (int s = 0; s < 16; s++) acc[s] = 0.0;
        const double inv_ns = 1.0;   // divisions kept explicit below to follow the reference's rounding
        (void)inv_ns;
        for (int j0 = 0; j0 < ns; j0 += 32) {
            const int j = j0 + lane;
            const bool live = j < ns;
            const double t = (double)j * P.dt;               // np.arange(0, T_max, dt)[j] (EP:251)
            const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
            const double px = cx[0] + cx[1] * t + cx[2] * t2 + cx[3] * t3 + cx[4] * t4 + cx[5] * t5;
            const double py = cy[0] + cy[1] * t + cy[2] * t2 + cy[3] * t3 + cy[4] * t4 + cy[5] * t5;
            const double b1[6] = {0.0, 1.0, 2.0 * t, 3.0 * t2, 4.0 * t3, 5.0 * t4};
            const double vx = cx[1] + cx[2] * b1[2] + cx[3] * b1[3] + cx[4] * b1[4] + cx[5] * b1[5];
            const double vy = cy[1] + cy[2] * b1[2] + cy[3] * b1[3] + cy[4] * b1[4] + cy[5] * b1[5];
            const double omg = (j == 0 || j == ns - 1) ? 0.5 : 1.0;   // EP:407
            // feasibility (EP:409-413, EP:441-451)
            const double vv = (vx * vx + vy * vy) - P.v_max2;
            const bool viol_v = live && vv > 0.0;
            // collision (EP:415-422, EP:453-466): nearest-cell lookup, trunc-toward-zero index (ESDF:61-65)
            const double fr = (py - map.oy) / map.res;
            const double fc = (px - map.ox) / map.res;
            if (live && (fr != fr || fc != fc)) bad = 6;       // int(nan) raises ValueError
            const double tr = trunc(fr), tc = trunc(fc);
            const bool inside = live && tr >= 0.0 && tr < (double)map.H && tc >= 0.0 && tc < (double)map.W;
            double dis = 10000.0;
            const Cell *cell = map.cells;
            if (inside) {
                cell = map.cells + ((size_t)(int)tr * map.W + (int)tc);
                dis = __ldg(&cell->d);
            }
            const double vd = P.safe_dis - dis;
            const bool viol_d = inside && vd > 0.0;
            out.ns += live ? 1u : 0u;
            if (viol_v) {
                const double vv2 = vv * vv, vv3 = vv2 * vv;
                acc[13] += (omg * P.dt) * vv3;
                if (want_grad) {
                    const double K = (3.0 * P.dt * omg) * vv2;
                    const double ax = 2.0 * cx[2] + 6.0 * t * cx[3] + 12.0 * t2 * cx[4] + 20.0 * t3 * cx[5];
                    const double ay = 2.0 * cy[2] + 6.0 * t * cy[3] + 12.0 * t2 * cy[4] + 20.0 * t3 * cy[5];
                    const double v2t = 2.0 * (ax * vx + ay * vy);
                    const double kx = (P.w2 * K) * (2.0 * vx), ky = (P.w2 * K) * (2.0 * vy);
#pragma unroll
                    for (int k = 1; k < 6; k++) { acc[2 * k] += b1[k] * kx; acc[2 * k + 1] += b1[k] * ky; }
                    acc[12] += P.w2 * (omg * vv3 / (double)ns + K * v2t * (double)j / (double)ns);
                }
                out.nv++;
            }
            if (viol_d) {
                const double vd2 = vd * vd, vd3 = vd2 * vd;
                acc[14] += (omg * P.dt) * vd3;
                if (want_grad) {
                    const double2 g = __ldg(reinterpret_cast<const double2 *>(cell));
                    const double K = (3.0 * P.dt * omg) * vd2;
                    const double p2t = -(g.x * vx + g.y * vy);
                    const double kx = -(P.w3 * K) * g.x, ky = -(P.w3 * K) * g.y;
                    const double b0[6] = {1.0, t, t2, t3, t4, t5};
#pragma unroll
                    for (int k = 0; k < 6; k++) { acc[2 * k] += b0[k] * kx; acc[2 * k + 1] += b0[k] * ky; }
                    acc[12] += P.w3 * (omg * vd3 / (double)ns + K * p2t * (double)j / (double)ns);
                }
                out.nc++;
            }
        }
        warp_reduce16(acc, lane);
        const int slot = lane >> 1;
        const double tot = acc[0];
        if (!(lane & 1)) {
            if (slot < 12) m.gC[12 * i + slot] += tot;
            else if (slot == 12) m.gT[i] += tot;
        }
        costs2 += __shfl_sync(FULL, tot, 26);
        costs3 += __shfl_sync(FULL, tot, 28);
    }
    bad = __reduce_max_sync(FULL, bad);
    out.ns = __reduce_add_sync(FULL, out.ns);
    out.nv = __reduce_add_sync(FULL, out.nv);
    out.nc = __reduce_add_sync(FULL, out.nc);
    out.costs[0] = costs0; out.costs[1] = costs1; out.costs[2] = costs2; out.costs[3] = costs3;
    out.f = costs0 * P.w0 + costs1 * P.w1 + costs2 * P.w2 + costs3 * P.w3;
    if (bad) { out.status = bad; return; }
    if (!want_grad) return;
    __syncwarp();

    // ---- adjoint (EP:494-537): A^T G = dW/dc with P A = L U  =>  U^T L^T (P G) = dW/dc -----------------
    solve_UT(m, M, lane, m.gC);
    solve_LT(m, M, lane, m.gC);
    const double *z = m.gC;     // z[perm_row] = G[reference row]
    double g_out = 0.0;
    if (lane < nq) {             // grad_q[d][i] = G[6i+3][d] (EP:506-508)
        const int d = lane / (M - 1), i = lane - d * (M - 1);
        g_out = z[perm_row(i, 0) * 2 + d];
    }
    // grad_T (EP:511-533): piece i < M-1 uses T_i; the last piece re-uses the loop variable T = ts[M-2]
    double gtau = 0.0;
    if (lane < M) {
        const int i = lane;
        const double Tq = (i < M - 1) ? T : m.ts[M - 2];
        const double T2 = Tq * Tq, T3 = T2 * Tq, T4 = T3 * Tq;
        const double *ci = m.c + 12 * i;
        double tr = 0.0;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const double c1 = ci[2 + d], c2 = ci[4 + d], c3 = ci[6 + d], c4 = ci[8 + d], c5 = ci[10 + d];
            const double vel = c1 + 2.0 * Tq * c2 + 3.0 * T2 * c3 + 4.0 * T3 * c4 + 5.0 * T4 * c5;
            const double ac = 2.0 * c2 + 6.0 * Tq * c3 + 12.0 * T2 * c4 + 20.0 * T3 * c5;
            const double jr = 6.0 * c3 + 24.0 * Tq * c4 + 60.0 * T2 * c5;
            if (i < M - 1) {
                const double sn = 24.0 * c4 + 120.0 * Tq * c5;
                const double cr = 120.0 * c5;
                tr += (z[perm_row(i, 0) * 2 + d] + z[perm_row(i, 1) * 2 + d]) * vel + z[perm_row(i, 2) * 2 + d] * ac
                    + z[perm_row(i, 3) * 2 + d] * jr + z[perm_row(i, 4) * 2 + d] * sn + z[perm_row(i, 5) * 2 + d] * cr;
            } else {
                tr += z[(N - 3) * 2 + d] * vel + z[(N - 2) * 2 + d] * ac + z[(N - 1) * 2 + d] * jr;
            }
        }
        const double gTi = m.gT[i] - tr;
        gtau = gTi * (P.T_max - P.T_min) * e / ((1.0 + e) * (1.0 + e));      // EP:485-492
    }
    // place grad_tau[i] (held by lane i) into lane nq+i
    const double gt_sh = __shfl_sync(FULL, gtau, (lane - nq) & 31);
    if (lane >= nq && lane < nq + M) g_out = gt_sh;
    out.g = g_out;
    __syncwarp();
}

#undef AB

}  // namespace neo
