// lbfgs_warp.cuh -- one warp runs scipy.optimize.minimize(method='L-BFGS-B', bounds=None, tol=1e-4,
// maxcor=10, maxls=20) exactly as the reference calls it (EP:213-225), with the whole optimizer state on chip:
// lane l < n owns component l of x, g, d and of the saved iterate; the (s, y) history lives in the warp's
// shared-memory slice; scalar logic (Moré-Thuente dcsrch/dcstep of MINPACK-2, as used by L-BFGS-B 3.0's
// lnsrlb) is executed redundantly and identically by all 32 lanes. No host round trips.
//
// Behaviour restated from scipy 1.18.1 (third-party; `scipy/optimize/_lbfgsb_py.py:336-470`, `_dcsrch.py`):
//   * ftol = gtol = 1e-4 (tol), factr = ftol/eps; stop if max|g| <= 1e-4 or (f_old-f) <= 1e-4*max(|f_old|,|f|,1)
//   * direction d = -H g (two-loop recursion, H0 = I/theta, theta = y'y/s'y of the newest pair)
//   * first step min(1/|d|, 1e10) on iteration 0, else 1; dcsrch(ftol=1e-3, gtol=0.9, xtol=0.1, stpmin=0,
//     stpmax=1e10); CONVERGENCE and WARNING exits are both accepted; a 21st evaluation request fails the search
//   * on a failed search: restore x, f, g; if the memory is empty -> ABNORMAL, else drop the memory and retry
//   * the pair (s, y) is stored unless s'y <= eps * (-g_old's)
//   * scipy's ScalarFunction does not re-evaluate an x identical to the last evaluated one (nfev bookkeeping)
#pragma once
#include "minco_warp.cuh"

namespace neo {

struct Dcsrch {
    bool brackt;
    int stage;
    double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
};

#define LS_FTOL 1e-3
#define LS_GTOL 0.9
#define LS_XTOL 0.1
#define LS_STPMIN 0.0
#define LS_STPMAX 1e10

__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(a, fmax(b, c)); }

__device__ inline void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                              double fp, double dp, bool &brackt, double stpmin, double stpmax)
{
    const double sgnd = dp * (dx / fabs(dx));
    double stpf, stpc, stpq, theta, s, gamma, p, q, r;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta; q = ((gamma - dx) + gamma) + dp; r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc; else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dx; r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta; q = (gamma + (dx - dp)) + gamma; r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
            stpf = fmin(stpmax, stpf); stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = max3(fabs(theta), fabs(dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dy; r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

__device__ __forceinline__ void dcsrch_start(Dcsrch &S, double stp, double f, double g)
{
    S.brackt = false; S.stage = 1; S.finit = f; S.ginit = g; S.gtest = LS_FTOL * g;
    S.width = LS_STPMAX - LS_STPMIN; S.width1 = S.width / 0.5;
    S.stx = 0.0; S.fx = f; S.gx = g; S.sty = 0.0; S.fy = f; S.gy = g;
    S.stmin = 0.0; S.stmax = stp + 4.0 * stp;
}

// 0 = evaluate at stp, 1 = CONVERGENCE, 2 = WARNING
__device__ inline int dcsrch_step(Dcsrch &S, double &stp, double f, double g)
{
    const double ftest = S.finit + stp * S.gtest;
    int task = 0;
    if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
    if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) task = 2;
    if (S.brackt && S.stmax - S.stmin <= LS_XTOL * S.stmax) task = 2;
    if (stp == LS_STPMAX && f <= ftest && g <= S.gtest) task = 2;
    if (stp == LS_STPMIN && (f > ftest || g >= S.gtest)) task = 2;
    if (f <= ftest && fabs(g) <= LS_GTOL * (-S.ginit)) task = 1;
    if (task) return task;
    {   // a modified function is used in stage 1 while the decrease is not yet sufficient
        const bool mod = S.stage == 1 && f <= S.fx && f > ftest;
        const double gt = mod ? S.gtest : 0.0;
        const double fm = f - stp * gt, gm = g - gt;
        double fxm = S.fx - S.stx * gt, fym = S.fy - S.sty * gt, gxm = S.gx - gt, gym = S.gy - gt;
        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
        S.fx = fxm + S.stx * gt; S.fy = fym + S.sty * gt;
        S.gx = gxm + gt; S.gy = gym + gt;
    }
    if (S.brackt) {
        if (fabs(S.sty - S.stx) >= 0.66 * S.width1) stp = S.stx + 0.5 * (S.sty - S.stx);
        S.width1 = S.width; S.width = fabs(S.sty - S.stx);
    }
    if (S.brackt) { S.stmin = fmin(S.stx, S.sty); S.stmax = fmax(S.stx, S.sty); }
    else { S.stmin = stp + 1.1 * (stp - S.stx); S.stmax = stp + 4.0 * (stp - S.stx); }
    stp = fmax(stp, LS_STPMIN); stp = fmin(stp, LS_STPMAX);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= LS_XTOL * S.stmax))
        stp = S.stx;
    return 0;
}

struct OptOut {
    double x;          // lane l < n: final x[l]
    double costs[4];   // at the last evaluated point (EP:233)
    int status, nit, nfev;
    unsigned long long ns, nv, nc;
};

// plan_once's minimize() (EP:213-225). x0l: lane l < n holds x0[l].
// Written as a state machine around ONE evaluation site so the (large) fused evaluator is instantiated once.
// cancel_word/cancel_mask: when (*cancel_word & cancel_mask) becomes non-zero (an earlier attempt of the same problem
// has been accepted, so this speculative attempt can never be the returned one) the run stops with ST_CANCELLED.
constexpr int ST_CANCELLED = 7;

// ---- speculative restarts (EXPERIMENTAL, off by default: NEO_SHADOW=1; protocol = oracle/shadow_sim.c) --------------
// The owner of a task publishes, at the start of every line search that has memory, the state the optimizer would
// restart from if that search failed. A warp with nothing else to do claims the request and computes the restart
// speculatively; a successful search cancels it (epoch change), a failed one hands the task over to it.
constexpr int ST_SUPERSEDED = 8;     // a speculative restart that was not needed (never reported)
constexpr int ST_HANDED_OFF = 9;     // the owner retired in favour of its claimant (never reported)
enum { SL_IDLE = 0, SL_REQUESTED = 1, SL_CLAIMED = 2, SL_CONFIRMED = 3 };
__device__ __forceinline__ unsigned sl_ctl(unsigned epoch, unsigned state) { return epoch * 4u + state; }

struct ShadowSlot {                  // one per task (attempt * B + problem), zero-initialised per launch
    unsigned ctl;                    // epoch * 4 + state
    int nit, nfev, nfev_after;       // counters at the start of the search / at its failure (hand-off)
    unsigned long long ns, nv, nc, nanos;                          // work counters at the start of the search
    unsigned long long ns_after, nv_after, nc_after, nanos_after;  // ... at the failure
    double f;
    double costs[4];
    double x[32], g[32];
};

struct ShadowCtx {                   // warp-uniform
    ShadowSlot *slot;
    unsigned *req_word;              // discovery bitmap (one bit per task with a request out; ctl stays authoritative)
    unsigned req_bit;
    unsigned epoch;                  // owner: last published epoch; claimant: the epoch it claimed
    bool owner;                      // true: owns the task and publishes; false: speculative
    bool published;                  // owner: the running line search has a request out
    unsigned long long t_start;      // %globaltimer when this warp started on the task
    unsigned long long nanos_base;   // time earlier owners spent on the task
};

__device__ __forceinline__ unsigned sl_load(const ShadowSlot *slot)
{
    return *reinterpret_cast<const volatile unsigned *>(&slot->ctl);
}

template <int MODE, bool SHADOW = false>
__device__ __forceinline__ void lbfgsb_warp(const DevParams &P, const MapView &map, const WarpMem &m, int M, int lane,
                                   double x0l, OptOut &o, const unsigned *cancel_word = nullptr,
                                   unsigned cancel_mask = 0u, bool lockstep = false, ShadowCtx *sc = nullptr)
{
    const int n = 3 * M - 2;
    const bool mine = lane < n;
    const double pgtol = 1e-4, ftol = 1e-4, epsmch = 2.220446049250313e-16;
    const double tol = (ftol / epsmch) * epsmch;
    const int maxls = 20, maxiter = 15000, maxfun = 15000;
    double x = mine ? x0l : 0.0, g = 0.0, d = 0.0, t = 0.0, r = 0.0, xlast = 0.0;
    double f = 0.0, fold = 0.0, theta = 1.0, stp = 0.0, gd = 0.0, gdold = 0.0;
    int col = 0, head = 0, nit = 0, nfev = 0, st = 0, ifun = 0;
    bool first = true;
    Dcsrch ls;
    o.ns = o.nv = o.nc = 0;
    bool resume = false;
    if constexpr (SHADOW) {
        if (!sc->owner) {
            // claimant: continue from the published restart state (memory empty, no evaluation at the start point)
            const ShadowSlot *sl = sc->slot;
            x = mine ? __ldcg(&sl->x[lane]) : 0.0; g = mine ? __ldcg(&sl->g[lane]) : 0.0;
            f = __ldcg(&sl->f); nit = __ldcg(&sl->nit); nfev = __ldcg(&sl->nfev);
            o.ns = __ldcg(&sl->ns); o.nv = __ldcg(&sl->nv); o.nc = __ldcg(&sl->nc);
            sc->nanos_base = __ldcg(&sl->nanos);
#pragma unroll
            for (int k = 0; k < 4; k++) o.costs[k] = __ldcg(&sl->costs[k]);
            __threadfence();
            const unsigned c = sl_load(sl);           // seqlock: the copy counts only if the owner has not moved on
            if (c != sl_ctl(sc->epoch, SL_CLAIMED) && c != sl_ctl(sc->epoch, SL_CONFIRMED)) { o.status = ST_SUPERSEDED; return; }
            xlast = __longlong_as_double(0x7ff8000000000000LL);       // the failed search's last trial point is unknown
            first = false; resume = true;
        }
    }
    for (;;) {
        // Large batches: the warps of a CTA meet here before every evaluation so that they walk through the (~60 KB)
        // evaluator together and share its instruction fetches (the SM's instruction cache holds 32 KB).
        if (lockstep) __syncthreads_and(0);
        // ---- the evaluation site: f, g at x (skipped when x is bit-identical to the last evaluated point) ------
        const bool same = !first && __all_sync(FULL, !mine || x == xlast);
        if constexpr (SHADOW) {
            if (!resume && !same && !sc->owner) {
                // claimant: still wanted? (CLAIMED: the owner's search is running; CONFIRMED: it failed, the task is ours)
                const unsigned c = sl_load(sc->slot);
                if (c == sl_ctl(sc->epoch, SL_CONFIRMED)) {
                    __threadfence();
                    const ShadowSlot *sl = sc->slot;
                    nfev += __ldcg(&sl->nfev_after) - __ldcg(&sl->nfev);
                    o.ns += __ldcg(&sl->ns_after) - __ldcg(&sl->ns); o.nv += __ldcg(&sl->nv_after) - __ldcg(&sl->nv);
                    o.nc += __ldcg(&sl->nc_after) - __ldcg(&sl->nc);
                    sc->nanos_base = __ldcg(&sl->nanos_after);
                    sc->owner = true; sc->published = false;
                } else if (c != sl_ctl(sc->epoch, SL_CLAIMED)) { o.status = ST_SUPERSEDED; return; }
            }
        }
        if (!same && !resume) {
            EvalOut ev;
            eval_fg<MODE>(P, map, m, M, lane, x, true, ev);
            o.ns += ev.ns; o.nv += ev.nv; o.nc += ev.nc;
            if constexpr (SHADOW) {
                if (ev.status) { st = ev.status; goto done; }      // the exit below handles requests and pending verdicts
            }
            if (ev.status) { o.status = ev.status; o.nit = nit; o.nfev = nfev; o.x = x; return; }
            f = ev.f; g = mine ? ev.g : 0.0; nfev++; xlast = x;
#pragma unroll
            for (int k = 0; k < 4; k++) o.costs[k] = ev.costs[k];
        }
        bool new_dir;
        if (SHADOW && resume) { resume = false; new_dir = true; }
        else if (first) {
            first = false;
            if (warp_max(fabs(g)) <= pgtol) { st = 1; break; }
            new_dir = true;
        } else {
            gd = warp_sum(g * d);
            if (dcsrch_step(ls, stp, f, gd) == 0) new_dir = false;        // FG: another trial point
            else {
                // ---- the line search accepted the last evaluated point ---------------------------------------
                if constexpr (SHADOW) {
                    if (sc->owner && sc->published) {       // this epoch is over: a claimant sees the change and stops
                        if (lane == 0) { atomicExch(&sc->slot->ctl, sl_ctl(sc->epoch, SL_IDLE)); atomicAnd(sc->req_word, ~sc->req_bit); }
                        sc->published = false;
                    }
                }
                nit++;
                if (cancel_mask && (*reinterpret_cast<const volatile unsigned *>(cancel_word) & cancel_mask)) {
                    st = ST_CANCELLED; break;
                }
                if (warp_max(fabs(g)) <= pgtol) { st = 1; break; }
                if (fold - f <= tol * max3(fabs(fold), fabs(f), 1.0)) { st = 0; break; }
                if (nit >= maxiter || nfev > maxfun) { st = 3; break; }
                const double y = g - r;                                     // matupd
                const double rr = warp_sum(y * y);
                double dr, ddum, s;
                if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; s = d; }
                else { dr = (gd - gdold) * stp; s = d * stp; ddum = -gdold * stp; }
                if (!(dr <= epsmch * ddum)) {
                    int slot;
                    if (col < HIST) { slot = (head + col) % HIST; col++; }
                    else { slot = head; head = (head + 1) % HIST; }
                    if (mine) { m.S[slot * n + lane] = s; m.Y[slot * n + lane] = y; }
                    if (lane == 0) m.rho[slot] = 1.0 / dr;
                    theta = rr / dr;
                    __syncwarp();
                }
                new_dir = true;
            }
        }
        for (;;) {      // (re)start a line search; loops only when a failed search drops the memory
            if (new_dir) {
                // ---- d = -H g (two-loop recursion, H0 = I/theta) --------------------------------------------
                double q = g;
#pragma unroll 1
                for (int k = col - 1; k >= 0; k--) {
                    const int j = (head + k) % HIST;
                    const double sj = mine ? m.S[j * n + lane] : 0.0, yj = mine ? m.Y[j * n + lane] : 0.0;
                    const double al = m.rho[j] * warp_sum(sj * q);
                    if (lane == 0) m.al[k] = al;
                    q -= al * yj;
                }
                if (theta != 1.0) q /= theta;
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < col; k++) {
                    const int j = (head + k) % HIST;
                    const double sj = mine ? m.S[j * n + lane] : 0.0, yj = mine ? m.Y[j * n + lane] : 0.0;
                    const double b = m.rho[j] * warp_sum(yj * q);
                    q += (m.al[k] - b) * sj;
                }
                d = -q;
                // ---- lnsrlb: set up the search -----------------------------------------------------------------
                const double dnorm = sqrt(warp_sum(d * d));
                stp = (nit == 0) ? fmin(1.0 / dnorm, LS_STPMAX) : 1.0;
                t = x; r = g; fold = f;
                if constexpr (SHADOW) {
                    if (sc->owner && col > 0) {
                        ShadowSlot *sl = sc->slot;
                        if (mine) { sl->x[lane] = x; sl->g[lane] = g; }
                        if (lane < 4) sl->costs[lane] = o.costs[lane];
                        if (lane == 0) {
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            sl->f = f; sl->nit = nit; sl->nfev = nfev; sl->ns = o.ns; sl->nv = o.nv; sl->nc = o.nc;
                            sl->nanos = sc->nanos_base + (now - sc->t_start);
                        }
                        __threadfence();
                        __syncwarp();
                        sc->epoch++;
                        if (lane == 0) { atomicExch(&sl->ctl, sl_ctl(sc->epoch, SL_REQUESTED)); atomicOr(sc->req_word, sc->req_bit); }
                        sc->published = true;
                    }
                }
                gd = warp_sum(g * d);
                gdold = gd;
                ifun = 0;
                if (gd >= 0.0) ifun = maxls + 1;                            // not a descent direction: fail
                else dcsrch_start(ls, stp, f, gd);
            }
            ifun++;
            if (ifun - 1 < maxls) break;                                    // evaluate the trial point
            // ---- failed search: restore the iterate; ABNORMAL if the memory is already empty ------------------
            if constexpr (SHADOW) {
                if (sc->owner && sc->published) {
                    sc->published = false;
                    ShadowSlot *sl = sc->slot;
                    unsigned old = 0;
                    if (lane == 0) {
                        old = atomicCAS(&sl->ctl, sl_ctl(sc->epoch, SL_REQUESTED), sl_ctl(sc->epoch, SL_IDLE));
                        atomicAnd(sc->req_word, ~sc->req_bit);
                    }
                    old = __shfl_sync(FULL, old, 0);
                    if (old != sl_ctl(sc->epoch, SL_REQUESTED)) {
                        // CLAIMED: the claimant has been computing this restart; give it the task and retire
                        if (lane == 0) {
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            sl->nfev_after = nfev; sl->ns_after = o.ns; sl->nv_after = o.nv; sl->nc_after = o.nc;
                            sl->nanos_after = sc->nanos_base + (now - sc->t_start);
                            __threadfence();
                            atomicExch(&sl->ctl, sl_ctl(sc->epoch, SL_CONFIRMED));
                        }
                        __syncwarp();
                        o.status = ST_HANDED_OFF;
                        return;
                    }
                }
            }
            x = t; g = r; f = fold;
            if (col == 0) { st = 2; goto done; }
            col = 0; head = 0; theta = 1.0;
            new_dir = true;
        }
        x = (stp == 1.0) ? (t + d) : (stp * d + t);
    }
done:
    if constexpr (SHADOW) {
        if (sc->owner && sc->published) {           // exits other than accept/fail (exception, cancel): withdraw the request
            if (lane == 0) { atomicExch(&sc->slot->ctl, sl_ctl(sc->epoch, SL_IDLE)); atomicAnd(sc->req_word, ~sc->req_bit); }
            sc->published = false;
        }
        while (!sc->owner) {
            // finished while still speculative: wait for the owner's verdict (it runs on its own resident warp)
            const unsigned c = sl_load(sc->slot);
            if (c == sl_ctl(sc->epoch, SL_CONFIRMED)) {
                __threadfence();
                const ShadowSlot *sl = sc->slot;
                nfev += __ldcg(&sl->nfev_after) - __ldcg(&sl->nfev);
                o.ns += __ldcg(&sl->ns_after) - __ldcg(&sl->ns); o.nv += __ldcg(&sl->nv_after) - __ldcg(&sl->nv);
                o.nc += __ldcg(&sl->nc_after) - __ldcg(&sl->nc);
                sc->nanos_base = __ldcg(&sl->nanos_after);
                sc->owner = true;
                break;
            }
            if (c != sl_ctl(sc->epoch, SL_CLAIMED)) { o.status = ST_SUPERSEDED; return; }
            if (cancel_mask && (*reinterpret_cast<const volatile unsigned *>(cancel_word) & cancel_mask)) {
                o.status = ST_SUPERSEDED;           // the owner stops at its own cancel check and reports the task
                return;
            }
            __nanosleep(200);
        }
    }
    o.x = x; o.status = st; o.nit = nit; o.nfev = nfev;
}

}  // namespace neo
