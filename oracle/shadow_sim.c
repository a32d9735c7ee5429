/* shadow_sim.c -- TEST INFRASTRUCTURE ONLY. Protocol simulator for "speculative restarts" (DESIGN.md section 9, item 1).
 *
 * Threads play the persistent warps of k_optimize: they pull (attempt, problem) tasks from an attempt-major queue, run
 * plan_once with the CPU checker's L-BFGS-B core, and -- once the queue is empty -- look for restart states that running
 * tasks publish at the start of every line search that has memory, claim one and compute the restart speculatively.
 * If the owner's line search fails it hands the task over to the claimant; if it succeeds the claimant is cancelled.
 * The simulator exists to check the protocol (slot states, epochs, hand-off of counters, no lost or duplicated task
 * results, no deadlock) before it is written in CUDA: its outputs must equal sequential orc_plan_batch bit for bit.
 *
 *   gcc -O2 -fPIC -std=c11 -ffp-contract=off -pthread -shared -o libshadow_sim.so shadow_sim.c -lm
 */
#define _GNU_SOURCE
#include "minco_oracle.c"

#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>

enum { SL_IDLE = 0, SL_REQUESTED = 1, SL_CLAIMED = 2, SL_CONFIRMED = 3 };
#define CTL(epoch, state) ((unsigned)(epoch) * 4u + (unsigned)(state))

typedef struct {
    _Atomic unsigned ctl;            /* epoch * 4 + state */
    /* payload, written by the owner before REQUESTED(epoch) */
    double x[ORC_MAXN], g[ORC_MAXN], f, costs[4];
    int nit, nfev;
    /* hand-off, written by the owner before CONFIRMED(epoch) */
    int nfev_after;
    double costs_after[4];
} slot_t;

typedef struct {
    double x[ORC_MAXN], costs[4];
    int status, nit, nfev, accepted, ran;
} task_result;

typedef struct {
    const orc_params *p; const orc_map *map;
    int B, M, A;
    const double *head, *tail, *q0, *ts0, *retry_q, *retry_ts;
    slot_t *slots;                   /* [A * B] */
    task_result *res;                /* [A * B] */
    _Atomic unsigned *pstate;        /* [B]: bit a = attempt a done, bit 8 + a = attempt a accepted */
    _Atomic int queue, resolved;
    _Atomic long claims, handoffs, cancelled_shadows, shadow_evals;
} sim_t;

typedef struct {
    sim_t *s;
    int task, b, a;
    int owner;                       /* 1: owns the task, publishes restart states; 0: speculative */
    unsigned epoch;                  /* owner: last published epoch; shadow: the epoch it claimed */
    int nfev_delta;                  /* evaluations of failed searches this (former) shadow did not run itself */
    int published;                   /* owner: the current line search has a REQUESTED/CLAIMED slot */
} wctx;

static int lower_accepted(sim_t *s, int b, int a)
{
    const unsigned st = atomic_load(&s->pstate[b]);
    return ((st >> 8) & ((1u << a) - 1u)) != 0;
}

/* ---- hooks ------------------------------------------------------------------------------------------------ */
static void hk_ls_start(void *c, int n, const double *x, const double *g, double f, const double *costs, int nit, int nfev)
{
    wctx *w = (wctx *)c;
    if (!w->owner) return;
    slot_t *sl = &w->s->slots[w->task];
    memcpy(sl->x, x, sizeof(double) * n); memcpy(sl->g, g, sizeof(double) * n);
    sl->f = f; memcpy(sl->costs, costs, sizeof(sl->costs));
    sl->nit = nit; sl->nfev = nfev + w->nfev_delta;
    w->epoch++;
    atomic_store_explicit(&sl->ctl, CTL(w->epoch, SL_REQUESTED), memory_order_release);
    w->published = 1;
}

static void hk_ls_end(void *c)
{
    wctx *w = (wctx *)c;
    if (!w->owner || !w->published) return;
    /* whatever the state (REQUESTED or CLAIMED), this epoch is over: a claimant sees the change and stops */
    atomic_store_explicit(&w->s->slots[w->task].ctl, CTL(w->epoch, SL_IDLE), memory_order_release);
    w->published = 0;
}

static int hk_ls_fail(void *c, int nfev, const double *costs)
{
    wctx *w = (wctx *)c;
    if (!w->owner || !w->published) return 0;
    slot_t *sl = &w->s->slots[w->task];
    unsigned expect = CTL(w->epoch, SL_REQUESTED);
    w->published = 0;
    if (atomic_compare_exchange_strong(&sl->ctl, &expect, CTL(w->epoch, SL_IDLE))) return 0;   /* nobody claimed */
    /* CLAIMED(epoch): hand the task over */
    sl->nfev_after = nfev + w->nfev_delta;
    memcpy(sl->costs_after, costs, sizeof(sl->costs_after));
    atomic_store_explicit(&sl->ctl, CTL(w->epoch, SL_CONFIRMED), memory_order_release);
    atomic_fetch_add(&w->s->handoffs, 1);
    return 1;
}

static void promote(wctx *w)      /* shadow -> owner after CONFIRMED */
{
    slot_t *sl = &w->s->slots[w->task];
    w->nfev_delta = sl->nfev_after - sl->nfev;      /* the failed search's evaluations, counted by the former owner */
    w->owner = 1; w->published = 0;
}

static int hk_poll(void *c)
{
    wctx *w = (wctx *)c;
    if (lower_accepted(w->s, w->b, w->a)) return 1;                  /* the retry speculation's cancel word */
    if (w->owner) return 0;
    const unsigned ctl = atomic_load_explicit(&w->s->slots[w->task].ctl, memory_order_acquire);
    if (ctl == CTL(w->epoch, SL_CLAIMED)) { atomic_fetch_add(&w->s->shadow_evals, 1); return 0; }
    if (ctl == CTL(w->epoch, SL_CONFIRMED)) { promote(w); return 0; }
    return 1;                                                        /* the owner's search succeeded */
}

/* ---- task completion (what the finishing warp does in k_optimize) ----------------------------------------------- */
static void finalize(sim_t *s, wctx *w, const orc_result *r, int raised_before_minimize, int st0)
{
    const int n = ORC_D * (s->M - 1) + s->M;
    task_result *tr = &s->res[w->task];
    tr->ran = 1;
    if (raised_before_minimize) { tr->status = st0; tr->nit = 0; tr->nfev = 0; tr->accepted = 0; }
    else {
        tr->status = r->status; tr->nit = r->nit; tr->nfev = r->nfev + w->nfev_delta;
        memcpy(tr->x, r->x, sizeof(double) * n); memcpy(tr->costs, r->costs, sizeof(tr->costs));
        double ts[ORC_MAXM];
        tr->accepted = r->status < ORC_OVERFLOW && orc_tau2T(s->p, s->M, r->x + ORC_D * (s->M - 1), ts) == 0 &&
                       !(r->costs[3] * s->p->w[3] > s->p->collision_cost_tol);
        if (r->status < ORC_OVERFLOW && orc_tau2T(s->p, s->M, r->x + ORC_D * (s->M - 1), ts) != 0) tr->status = ORC_OVERFLOW;
    }
    const unsigned bits = (1u << w->a) | (tr->accepted ? (1u << (8 + w->a)) : 0u);
    const unsigned before = atomic_fetch_or(&s->pstate[w->b], bits);
    const unsigned after = before | bits;
    /* resolved: the lowest accepted attempt has all its predecessors done, or every attempt is done */
    unsigned done = after & 0xffu, ok = (after >> 8) & 0xffu;
    int res_now = 0, res_before = 0;
    for (int pass = 0; pass < 2; pass++) {
        const unsigned d = pass ? (before & 0xffu) : done, o = pass ? ((before >> 8) & 0xffu) : ok;
        int r2;
        if (o) { const unsigned lower = (o & (~o + 1u)) - 1u; r2 = (d & lower) == lower; }
        else r2 = d == (1u << s->A) - 1u;
        if (pass) res_before = r2; else res_now = r2;
    }
    if (res_now && !res_before) atomic_fetch_add(&s->resolved, 1);
}

static void mark_skipped(sim_t *s, int task, int b, int a)
{
    s->res[task].ran = 0;
    wctx w = {s, task, b, a, 1, 0, 0, 0};
    /* a skipped/cancelled task only reports "done" */
    const unsigned before = atomic_fetch_or(&s->pstate[b], 1u << a);
    (void)before; (void)w;
}

static void run_owner(sim_t *s, int task)
{
    const int M = s->M, nq = ORC_D * (M - 1), n = nq + M, a = task / s->B, b = task % s->B;
    if (lower_accepted(s, b, a)) { mark_skipped(s, task, b, a); return; }
    wctx w = {s, task, b, a, 1, 0, 0, 0};
    const double *head = s->head + (size_t)b * 3 * ORC_D, *tail = s->tail + (size_t)b * 3 * ORC_D;
    const double *q = a == 0 ? s->q0 + (size_t)b * nq : s->retry_q + ((size_t)b * (s->A - 1) + (a - 1)) * nq;
    const double *ts = a == 0 ? s->ts0 + (size_t)b * M : s->retry_ts;
    double x0[ORC_MAXN];
    orc_result r;
    memcpy(x0, q, sizeof(double) * nq);
    int st = orc_T2tau(s->p, M, ts, x0 + nq);
    if (st) { finalize(s, &w, &r, 1, st); return; }
    orc_hooks hk = {&w, hk_ls_start, hk_ls_end, hk_ls_fail, hk_poll};
    orc_hooks_cur = &hk;
    st = orc_lbfgsb(s->p, s->map, M, head, tail, x0, &r);
    orc_hooks_cur = 0;
    (void)n;
    if (st == ORC_HANDED_OFF) return;                       /* the claimant owns the task now */
    if (st == ORC_STOPPED) {                                /* cancelled by an accepted earlier attempt */
        if (w.published) hk_ls_end(&w);
        mark_skipped(s, task, b, a);
        return;
    }
    finalize(s, &w, &r, 0, 0);
}

static void run_shadow(sim_t *s, int task, unsigned epoch)
{
    const int M = s->M, n = ORC_D * (M - 1) + M, a = task / s->B, b = task % s->B;
    slot_t *sl = &s->slots[task];
    orc_restart rs;
    memcpy(rs.x, sl->x, sizeof(double) * n); memcpy(rs.g, sl->g, sizeof(double) * n);
    rs.f = sl->f; memcpy(rs.costs, sl->costs, sizeof(rs.costs));
    rs.nit = sl->nit; rs.nfev = sl->nfev; rs.nfev_after = sl->nfev; rs.failed = 1;
    /* seqlock: the copy is valid only if the owner has not moved on meanwhile */
    unsigned ctl = atomic_load_explicit(&sl->ctl, memory_order_acquire);
    if (ctl != CTL(epoch, SL_CLAIMED) && ctl != CTL(epoch, SL_CONFIRMED)) { atomic_fetch_add(&s->cancelled_shadows, 1); return; }
    wctx w = {s, task, b, a, 0, epoch, 0, 0};
    if (ctl == CTL(epoch, SL_CONFIRMED)) promote(&w);
    const double *head = s->head + (size_t)b * 3 * ORC_D, *tail = s->tail + (size_t)b * 3 * ORC_D;
    orc_result r;
    orc_hooks hk = {&w, hk_ls_start, hk_ls_end, hk_ls_fail, hk_poll};
    orc_hooks_cur = &hk;
    int st = orc_lbfgsb_resume(s->p, s->map, M, head, tail, &rs, &r);
    orc_hooks_cur = 0;
    if (st == ORC_HANDED_OFF) return;
    if (st == ORC_STOPPED) {
        if (w.owner) { if (w.published) hk_ls_end(&w); mark_skipped(s, task, b, a); }
        else atomic_fetch_add(&s->cancelled_shadows, 1);
        return;
    }
    /* finished while still speculative: wait for the owner's verdict (it runs on another thread and always reaches one) */
    while (!w.owner) {
        ctl = atomic_load_explicit(&sl->ctl, memory_order_acquire);
        if (ctl == CTL(epoch, SL_CONFIRMED)) { promote(&w); break; }
        if (ctl != CTL(epoch, SL_CLAIMED)) { atomic_fetch_add(&s->cancelled_shadows, 1); return; }
        if (lower_accepted(s, b, a)) {
            /* the owner will stop at its next poll and report the task as done; nothing to do here */
            atomic_fetch_add(&s->cancelled_shadows, 1);
            return;
        }
        sched_yield();
    }
    finalize(s, &w, &r, 0, 0);
}

static void *worker(void *arg)
{
    sim_t *s = (sim_t *)arg;
    const int total = s->A * s->B;
    for (;;) {
        const int t = atomic_fetch_add(&s->queue, 1);
        if (t >= total) break;
        run_owner(s, t);
    }
    /* idle phase: serve restart requests until every problem is resolved */
    unsigned start = (unsigned)(size_t)pthread_self() * 2654435761u;
    while (atomic_load(&s->resolved) < s->B) {
        int found = 0;
        for (int k = 0; k < total && !found; k++) {
            const int t = (int)((start + (unsigned)k) % (unsigned)total);
            unsigned ctl = atomic_load_explicit(&s->slots[t].ctl, memory_order_acquire);
            if ((ctl & 3u) != SL_REQUESTED) continue;
            if (atomic_compare_exchange_strong(&s->slots[t].ctl, &ctl, (ctl & ~3u) | SL_CLAIMED)) {
                atomic_fetch_add(&s->claims, 1);
                run_shadow(s, t, ctl >> 2);
                found = 1;
            }
        }
        if (!found) sched_yield();
        start += 7919u;
    }
    return 0;
}

/* Same inputs and outputs as orc_plan_batch, computed by `threads` workers with speculative retries AND speculative
 * restarts. stats[4] = claims, hand-offs, cancelled shadows, evaluations done speculatively. */
int sim_plan_batch(const orc_params *p, const orc_map *map, int B, int M, const double *head, const double *tail,
                   const double *q0, const double *ts0, const double *retry_q, const double *retry_ts, int max_attempts,
                   int threads, double *x, double *ts, double *costs, int32_t *status, int32_t *ok, int32_t *attempt,
                   int32_t *nit, int32_t *runs, int32_t *nfev, int64_t *stats)
{
    const int nq = ORC_D * (M - 1), n = nq + M, total = max_attempts * B;
    sim_t s;
    memset(&s, 0, sizeof(s));
    s.p = p; s.map = map; s.B = B; s.M = M; s.A = max_attempts;
    s.head = head; s.tail = tail; s.q0 = q0; s.ts0 = ts0; s.retry_q = retry_q; s.retry_ts = retry_ts;
    s.slots = (slot_t *)calloc((size_t)total, sizeof(slot_t));
    s.res = (task_result *)calloc((size_t)total, sizeof(task_result));
    s.pstate = (_Atomic unsigned *)calloc((size_t)B, sizeof(unsigned));
    if (!s.slots || !s.res || !s.pstate) return -1;
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int i = 0; i < threads; i++) pthread_create(&th[i], 0, worker, &s);
    for (int i = 0; i < threads; i++) pthread_join(th[i], 0);
    /* assemble like the resolving warp: lowest accepted attempt (or the last one), counters summed over 0..that one */
    for (int b = 0; b < B; b++) {
        int acc = -1;
        for (int a = 0; a < max_attempts && acc < 0; a++) if (s.res[a * B + b].ran && s.res[a * B + b].accepted) acc = a;
        const int last = acc >= 0 ? acc : max_attempts - 1;
        int sn = 0, sr = 0, sf = 0, have = -1;
        for (int a = 0; a <= last; a++) {
            const task_result *tr = &s.res[a * B + b];
            if (!tr->ran) return -2 - b;                     /* an attempt that had to run is missing */
            sf += tr->nfev;
            if (tr->status < ORC_OVERFLOW) { sn += tr->nit; sr += 1; have = a; }
        }
        const task_result *fin = &s.res[last * B + b];
        status[b] = fin->status; ok[b] = acc >= 0; attempt[b] = last; nit[b] = sn; runs[b] = sr; nfev[b] = sf;
        memset(x + (size_t)b * n, 0, sizeof(double) * n); memset(ts + (size_t)b * M, 0, sizeof(double) * M);
        memset(costs + (size_t)b * 4, 0, sizeof(double) * 4);
        if (have >= 0) {
            const task_result *hv = &s.res[have * B + b];
            memcpy(x + (size_t)b * n, hv->x, sizeof(double) * n);
            orc_tau2T(p, M, hv->x + nq, ts + (size_t)b * M);
            memcpy(costs + (size_t)b * 4, hv->costs, sizeof(double) * 4);
        }
    }
    stats[0] = atomic_load(&s.claims); stats[1] = atomic_load(&s.handoffs); stats[2] = atomic_load(&s.cancelled_shadows);
    stats[3] = atomic_load(&s.shadow_evals);
    free(s.slots); free(s.res); free((void *)s.pstate); free(th);
    return 0;
}
