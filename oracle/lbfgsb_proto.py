"""lbfgsb_proto.py -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).

Pure-Python restatement (no LAPACK/BLAS calls) of scipy 1.18.1's L-BFGS-B, unbounded case, as built against OpenBLAS
0.3.31 (SkylakeX kernels): the readable twin of the L-BFGS-B section of oracle/minco_oracle.c. It was derived call by
call against scipy.optimize._lbfgsb.setulb's workspace (ws, wy, sy, ss, wn, wn1, z, d) and ends in the same x bit for
bit (300/300 runs at M = 3, 60/60 at M = 10 when driven by the reference's own get_cost/get_grad). Kept for study;
the tests use the C version."""
import math, ctypes
import numpy as np
_libm = ctypes.CDLL('libm.so.6'); _libm.fma.restype = ctypes.c_double; _libm.fma.argtypes = [ctypes.c_double] * 3
def fma(a, b, c): return _libm.fma(float(a), float(b), float(c))
EPS = np.finfo(float).eps

def dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx: gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx): stpf = stpc
        else: stpf = stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx: gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if abs(stpc - stp) > abs(stpq - stp): stpf = stpc
        else: stpf = stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx: gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0 and gamma != 0: stpc = stp + r * (stx - stp)
        elif stp > stx: stpc = stpmax
        else: stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            if abs(stpc - stp) < abs(stpq - stp): stpf = stpc
            else: stpf = stpq
            if stp > stx: stpf = min(stp + 0.66 * (sty - stp), stpf)
            else: stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            if abs(stpc - stp) > abs(stpq - stp): stpf = stpc
            else: stpf = stpq
            stpf = min(stpmax, stpf); stpf = max(stpmin, stpf)
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * math.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty: gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpc = stp + r * (sty - stp)
            stpf = stpc
        elif stp > stx: stpf = stpmax
        else: stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0: sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt

class Dcsrch:
    def __init__(self, ftol=1e-3, gtol=0.9, xtol=0.1, stpmin=0.0, stpmax=1e10):
        self.ftol, self.gtol, self.xtol, self.stpmin, self.stpmax = ftol, gtol, xtol, stpmin, stpmax
    def start(self, stp, f, g):
        self.brackt = False; self.stage = 1
        self.finit = f; self.ginit = g; self.gtest = self.ftol * g
        self.width = self.stpmax - self.stpmin; self.width1 = self.width / 0.5
        self.stx = 0.0; self.fx = f; self.gx = g
        self.sty = 0.0; self.fy = f; self.gy = g
        self.stmin = 0.0; self.stmax = stp + 4.0 * stp
        return stp, 'FG'
    def step(self, stp, f, g):
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and g >= 0: self.stage = 2
        task = 'FG'
        if self.brackt and (stp <= self.stmin or stp >= self.stmax): task = 'WARN_ROUND'
        if self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax: task = 'WARN_XTOL'
        if stp == self.stpmax and f <= ftest and g <= self.gtest: task = 'WARN_STPMAX'
        if stp == self.stpmin and (f > ftest or g >= self.gtest): task = 'WARN_STPMIN'
        if f <= ftest and abs(g) <= self.gtol * -self.ginit: task = 'CONV'
        if task != 'FG': return stp, task
        if self.stage == 1 and f <= self.fx and f > ftest:
            fm = f - stp * self.gtest
            fxm = self.fx - self.stx * self.gtest
            fym = self.fy - self.sty * self.gtest
            gm = g - self.gtest; gxm = self.gx - self.gtest; gym = self.gy - self.gtest
            self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt = dcstep(self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm, self.brackt, self.stmin, self.stmax)
            self.fx = fxm + self.stx * self.gtest; self.fy = fym + self.sty * self.gtest
            self.gx = gxm + self.gtest; self.gy = gym + self.gtest
        else:
            self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, self.brackt = dcstep(self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, g, self.brackt, self.stmin, self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= 0.66 * self.width1:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.width1 = self.width; self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin = min(self.stx, self.sty); self.stmax = max(self.stx, self.sty)
        else:
            self.stmin = stp + 1.1 * (stp - self.stx); self.stmax = stp + 4.0 * (stp - self.stx)
        stp = max(stp, self.stpmin); stp = min(stp, self.stpmax)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or (self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax):
            stp = self.stx
        return stp, 'FG'


def dot_loop(a, b):
    s = 0.0
    for u, v in zip(a, b): s = s + u * v
    return s

def ddot(a, b):
    n = len(a); n1 = n & -16; assert n < 32
    s = 0.0
    if n1:
        p = [a[i] * b[i] for i in range(16)]
        v = [((p[l] + p[4 + l]) + p[8 + l]) + p[12 + l] for l in range(4)]
        s = (v[0] + v[2]) + (v[1] + v[3])
    for i in range(n1, n): s = fma(a[i], b[i], s)
    return s

def dnrm2(a):
    """OpenBLAS nrm2.S (x87, 80-bit extended = np.longdouble on x86-64): four accumulators -- element i of the leading
    blocks of 8 goes to accumulator i mod 4, the remaining n mod 8 elements to accumulator 0 -- D + ((C + A) + B), fsqrt."""
    acc = [np.longdouble(0)] * 4
    n8 = len(a) & ~7
    for i, v in enumerate(a):
        v = np.longdouble(v)
        k = i & 3 if i < n8 else 0
        acc[k] = acc[k] + v * v
    return float(np.sqrt(acc[3] + ((acc[2] + acc[0]) + acc[1])))

def potf2(a, o, n):
    for j in range(n):
        ajj = a[o+j, o+j] - ddot(a[o:o+j, o+j], a[o:o+j, o+j])
        if ajj <= 0.0: return False
        ajj = math.sqrt(ajj); a[o+j, o+j] = ajj
        if j == n - 1: break
        m1 = j & -4
        for i in range(j + 1, n):
            col = a[o:o+j, o+i]; x = a[o:o+j, o+j]; v = a[o+j, o+i]
            if m1:
                acc = [col[l] * x[l] for l in range(4)]
                if m1 == 8: acc = [acc[l] + col[4+l] * x[4+l] for l in range(4)]
                v = v - ((acc[0] + acc[2]) + (acc[1] + acc[3]))
            ks = range(m1, j); k3 = j - m1
            if k3 == 1: v = fma(col[m1], -x[m1], v)
            elif k3 == 2: v = v + fma(col[m1], -x[m1], col[m1+1] * -x[m1+1])
            elif k3 == 3: v = v + fma(col[m1+2], -x[m1+2], fma(col[m1], -x[m1], col[m1+1] * -x[m1+1]))
            a[o+j, o+i] = v
        r = 1.0 / ajj
        for i in range(j + 1, n): a[o+j, o+i] = a[o+j, o+i] * r
    return True

def trsv_T(u, o, n, b):
    for i in range(n):
        if i > 0: b[i] = b[i] - ddot(u[o:o+i, o+i], b[0:i])
        b[i] = b[i] / u[o+i, o+i]

def trsv_N(u, o, n, b):
    for i in range(n - 1, -1, -1):
        b[i] = b[i] / u[o+i, o+i]
        t = -b[i]
        for k in range(i): b[k] = fma(t, u[o+k, o+i], b[k])

def blocks_desc(m):
    out = []; kk = 0; i = 8
    while i > 0:
        if m & i: out.append((kk, i)); kk += i
        i >>= 1
    return out

def trsm_LT(u, n, B):
    """u[0:n,0:n] upper; solves u^T X = B in place, B: (n, nrhs) view, nrhs >= 2, n <= 15"""
    for c in range(B.shape[1]):
        b = B[:, c]
        for kk, bs in blocks_desc(n):
            if kk > 0:
                for i in range(kk, kk + bs):
                    acc = 0.0
                    for k in range(kk): acc = fma(u[k, i], b[k], acc)
                    b[i] = b[i] - acc
            for i in range(kk, kk + bs):
                b[i] = b[i] * (1.0 / u[i, i])
                for k in range(i + 1, kk + bs): b[k] = fma(-b[i], u[i, k], b[k])

class Mem:
    def __init__(self, n, m): self.n, self.m = n, m; self.reset()
    def reset(self):
        m = self.m
        self.col = 0; self.theta = 1.0; self.iupdat = 0
        self.ws = []; self.wy = []
        self.sy = np.zeros((m, m)); self.yy = np.zeros((m, m)); self.rz = np.zeros((m, m))
    def update(self, s, y, rr, dr):
        m = self.m
        self.iupdat += 1
        if self.iupdat <= m: self.col = self.iupdat
        else: self.ws.pop(0); self.wy.pop(0)
        self.ws.append(s.copy()); self.wy.append(y.copy())
        self.theta = rr / dr
        col = self.col
        if self.iupdat > m:
            for j in range(m - 1):
                self.sy[j:m-1, j] = self.sy[j+1:m, j+1]       # only the diagonal is used by the unbounded path
                self.yy[j:m-1, j] = self.yy[j+1:m, j+1]
                self.rz[0:m-1, j] = self.rz[1:m, j+1]
        self.sy[col-1, col-1] = dr
        for jy in range(col): self.yy[col-1, jy] = dot_loop(self.wy[col-1], self.wy[jy])
        for i in range(col): self.rz[i, col-1] = dot_loop(self.ws[i], self.wy[col-1])
    def factor(self):
        col, theta = self.col, self.theta
        wn = np.zeros((2*col, 2*col))
        for iy in range(col):
            is_ = col + iy
            for jy in range(iy + 1):
                wn[jy, iy] = self.yy[iy, jy] / theta
                wn[col + jy, is_] = 0.0 * theta
            for jy in range(iy): wn[jy, is_] = -0.0
            for jy in range(iy, col): wn[jy, is_] = self.rz[iy, jy]
            wn[iy, iy] = wn[iy, iy] + self.sy[iy, iy]
        if not potf2(wn, 0, col): return None
        if col == 1: trsv_T(wn, 0, 1, wn[0:1, 1])
        else: trsm_LT(wn, col, wn[0:col, col:2*col])
        for is_ in range(col, 2*col):
            for js in range(is_, 2*col):
                wn[is_, js] = wn[is_, js] + ddot(wn[0:col, is_], wn[0:col, js])
        if not potf2(wn, col, col): return None
        return wn
    def step(self, g, wn):
        col, theta, n = self.col, self.theta, self.n
        d = -g
        wv = np.zeros(2*col)
        for i in range(col):
            wv[i] = dot_loop(self.wy[i], d); wv[col + i] = theta * dot_loop(self.ws[i], d)
        trsv_T(wn, 0, 2*col, wv)
        for i in range(col): wv[i] = -wv[i]
        trsv_N(wn, 0, 2*col, wv)
        for jy in range(col):
            for i in range(n):
                d[i] = d[i] + self.wy[jy][i] * wv[jy] / theta + self.ws[jy][i] * wv[col + jy]
        return d * (1.0 / theta)

def lbfgsb(fg, x0, m=10, ftol=1e-4, pgtol=1e-4, maxls=20):
    n = len(x0); tol = (ftol / EPS) * EPS
    x = np.array(x0, dtype=float)
    f, g = fg(x); g = np.array(g, dtype=float)
    nfev = 1; it = 0; xlast = x.copy()
    C = Mem(n, m)
    if np.max(np.abs(g)) <= pgtol: return dict(x=x, f=f, nit=0, nfev=nfev, status='CONV_PG')
    while True:
        if C.col == 0: z = x - g
        else:
            wn = C.factor()
            if wn is None: C.reset(); continue
            z = x + C.step(g, wn)
        d = z - x
        dnorm = dnrm2(d)
        stp = min(1.0 / dnorm, 1e10) if it == 0 else 1.0
        t = x.copy(); r = g.copy(); fold = f
        gd = ddot(g, d); gdold = gd
        fail = gd >= 0
        if not fail:
            ls = Dcsrch(); stp, task = ls.start(stp, f, gd); ifun = 0
            while True:
                ifun += 1; nfev += 1
                if ifun - 1 >= maxls: fail = True; nfev -= 1; break
                xn = z.copy() if stp == 1.0 else stp * d + t
                if np.array_equal(xn, xlast): nfev -= 1
                x = xn
                f, g = fg(x); g = np.array(g, dtype=float); xlast = x.copy()
                gd = ddot(g, d)
                stp, task = ls.step(stp, f, gd)
                if task != 'FG': break
        if fail:
            x = t; g = r; f = fold
            if C.col == 0: return dict(x=x, f=f, nit=it, nfev=nfev, status='ABNORMAL')
            C.reset(); continue
        it += 1
        if np.max(np.abs(g)) <= pgtol: return dict(x=x, f=f, nit=it, nfev=nfev, status='CONV_PG')
        if fold - f <= tol * max(abs(fold), abs(f), 1.0): return dict(x=x, f=f, nit=it, nfev=nfev, status='CONV_F')
        yv = g - r; rr = dnrm2(yv) ** 2
        if stp == 1.0: dr = gd - gdold; ddum = -gdold; sv = d.copy()
        else: dr = (gd - gdold) * stp; sv = d * stp; ddum = -gdold * stp
        if dr <= EPS * ddum: continue
        C.update(sv, yv, rr, dr)
