"""numpy emulation of the CUDA kernels, driven by the SAME plan data that is uploaded to HBM.

TEST INFRASTRUCTURE.  Used by the CPU tests to validate the host symbolic layer (row lists,
relative indices, gather lists, task order) without a GPU, and as an executable
specification of k_linearize / k_factor / k_backsolve index arithmetic.  Not shipped, not a
fallback: nothing under aprilsam_b200/ imports this.
"""
from __future__ import annotations

import numpy as np

TR_FLAG = 1 << 30
TWOPI = 6.2831853071795862319959


def mod2pi(v):
    w = v + np.pi
    return (w - TWOPI * np.floor(w / TWOPI)) - np.pi


def xyt_eval(pa, pb, z):
    ca, sa = np.cos(pa[2]), np.sin(pa[2])
    dx, dy = pb[0] - pa[0], pb[1] - pa[1]
    Ja = np.array([[-ca, -sa, -sa * dx + ca * dy], [sa, -ca, -ca * dx - sa * dy], [0, 0, -1.0]])
    Jb = np.array([[ca, sa, 0], [-sa, ca, 0], [0, 0, 1.0]])
    r = np.array([z[0] - (ca * dx + sa * dy), z[1] - (-sa * dx + ca * dy), mod2pi(z[2] - (pb[2] - pa[2]))])
    return Ja, Jb, r


class Hessian:
    """Adiag / Aoff / Bq exactly as k_linearize leaves them in HBM (node-id space)."""

    def __init__(self, n_nodes, n_slots):
        self.Adiag = np.zeros((n_nodes, 3, 3))
        self.Aoff = np.zeros((n_slots, 3, 3))
        self.B = np.zeros((n_nodes, 3))

    def grow(self, n_nodes, n_slots):
        if n_nodes > len(self.Adiag):
            self.Adiag = np.concatenate([self.Adiag, np.zeros((n_nodes - len(self.Adiag), 3, 3))])
            self.B = np.concatenate([self.B, np.zeros((n_nodes - len(self.B), 3))])
        if n_slots > len(self.Aoff):
            self.Aoff = np.concatenate([self.Aoff, np.zeros((n_slots - len(self.Aoff), 3, 3))])

    def reset(self, n_lambda, lam):
        self.Adiag[:] = 0
        self.Aoff[:] = 0
        self.B[:] = 0
        for k in range(3):
            self.Adiag[:n_lambda, k, k] = lam

    def linearize(self, f_range, ftype, fa, fb, fz, fW, lp, st, node2q, fslot, pts=None):
        for k, f in enumerate(f_range):
            W = fW[f].reshape(3, 3)
            a = fa[f]
            if ftype[f] == 2:
                p = st[a] if pts is None else pts[k, :3]
                r = np.array([fz[f, 0] - p[0], fz[f, 1] - p[1], mod2pi(fz[f, 2] - p[2])])
                self.Adiag[a] += np.triu(W)
                self.B[a] += W @ r
                continue
            b = fb[f]
            pa, pb = (lp[a], lp[b]) if pts is None else (pts[k, :3], pts[k, 3:])
            Ja, Jb, r = xyt_eval(pa, pb, fz[f])
            JatW, JbtW = Ja.T @ W, Jb.T @ W
            self.Adiag[a] += np.triu(JatW @ Ja)
            self.Adiag[b] += np.triu(JbtW @ Jb)
            if node2q[a] < node2q[b]:
                H, early = JatW @ Jb, a
            else:
                H, early = JbtW @ Ja, b
            self.Aoff[fslot[f]] += H if early == min(a, b) else H.T
            self.B[a] += JatW @ r
            self.B[b] += JbtW @ r

    def dense(self, pairs):
        """Full symmetric matrix in node-id order. pairs[slot] = (lo, hi)."""
        import scipy.sparse as sp
        n = len(self.Adiag)
        rows, cols, vals = [], [], []
        for i in range(n):
            D = np.triu(self.Adiag[i]) + np.triu(self.Adiag[i], 1).T
            for p in range(3):
                for q in range(3):
                    rows.append(3 * i + p); cols.append(3 * i + q); vals.append(D[p, q])
        for s, (lo, hi) in enumerate(pairs):
            S = self.Aoff[s]
            for p in range(3):
                for q in range(3):
                    rows.append(3 * lo + p); cols.append(3 * hi + q); vals.append(S[p, q])
                    rows.append(3 * hi + q); cols.append(3 * lo + p); vals.append(S[p, q])
        return sp.csc_matrix((vals, (rows, cols)), shape=(3 * n, 3 * n))


class Fronts:
    """Device arena emulation: one (m x m) column-major front + m rhs doubles per supernode."""

    def __init__(self):
        self.F = {}    # f_off -> (m x m) array (we index [row, col])
        self.rhs = {}  # f_off -> m
        self.y = None
        self.x = None

    def ensure(self, N):
        if self.y is None or len(self.y) < 3 * N:
            ny = np.zeros(3 * N)
            nx = np.zeros(3 * N)
            if self.y is not None:
                ny[:len(self.y)] = self.y
                nx[:len(self.x)] = self.x
            self.y, self.x = ny, nx


def seg_views(desc, ipool, s):
    mb, cb, ch, ac, seg = (int(desc[k][s]) for k in ("mb", "cb", "ch_cnt", "a_cnt", "seg"))
    o = seg
    rows = ipool[o:o + mb]; o += mb
    rel = ipool[o:o + mb]; o += mb
    children = ipool[o:o + ch]; o += ch
    a_slot = ipool[o:o + ac]; o += ac
    a_rb = ipool[o:o + ac]; o += ac
    a_cb = ipool[o:o + ac]
    return rows, rel, children, a_slot, a_rb, a_cb


def factor(fr: Fronts, H: Hessian, desc, ipool, q2node, tasks, nwait=None, check_order=True, prior=(), keep=None,
           count_prior=True, panel=0, syrk=None):
    """k_factor: assemble + eliminate the listed supernodes (children first).  `prior` = supernodes
    factored by an earlier launch of the same solve; count_prior: their arrivals are part of nwait
    (k_factor_leaf before k_factor) or not (arrival counters zeroed in between: the multi-GPU top)."""
    done = set()
    intask = set(int(t) for t in tasks) | (set(int(t) for t in prior) if count_prior else set())
    done |= set(int(t) for t in prior)
    for ti, s in enumerate(tasks):
        s = int(s)
        if s in done:  # further workers of a multi-CTA team: same supernode
            assert (int(nwait[ti]) >> 24) & 0x7f > 1
            continue
        rows, rel, children, a_slot, a_rb, a_cb = seg_views(desc, ipool, s)
        mb, cb, first = int(desc["mb"][s]), int(desc["cb"][s]), int(desc["first"][s])
        m, c = 3 * mb, 3 * cb
        assert list(rows[:cb]) == list(range(first, first + cb)), "own columns first"
        assert np.all(np.diff(rows) > 0), "rows ascending"
        F = np.zeros((m, m))
        rhs = np.zeros(m)
        for k in range(cb):
            node = q2node[first + k]
            D = H.Adiag[node]
            for p in range(3):
                for q in range(p + 1):
                    F[3 * k + p, 3 * k + q] = D[q, p]
            rhs[3 * k:3 * k + 3] = H.B[node]
        for i in range(len(a_slot)):
            rb = int(a_rb[i]) & ~TR_FLAG
            S = H.Aoff[a_slot[i]]
            blk = S if (int(a_rb[i]) & TR_FLAG) else S.T  # F[row p, col q] = S[p,q] if TR else S[q,p]
            assert rb >= int(a_cb[i]) and rb != int(a_cb[i]), "gather target strictly below the diagonal block"
            F[3 * rb:3 * rb + 3, 3 * int(a_cb[i]):3 * int(a_cb[i]) + 3] = blk
        nw = 0
        for cs in children:
            cs = int(cs)
            if cs in intask:
                nw += 1
                if check_order:
                    assert cs in done, f"child {cs} of {s} scheduled after its parent"
            assert int(desc["parent"][cs]) == s
            crows, crel, *_ = seg_views(desc, ipool, cs)
            cmb, ccb = int(desc["mb"][cs]), int(desc["cb"][cs])
            CF = fr.F[int(desc["f_off"][cs])]
            crhs = fr.rhs[int(desc["f_off"][cs])]
            idx = np.concatenate([3 * int(crel[k]) + np.arange(3) for k in range(ccb, cmb)]) if cmb > ccb else np.zeros(0, int)
            # parent's row list must hold the same positions
            assert all(rows[int(crel[k])] == crows[k] for k in range(ccb, cmb)), "rel index mismatch"
            U = CF[3 * ccb:, 3 * ccb:]
            F[np.ix_(idx, idx)] += np.tril(U)
            rhs[idx] += crhs[3 * ccb:]
        if nwait is not None:
            assert nw == (int(nwait[ti]) & 0xffff), f"nwait mismatch for supernode {s}: {nw} vs {nwait[ti]}"
        # partial Cholesky of the first c columns (right-looking, lower triangle only)
        if panel > 0:
            # blocked like the kernels: a panel of `panel` columns in double precision, then ONE product for the
            # trailing matrix, computed by syrk(P) ~ P P' (tools/ozaki_study.py plugs reduced-precision products in)
            for k0 in range(0, c, panel):
                k1 = min(c, k0 + panel)
                for k in range(k0, k1):
                    d = F[k, k]
                    if not d > 0:
                        raise np.linalg.LinAlgError(f"pivot <= 0 in supernode {s}")
                    piv = np.sqrt(d)
                    F[k, k] = piv
                    F[k + 1:, k] /= piv
                    rhs[k] /= piv
                    lk = F[k + 1:, k]
                    F[k + 1:, k + 1:k1] -= np.outer(lk, lk[:k1 - k - 1])  # the panel's own columns only
                    rhs[k + 1:] -= lk * rhs[k]
                if k1 < m:
                    P = F[k1:, k0:k1]
                    F[k1:, k1:] -= np.tril(syrk(P) if syrk is not None else P @ P.T)
                F[:, :] = np.tril(F)
        for k in range(c if panel <= 0 else 0):
            d = F[k, k]
            if not d > 0:
                raise np.linalg.LinAlgError(f"pivot <= 0 in supernode {s}")
            piv = np.sqrt(d)
            F[k, k] = piv
            F[k + 1:, k] /= piv
            rhs[k] /= piv
            lk = F[k + 1:, k]
            F[k + 1:, k + 1:] -= np.tril(np.outer(lk, lk))
            rhs[k + 1:] -= lk * rhs[k]
        off = int(desc["f_off"][s])
        if keep is not None and int(keep[ti]) != 0:
            # partial re-factorisation (cta_front, keepw): the host claims that the first kb poses' columns of the
            # retained front are unchanged by this step -- check it against the full elimination just done
            kb, omb = int(keep[ti]) >> 16, int(keep[ti]) & 0xffff
            kc, mo = 3 * kb, 3 * omb
            assert off in fr.F, "kept columns need the retained front at the same offset"
            Fo, ro = fr.F[off], fr.rhs[off]
            assert Fo.shape[0] == mo and 0 < kc < c and mo <= m
            scale = max(1.0, np.abs(Fo).max())
            assert np.abs(np.tril(F[:mo, :kc]) - np.tril(Fo[:, :kc])).max() < 1e-9 * scale, f"supernode {s}: kept L changed"
            assert np.abs(F[mo:, :kc]).max(initial=0.0) == 0.0, f"supernode {s}: appended rows are not zero in kept columns"
            assert np.abs(rhs[:kc] - ro[:kc]).max() < 1e-9 * max(1.0, np.abs(ro).max()), f"supernode {s}: kept y changed"
            fr.kept_cols = getattr(fr, "kept_cols", 0) + kc
        fr.F[off] = F
        fr.rhs[off] = rhs
        fr.y[3 * first:3 * first + c] = rhs[:c]
        done.add(s)


def backsolve(fr: Fronts, desc, ipool, btasks):
    """k_backsolve: parents first; list must be closed under ancestors.  An entry is a supernode id, or -- wide
    supernodes of a batch schedule -- supernode | (block + 1) << 24: one 96-column block per entry, the blocks of
    a supernode consecutive, last block first (cta_backsolve, blk_only)."""
    done = set()
    blocks_seen = {}
    for e in btasks:
        e = int(e)
        s, blk = e & 0xffffff, ((e >> 24) & 0x7f) - 1
        P = int(desc["parent"][s])
        assert P < 0 or P in done, f"parent {P} of {s} not solved first"
        rows, *_ = seg_views(desc, ipool, s)
        mb, cb, first = int(desc["mb"][s]), int(desc["cb"][s]), int(desc["first"][s])
        c = 3 * cb
        L = fr.F[int(desc["f_off"][s])]
        xs = np.concatenate([fr.x[3 * int(r):3 * int(r) + 3] for r in rows[cb:]]) if mb > cb else np.zeros(0)
        if blk < 0:
            w = fr.y[3 * first:3 * first + c] - L[c:, :c].T @ xs
            L11 = np.tril(L[:c, :c])
            fr.x[3 * first:3 * first + c] = np.linalg.solve(L11.T, w)
            done.add(s)
            continue
        nblk = (c + 95) // 96
        assert c > 96 and 0 <= blk < nblk
        assert blocks_seen.get(s, nblk) == blk + 1, f"blocks of supernode {s} must come last block first, consecutively"
        blocks_seen[s] = blk
        b0, be = 96 * blk, min(96 * blk + 96, c)
        xlater = fr.x[3 * first + be:3 * first + c]  # the later blocks of this supernode: solved by earlier entries
        w = fr.y[3 * first + b0:3 * first + be] - L[c:, b0:be].T @ xs - L[be:c, b0:be].T @ xlater
        fr.x[3 * first + b0:3 * first + be] = np.linalg.solve(np.tril(L[b0:be, b0:be]).T, w)
        if blk == 0:
            done.add(s)
