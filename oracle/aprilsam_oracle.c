/* aprilsam_oracle.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 * Plain-C restatement of the reference's batch Gauss-Newton step and chi2 for the path named by
 * BASELINE.json (april_graph_cholesky / april_graph_chi2).  Every function cites the reference
 * lines it restates (paths relative to /root/reference/).  Scalar, single-threaded, written for
 * clarity: it exists so that parity can be checked where the compiled reference (oracle/_ref) is
 * absent, and is itself PINNED against the golden vectors produced by the real reference
 * (tests/test_host_cpu.py::test_oracle_port_matches_golden; tests/golden/ made by
 * tools/make_golden.py).
 *
 * Scope of the port: factor evaluation, J'WJ assembly (+ lambda I), sparse up-looking Cholesky,
 * the two triangular solves, the node update, chi2.  The fill-reducing ordering only affects
 * speed, never the batch solution, so the port uses a plain greedy minimum degree (graphs up to
 * a few thousand poses).  NOT ported: the incremental path (aprilsam.c:377-987), whose
 * observable behaviour depends on the reference's exact ordering and tree; its parity is anchored
 * on oracle/_ref in lock-step and on the committed per-step golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may use
 * anything under oracle/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PORT_API __attribute__((visibility("default")))

/* common/math_util.h:113-122 */
static double port_mod2pi(double v)
{
    const double twopi = 6.2831853071795862319959;
    const double pi = 3.141592653589793238462643383279502884196;
    double w = v + pi;
    return (w - twopi * floor(w / twopi)) - pi;
}

/* april_graph_xyt.c:62-124 (at l_point) and :126-188 (at state): residual + Jacobians */
static void port_xyt_eval(const double *pa, const double *pb, const double *z, double Ja[9], double Jb[9], double r[3])
{
    double ca = cos(pa[2]), sa = sin(pa[2]);
    double dx = pb[0] - pa[0], dy = pb[1] - pa[1];
    double Ja_[9] = { -ca, -sa, -sa * dx + ca * dy, sa, -ca, -ca * dx - sa * dy, 0, 0, -1 };
    double Jb_[9] = { ca, sa, 0, -sa, ca, 0, 0, 0, 1 };
    memcpy(Ja, Ja_, sizeof(Ja_));
    memcpy(Jb, Jb_, sizeof(Jb_));
    r[0] = z[0] - (ca * dx + sa * dy);
    r[1] = z[1] - (-sa * dx + ca * dy);
    r[2] = port_mod2pi(z[2] - (pb[2] - pa[2]));
}

/* r' W r  (april_graph_xyt.c:110-121) */
static double port_rWr(const double *W, const double *r)
{
    double X[3];
    for (int i = 0; i < 3; i++)
        X[i] = W[3 * i] * r[0] + W[3 * i + 1] * r[1] + W[3 * i + 2] * r[2];
    return r[0] * X[0] + r[1] * X[1] + r[2] * X[2];
}

/* april_graph.c:79-98: 0.5 r'Wr for xyt factors (at state), 1.0 r'Wr for everything else */
PORT_API double oracle_chi2(int N, const double *state, int F, const int *type, const int *fa, const int *fb,
                            const double *z, const double *W)
{
    (void) N;
    double chi2 = 0;
    for (int f = 0; f < F; f++) {
        double Ja[9], Jb[9], r[3];
        if (type[f] == 1) {
            port_xyt_eval(&state[3 * fa[f]], &state[3 * fb[f]], &z[3 * f], Ja, Jb, r);
            chi2 += 0.5 * port_rWr(&W[9 * f], r);
        } else { /* april_graph_xytpos.c:63-102 */
            const double *s = &state[3 * fa[f]];
            r[0] = z[3 * f] - s[0];
            r[1] = z[3 * f + 1] - s[1];
            r[2] = port_mod2pi(z[3 * f + 2] - s[2]);
            chi2 += port_rWr(&W[9 * f], r);
        }
    }
    return chi2;
}

/* ---- sparse symmetric matrix as a list of scalar triplets, upper triangle ------------------- */
typedef struct {
    int *i, *j;
    double *v;
    int n, cap;
} trip_t;

static void trip_add(trip_t *t, int i, int j, double v)
{
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 1024;
        t->i = realloc(t->i, sizeof(int) * t->cap);
        t->j = realloc(t->j, sizeof(int) * t->cap);
        t->v = realloc(t->v, sizeof(double) * t->cap);
    }
    t->i[t->n] = i;
    t->j[t->n] = j;
    t->v[t->n] = v;
    t->n++;
}

/* C = A'B (matd_op "M'*M": common/matd.c:793-840 -> transpose + matd_multiply :230-254) */
static void port_atb(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++)
                acc += A[3 * k + i] * B[3 * k + j];
            C[3 * i + j] = acc;
        }
}

static void port_ab(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++)
                acc += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = acc;
        }
}

/* greedy minimum degree on the pose graph (stands in for aprilsam.c:999-1249; see header) */
static int *port_ordering(int N, int F, const int *type, const int *fa, const int *fb)
{
    unsigned char *adj = calloc((size_t) N * N, 1);
    int *deg = calloc(N, sizeof(int)), *order = malloc(sizeof(int) * N);
    char *gone = calloc(N, 1);
    for (int f = 0; f < F; f++)
        if (type[f] == 1 && fa[f] != fb[f] && !adj[(size_t) fa[f] * N + fb[f]]) {
            adj[(size_t) fa[f] * N + fb[f]] = adj[(size_t) fb[f] * N + fa[f]] = 1;
            deg[fa[f]]++;
            deg[fb[f]]++;
        }
    int *nb = malloc(sizeof(int) * N);
    for (int k = 0; k < N; k++) {
        int best = -1;
        for (int i = 0; i < N; i++)
            if (!gone[i] && (best < 0 || deg[i] < deg[best]))
                best = i;
        order[k] = best;
        gone[best] = 1;
        int n = 0;
        for (int i = 0; i < N; i++)
            if (adj[(size_t) best * N + i] && !gone[i])
                nb[n++] = i;
        for (int a = 0; a < n; a++) {
            adj[(size_t) nb[a] * N + best] = 0;
            deg[nb[a]]--;
            for (int b = a + 1; b < n; b++)
                if (!adj[(size_t) nb[a] * N + nb[b]]) {
                    adj[(size_t) nb[a] * N + nb[b]] = adj[(size_t) nb[b] * N + nb[a]] = 1;
                    deg[nb[a]]++;
                    deg[nb[b]]++;
                }
        }
    }
    free(adj);
    free(deg);
    free(gone);
    free(nb);
    return order;
}

typedef struct {
    int i;
    double v;
} ent_t;

static int cmp_ent(const void *a, const void *b) { return ((const ent_t *) a)->i - ((const ent_t *) b)->i; }

/* One batch Gauss-Newton step (aprilsam.c:87-375).  state[3N] in/out; lpoint_out[3N] (may be
 * NULL) receives the linearisation point.  Returns 0, or 1 if a pivot is not positive. */
PORT_API int oracle_batch_step(int N, double *state, int F, const int *type, const int *fa, const int *fb,
                               const double *z, const double *W, double lambda, double *lpoint_out)
{
    if (N <= 0 || F <= 0)
        return 0;
    /* relinearise: l_point <- state (aprilsam.c:131-135, april_graph_xyt.c:316-320) */
    double *lp = malloc(sizeof(double) * 3 * N);
    memcpy(lp, state, sizeof(double) * 3 * N);
    if (lpoint_out)
        memcpy(lpoint_out, lp, sizeof(double) * 3 * N);

    /* idxs[node] = 3 * position (aprilsam.c:141-148) */
    int *order = port_ordering(N, F, type, fa, fb);
    int *idx = malloc(sizeof(int) * N);
    for (int p = 0; p < N; p++)
        idx[order[p]] = 3 * p;
    const int n = 3 * N;

    /* assembly: upper triangle of the permuted A, and B (aprilsam.c:154-195) */
    trip_t T = { 0 };
    double *B = calloc(n, sizeof(double));
    for (int f = 0; f < F; f++) {
        double J[2][9], r[3];
        int nodes[2], nn;
        if (type[f] == 1) {
            port_xyt_eval(&lp[3 * fa[f]], &lp[3 * fb[f]], &z[3 * f], J[0], J[1], r);
            nodes[0] = fa[f];
            nodes[1] = fb[f];
            nn = 2;
        } else { /* xytpos: J = I, r = z - state (april_graph_xytpos.c:63-102) */
            memset(J[0], 0, sizeof(J[0]));
            J[0][0] = J[0][4] = J[0][8] = 1;
            const double *s = &state[3 * fa[f]];
            r[0] = z[3 * f] - s[0];
            r[1] = z[3 * f + 1] - s[1];
            r[2] = port_mod2pi(z[3 * f + 2] - s[2]);
            nodes[0] = fa[f];
            nn = 1;
        }
        for (int z0 = 0; z0 < nn; z0++) {
            double JatW[9];
            port_atb(J[z0], &W[9 * f], JatW);
            for (int z1 = 0; z1 < nn; z1++) {
                double H[9];
                port_ab(JatW, J[z1], H);
                for (int row = 0; row < 3; row++)
                    for (int col = 0; col < 3; col++) {
                        int a = row + idx[nodes[z0]], b = col + idx[nodes[z1]];
                        if (a > b)
                            continue; /* aprilsam.c:171-172 */
                        trip_add(&T, a, b, H[3 * row + col]);
                    }
            }
            for (int row = 0; row < 3; row++)
                B[idx[nodes[z0]] + row] += JatW[3 * row] * r[0] + JatW[3 * row + 1] * r[1] + JatW[3 * row + 2] * r[2];
        }
    }
    if (lambda > 0) /* aprilsam.c:197-204 */
        for (int i = 0; i < n; i++)
            trip_add(&T, i, i, lambda);

    /* upper triangle -> compressed columns with duplicates summed (cs_triplet csparse.c:1913-1932;
     * cs_symperm with no permutation keeps the upper part, :1825-1861) */
    int *cp = calloc(n + 1, sizeof(int));
    for (int k = 0; k < T.n; k++)
        cp[T.j[k] + 1]++;
    for (int j = 0; j < n; j++)
        cp[j + 1] += cp[j];
    ent_t *ce = malloc(sizeof(ent_t) * (T.n + 1));
    int *fill = malloc(sizeof(int) * n);
    memcpy(fill, cp, sizeof(int) * n);
    for (int k = 0; k < T.n; k++) {
        ent_t e = { T.i[k], T.v[k] };
        ce[fill[T.j[k]]++] = e;
    }
    int *ccnt = malloc(sizeof(int) * n);
    for (int j = 0; j < n; j++) {
        qsort(ce + cp[j], cp[j + 1] - cp[j], sizeof(ent_t), cmp_ent);
        int w = cp[j];
        for (int k = cp[j]; k < cp[j + 1]; k++) {
            if (w > cp[j] && ce[w - 1].i == ce[k].i)
                ce[w - 1].v += ce[k].v;
            else
                ce[w++] = ce[k];
        }
        ccnt[j] = w - cp[j];
    }

    /* elimination tree (cs_etree, csparse.c:906-933) */
    int *parent = malloc(sizeof(int) * n), *anc = malloc(sizeof(int) * n);
    for (int k = 0; k < n; k++) {
        parent[k] = anc[k] = -1;
        for (int p = cp[k]; p < cp[k] + ccnt[k]; p++) {
            int i = ce[p].i;
            while (i != -1 && i < k) {
                int inext = anc[i];
                anc[i] = k;
                if (inext == -1)
                    parent[i] = k;
                i = inext;
            }
        }
    }

    /* up-looking numeric Cholesky, L stored by columns (cs_chol csparse.c:462-513 with the
     * row patterns found by cs_ereach :441-459); columns grow dynamically instead of cs_counts */
    int **Li = calloc(n, sizeof(int *)), *Ln = calloc(n, sizeof(int)), *Lc = calloc(n, sizeof(int));
    double **Lx = calloc(n, sizeof(double *));
    double *x = calloc(n, sizeof(double));
    int *mark = malloc(sizeof(int) * n), *stack = malloc(sizeof(int) * n), *pat = malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++)
        mark[i] = -1;
    int rc = 0;
    for (int k = 0; k < n && !rc; k++) {
        int top = n;
        mark[k] = k;
        double d = 0;
        for (int p = cp[k]; p < cp[k] + ccnt[k]; p++) {
            int i = ce[p].i;
            if (i == k) {
                d = ce[p].v;
                continue;
            }
            x[i] = ce[p].v;
            int len = 0;
            for (; mark[i] != k; i = parent[i]) { /* walk up the etree */
                stack[len++] = i;
                mark[i] = k;
            }
            while (len > 0)
                pat[--top] = stack[--len];
        }
        for (; top < n; top++) { /* pattern in topological order */
            int i = pat[top];
            double lki = x[i] / Lx[i][0]; /* L(k,i) = x(i) / L(i,i) */
            x[i] = 0;
            for (int p = 1; p < Ln[i]; p++)
                x[Li[i][p]] -= Lx[i][p] * lki;
            d -= lki * lki;
            if (Ln[i] == Lc[i]) {
                Lc[i] = Lc[i] ? 2 * Lc[i] : 8;
                Li[i] = realloc(Li[i], sizeof(int) * Lc[i]);
                Lx[i] = realloc(Lx[i], sizeof(double) * Lc[i]);
            }
            Li[i][Ln[i]] = k;
            Lx[i][Ln[i]] = lki;
            Ln[i]++;
        }
        if (!(d > 0)) { /* csparse.c:505-506 */
            rc = 1;
            break;
        }
        Lc[k] = 8;
        Li[k] = malloc(sizeof(int) * 8);
        Lx[k] = malloc(sizeof(double) * 8);
        Li[k][0] = k;
        Lx[k][0] = sqrt(d);
        Ln[k] = 1;
    }

    if (!rc) {
        /* U'y = B then U x = y with U = L' (smatd_chol_solve_full, common/smatd.c:1100-1114,
         * :1051-1073 column scatter, :1075-1097 row dot) */
        double *y = B;
        for (int j = 0; j < n; j++) {
            y[j] /= Lx[j][0];
            for (int p = 1; p < Ln[j]; p++)
                y[Li[j][p]] -= Lx[j][p] * y[j];
        }
        for (int j = n - 1; j >= 0; j--) {
            double acc = y[j];
            for (int p = 1; p < Ln[j]; p++)
                acc -= Lx[j][p] * y[Li[j][p]];
            y[j] = acc / Lx[j][0];
        }
        /* state = l_point + dx, theta wrapped, skipped on NaN (april_graph_xyt.c:302-314) */
        for (int i = 0; i < N; i++) {
            const double *dx = &y[idx[i]];
            if (isnan(dx[0]) || isnan(dx[1]) || isnan(dx[2]))
                continue;
            for (int k = 0; k < 3; k++)
                state[3 * i + k] = lp[3 * i + k] + dx[k];
            state[3 * i + 2] = port_mod2pi(state[3 * i + 2]);
        }
    }
    for (int i = 0; i < n; i++) {
        free(Li[i]);
        free(Lx[i]);
    }
    free(Li); free(Lx); free(Ln); free(Lc); free(x); free(mark); free(stack); free(pat);
    free(parent); free(anc); free(cp); free(ce); free(fill); free(ccnt);
    free(T.i); free(T.j); free(T.v); free(B); free(idx); free(order); free(lp);
    return rc;
}

PORT_API int aprilsam_oracle_port_available(void) { return 1; }
