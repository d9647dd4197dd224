// asam_cuda.cu -- sm_100a CUDA kernels + C-ABI for the AprilSAM Gauss-Newton path.
//
// Kernels (see DESIGN.md for the roofline of each):
//   k_linearize   one thread per factor: residual, Jacobians, J'WJ / J'Wr, atomically
//                 scattered into the block Hessian (Adiag / Aoff / Bq).
//                 reference: april_graph_xyt.c:62-124, april_graph_xytpos.c:63-102,
//                            aprilsam.c:154-195 (batch), :508-542 (incremental)
//   k_factor      persistent, dependency-driven multifrontal supernodal Cholesky with the
//                 forward solve fused in (rhs carried as an extra front column).
//                 reference: csparse.c:462-513 (cs_chol), smatd.c:1051-1073, and for a
//                 subset of supernodes aprilsam.c:791-906 (reconstruct + re-eliminate)
//   k_backsolve   persistent, dependency-driven back-substitution L' x = y.
//                 reference: smatd.c:1075-1097, aprilsam.c:721-779
//   k_chi2*       deterministic reduction of the factor energies at `state`.
//                 reference: april_graph.c:79-98, april_graph_xyt.c:126-188
//
// There is no CPU implementation of any of this in the product: if the CUDA runtime or a
// device is missing, asam_dev_create() fails and the host API aborts.

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "asam_cuda.h"

#define ASAM_EXPORT extern "C" __attribute__((visibility("default")))
#define ASAM_TR_FLAG (1 << 30)

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return set_err("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

ASAM_EXPORT const char *asam_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0; // bytes
};

struct asam_dev {
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;

    // graph mirror
    Buf f_type, f_a, f_b, f_z, f_W, f_slot;
    Buf lp, st, node2q, q2node;
    // hessian
    Buf Adiag, Aoff, Bq, y, x;
    // plan
    Buf sn, ipool, arena;
    Buf arrive, xdone;
    // task lists
    Buf tasks_full, nwait_full, btasks_full;
    int ntasks_full = 0;
    int bt_start = 0, bt_count = 0, bt_cap = 0; // btasks_full holds [bt_start, bt_start+bt_count)
    Buf tasks_tmp, nwait_tmp, btasks_tmp;
    // misc
    Buf ctrl;     // int[8]: [0] ticket, [1] err, [2] ticket backsolve
    Buf partial;  // chi2 partial sums
    Buf patch_ids, patch_desc;
    Buf pts;
    int epoch = 0;

    // pinned staging
    char *pin = nullptr;
    size_t pin_cap = 0, pin_off = 0;

    // launch config
    int fac_threads = 256, fac_grid = 0, fac_smem = 0;
    int bs_threads = 128, bs_grid = 0, bs_smem = 0;

    int64_t n_launch = 0, n_h2d = 0, n_d2h = 0;

    int timing = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ev_set[3] = {0, 0, 0};
    cudaEvent_t tev[2] = {nullptr, nullptr};
    Buf flush;
    int flush_val = 0;
};

static int buf_reserve(asam_dev *d, Buf &b, size_t bytes, bool keep, bool zero_new)
{
    if (bytes <= b.cap)
        return 0;
    size_t want = b.cap ? b.cap : 4096;
    while (want < bytes)
        want = want + want / 2 + 4096;
    void *np = nullptr;
    CK(cudaMalloc(&np, want));
    if (zero_new)
        CK(cudaMemsetAsync(np, 0, want, d->stream));
    if (keep && b.p && b.cap)
        CK(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, d->stream));
    if (b.p) {
        CK(cudaStreamSynchronize(d->stream));
        CK(cudaFree(b.p));
    }
    b.p = np;
    b.cap = want;
    return 0;
}

static int upload(asam_dev *d, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return 0;
    d->n_h2d += (int64_t) bytes;
    if (bytes > d->pin_cap) { // too large to stage: synchronous pageable copy
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        d->pin_off = 0;
        return 0;
    }
    if (d->pin_off + bytes > d->pin_cap) {
        CK(cudaStreamSynchronize(d->stream));
        d->pin_off = 0;
    }
    memcpy(d->pin + d->pin_off, src, bytes);
    CK(cudaMemcpyAsync(dst, d->pin + d->pin_off, bytes, cudaMemcpyHostToDevice, d->stream));
    d->pin_off += (bytes + 255) & ~(size_t) 255;
    return 0;
}

static int download(asam_dev *d, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return 0;
    d->n_d2h += (int64_t) bytes;
    if (bytes <= d->pin_cap) {
        CK(cudaStreamSynchronize(d->stream)); // staging area is free after this
        d->pin_off = 0;
        CK(cudaMemcpyAsync(d->pin, src, bytes, cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        memcpy(dst, d->pin, bytes);
    } else {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        d->pin_off = 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double d_mod2pi(double v)
{
    // reference: common/math_util.h:113-122 (same constants, same operation order)
    const double twopi = 6.2831853071795862319959;
    const double pi = 3.141592653589793238462643383279502884196;
    double w = v + pi;
    return (w - twopi * floor(w / twopi)) - pi;
}

// C = A' * B for row-major 3x3 (matd_op("M'*M"): transpose then naive triple loop,
// reference common/matd.c:230-254)
__device__ __forceinline__ void d_atb(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++)
                acc += A[k * 3 + i] * B[k * 3 + j];
            C[i * 3 + j] = acc;
        }
}

__device__ __forceinline__ void d_ab(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++)
                acc += A[i * 3 + k] * B[k * 3 + j];
            C[i * 3 + j] = acc;
        }
}

__device__ __forceinline__ void d_av(const double *A, const double *v, double *r)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
        r[i] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// Residual + Jacobians of an xyt factor at (pa, pb)   (april_graph_xyt.c:62-124)
__device__ __forceinline__ void d_xyt_eval(const double *pa, const double *pb, const double *z, double *Ja,
                                           double *Jb, double *r)
{
    double ca, sa;
    sincos(pa[2], &sa, &ca);
    double dx = pb[0] - pa[0], dy = pb[1] - pa[1];
    double zh0 = ca * dx + sa * dy;
    double zh1 = -sa * dx + ca * dy;
    double zh2 = pb[2] - pa[2];
    Ja[0] = -ca; Ja[1] = -sa; Ja[2] = -sa * dx + ca * dy;
    Ja[3] = sa;  Ja[4] = -ca; Ja[5] = -ca * dx - sa * dy;
    Ja[6] = 0.0; Ja[7] = 0.0; Ja[8] = -1.0;
    Jb[0] = ca;  Jb[1] = sa;  Jb[2] = 0.0;
    Jb[3] = -sa; Jb[4] = ca;  Jb[5] = 0.0;
    Jb[6] = 0.0; Jb[7] = 0.0; Jb[8] = 1.0;
    r[0] = z[0] - zh0;
    r[1] = z[1] - zh1;
    r[2] = d_mod2pi(z[2] - zh2);
}

// ------------------------------------------------------------------------------------------
// kernel 1: linearise + scatter
// ------------------------------------------------------------------------------------------
struct LinArgs {
    const int *f_type, *f_a, *f_b, *f_slot;
    const double *f_z, *f_W;
    const double *lp, *st;
    const double *pts; // optional, indexed from f_first
    const int *node2q;
    double *Adiag, *Aoff, *Bq;
    int f_first, f_count;
};

__global__ void __launch_bounds__(128) k_linearize(LinArgs a)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.f_count)
        return;
    int f = a.f_first + t;
    int type = a.f_type[f];
    int na = a.f_a[f];
    double z[3], W[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
        z[i] = a.f_z[3 * (size_t) f + i];
#pragma unroll
    for (int i = 0; i < 9; i++)
        W[i] = a.f_W[9 * (size_t) f + i];
    if (type == 2) { // xytpos: J = I, r = z - state   (april_graph_xytpos.c:63-102)
        double p[3];
        const double *src = a.pts ? (a.pts + 6 * (size_t) t) : (a.st + 3 * (size_t) na);
        p[0] = src[0]; p[1] = src[1]; p[2] = src[2];
        double r[3] = { z[0] - p[0], z[1] - p[1], d_mod2pi(z[2] - p[2]) };
        // J'W = W ; (J'W) J = W ; keep scalar row <= col  (aprilsam.c:171-172)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++)
                atomicAdd(&a.Adiag[9 * (size_t) na + i * 3 + j], W[i * 3 + j]);
        double g[3];
        d_av(W, r, g);
#pragma unroll
        for (int i = 0; i < 3; i++)
            atomicAdd(&a.Bq[3 * (size_t) na + i], g[i]);
        return;
    }

    // xyt factor
    int nb = a.f_b[f];
    int qa = a.node2q[na], qb = a.node2q[nb];
    double pa[3], pb[3];
    if (a.pts) {
        const double *src = a.pts + 6 * (size_t) t;
#pragma unroll
        for (int i = 0; i < 3; i++) { pa[i] = src[i]; pb[i] = src[3 + i]; }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) { pa[i] = a.lp[3 * (size_t) na + i]; pb[i] = a.lp[3 * (size_t) nb + i]; }
    }
    double Ja[9], Jb[9], r[3];
    d_xyt_eval(pa, pb, z, Ja, Jb, r);

    double JatW[9], JbtW[9], H[9], g[3];
    d_atb(Ja, W, JatW); // J_a' W
    d_atb(Jb, W, JbtW); // J_b' W

    // diagonal blocks: entries with scalar row <= col only (aprilsam.c:171-172)
    d_ab(JatW, Ja, H);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++)
            atomicAdd(&a.Adiag[9 * (size_t) na + i * 3 + j], H[i * 3 + j]);
    d_ab(JbtW, Jb, H);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++)
            atomicAdd(&a.Adiag[9 * (size_t) nb + i * 3 + j], H[i * 3 + j]);

    // off-diagonal block: the reference keeps (J_early' W J_late) where "early" is the node
    // eliminated first; the mirrored block is dropped (matters for non-symmetric W).
    // The slot is stored as S[lower node id][higher node id]; H is [early][late].
    int slot = a.f_slot[f];
    int early;
    if (qa < qb) {
        d_ab(JatW, Jb, H);
        early = na;
    } else {
        d_ab(JbtW, Ja, H);
        early = nb;
    }
    const bool early_is_lo = early == (na < nb ? na : nb);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            atomicAdd(&a.Aoff[9 * (size_t) slot + (early_is_lo ? i * 3 + j : j * 3 + i)], H[i * 3 + j]);

    d_av(JatW, r, g);
#pragma unroll
    for (int i = 0; i < 3; i++)
        atomicAdd(&a.Bq[3 * (size_t) na + i], g[i]);
    d_av(JbtW, r, g);
#pragma unroll
    for (int i = 0; i < 3; i++)
        atomicAdd(&a.Bq[3 * (size_t) nb + i], g[i]);
}

__global__ void k_hessian_reset(double *Adiag, double *Aoff, double *Bq, int n_nodes, int n_slots, int n_lambda,
                                double lambda)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    size_t nd = 9 * (size_t) n_nodes, no = 9 * (size_t) n_slots, nb = 3 * (size_t) n_nodes;
    if (i < nd) {
        int e = (int) (i % 9);
        int q = (int) (i / 9);
        Adiag[i] = ((e == 0 || e == 4 || e == 8) && q < n_lambda) ? lambda : 0.0;
    } else if (i < nd + no) {
        Aoff[i - nd] = 0.0;
    } else if (i < nd + no + nb) {
        Bq[i - nd - no] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// kernel 2: persistent multifrontal factorisation (+ fused forward solve)
// ------------------------------------------------------------------------------------------
struct FacArgs {
    const asam_sn_desc_t *sn;
    const int *ipool;
    double *arena;
    const double *Adiag, *Aoff, *Bq;
    const int *q2node;
    double *y;
    int *arrive;
    const int *tasks, *nwait;
    int ntasks;
    int *ctrl; // [0] ticket, [1] err
    int smem_doubles;
    long long spin_limit;
};

__device__ __forceinline__ int ld_volatile(const int *p) { return *((const volatile int *) p); }

// Dense partial Cholesky of the first c columns of the m x m lower-triangular front F
// (column-major, leading dimension m), right-looking, all threads of the CTA.  rhs (m)
// is carried as an extra column: on exit rhs[0..c) = L11^-1 b1, rhs[c..m) = b2 - L21 y1.
__device__ void front_partial_cholesky(double *F, double *rhs, int m, int c, int sn_id, int *err)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int k = 0; k < c; ++k) {
        __syncthreads();
        double dkk = F[k + (size_t) k * m];
        if (!(dkk > 0.0)) {
            if (tid == 0)
                atomicCAS(err, 0, 1 + sn_id);
        }
        double piv = sqrt(dkk);
        __syncthreads(); // everyone has read F[k,k]
        double *col = F + (size_t) k * m;
        for (int i = k + tid; i < m; i += nt)
            col[i] = (i == k) ? piv : col[i] / piv;
        if (tid == 0)
            rhs[k] = rhs[k] / piv;
        __syncthreads();
        const int n = m - k - 1;
        const double *lk = col + k + 1; // L[k+1.., k]
        for (int j = warp; j < n; j += nwarps) {
            double ljk = lk[j];
            double *cj = F + (size_t) (k + 1 + j) * m + (k + 1);
            for (int i = j + lane; i < n; i += 32)
                cj[i] -= lk[i] * ljk;
        }
        double yk = rhs[k];
        for (int i = tid; i < n; i += nt)
            rhs[k + 1 + i] -= lk[i] * yk;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_factor(FacArgs a)
{
    extern __shared__ double sm[];
    __shared__ int s_task;
    __shared__ int s_abort;
    const int tid = threadIdx.x, nt = blockDim.x;
    int *err = a.ctrl + 1;

    for (;;) {
        if (tid == 0) {
            s_task = atomicAdd(&a.ctrl[0], 1);
            s_abort = 0;
        }
        __syncthreads();
        const int t = s_task;
        if (t >= a.ntasks)
            break;
        const int s = a.tasks[t];
        const int nw = a.nwait[t];
        const asam_sn_desc_t d = a.sn[s];
        const int m = 3 * d.mb, c = 3 * d.cb;
        const int *seg = a.ipool + d.seg;
        const int *children = seg + 2 * d.mb;
        const int *a_slot = children + d.ch_cnt;
        const int *a_rb = a_slot + d.a_cnt;
        const int *a_cb = a_rb + d.a_cnt;
        double *Fg = a.arena + d.f_off;
        const bool use_sm = ((long long) m * m + m) <= (long long) a.smem_doubles;
        double *F = use_sm ? sm : Fg;
        double *rhs = F + (size_t) m * m;

        // ---- 1. zero the front, gather original entries ---------------------------------
        for (size_t i = tid; i < (size_t) m * m + m; i += nt)
            F[i] = 0.0;
        __syncthreads();
        for (int e = tid; e < d.cb * 9; e += nt) {
            int k = e / 9, p = (e % 9) / 3, q = e % 3; // F[row 3k+p, col 3k+q], p >= q
            if (p >= q)
                F[(3 * k + p) + (size_t) (3 * k + q) * m] =
                    a.Adiag[9 * (size_t) a.q2node[d.first + k] + q * 3 + p];
        }
        for (int e = tid; e < c; e += nt)
            rhs[e] = a.Bq[3 * (size_t) a.q2node[d.first + e / 3] + e % 3];
        for (int e = tid; e < d.a_cnt * 9; e += nt) {
            int i = e / 9, p = (e % 9) / 3, q = e % 3; // late-node component p (row), early q (col)
            const int rbf = a_rb[i];
            const int rb = rbf & ~ASAM_TR_FLAG;
            // slot is S[lo id][hi id]; flag set when the early (column) node is the higher id
            const int si = (rbf & ASAM_TR_FLAG) ? (p * 3 + q) : (q * 3 + p);
            F[(3 * rb + p) + (size_t) (3 * a_cb[i] + q) * m] = a.Aoff[9 * (size_t) a_slot[i] + si];
        }

        // ---- 2. wait for the children that are being re-factored in this launch ---------
        if (nw > 0) {
            if (tid == 0) {
                long long spins = 0;
                while (ld_volatile(&a.arrive[s]) < nw) {
                    __nanosleep(64);
                    if (++spins > a.spin_limit || ld_volatile(err) < 0) {
                        atomicCAS(err, 0, -(1 + s));
                        s_abort = 1;
                        break;
                    }
                }
                a.arrive[s] = 0;
                __threadfence();
            }
        }
        __syncthreads();
        if (s_abort)
            break;

        // ---- 3. extend-add the children's update matrices --------------------------------
        for (int ci = 0; ci < d.ch_cnt; ++ci) {
            const int cs = children[ci];
            const asam_sn_desc_t cd = a.sn[cs];
            const int cm = 3 * cd.mb, cc = 3 * cd.cb, cr = cm - cc;
            const double *CF = a.arena + cd.f_off;
            const int *crel = a.ipool + cd.seg + cd.mb; // rel[]
            const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
            for (int j = warp; j < cr; j += nwarps) {
                const int dj = 3 * crel[(cc + j) / 3] + (cc + j) % 3;
                const double *ccol = CF + (size_t) (cc + j) * cm + cc;
                double *fcol = F + (size_t) dj * m;
                for (int i = j + lane; i < cr; i += 32) {
                    const int di = 3 * crel[(cc + i) / 3] + (cc + i) % 3;
                    fcol[di] += __ldcg(ccol + i);
                }
            }
            const double *crhs = CF + (size_t) cm * cm + cc;
            for (int i = tid; i < cr; i += nt) {
                const int di = 3 * crel[(cc + i) / 3] + (cc + i) % 3;
                rhs[di] += __ldcg(crhs + i);
            }
            __syncthreads();
        }

        // ---- 4. eliminate this supernode's columns ---------------------------------------
        front_partial_cholesky(F, rhs, m, c, s, err);

        // ---- 5. publish: y, L panel + update matrix ---------------------------------------
        for (int e = tid; e < c; e += nt)
            a.y[3 * (size_t) d.first + e] = rhs[e];
        if (use_sm) {
            for (size_t i = tid; i < (size_t) m * m + m; i += nt)
                Fg[i] = F[i];
        }
        __syncthreads();
        if (tid == 0 && d.parent >= 0) {
            __threadfence();
            atomicAdd(&a.arrive[d.parent], 1);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// kernel 3: persistent back-substitution
// ------------------------------------------------------------------------------------------
struct BsArgs {
    const asam_sn_desc_t *sn;
    const int *ipool;
    const double *arena;
    const double *y;
    double *x;
    int *xdone;
    const int *btasks;
    int ntasks;
    int *ctrl; // [2] ticket, [1] err
    int epoch;
    int smem_doubles;
    long long spin_limit;
};

__global__ void __launch_bounds__(128) k_backsolve(BsArgs a)
{
    extern __shared__ double sm[]; // xs[r] | w[c]
    __shared__ int s_task;
    __shared__ int s_abort;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    int *err = a.ctrl + 1;

    for (;;) {
        if (tid == 0) {
            s_task = atomicAdd(&a.ctrl[2], 1);
            s_abort = 0;
        }
        __syncthreads();
        const int t = s_task;
        if (t >= a.ntasks)
            break;
        const int s = a.btasks[t];
        const asam_sn_desc_t d = a.sn[s];
        const int m = 3 * d.mb, c = 3 * d.cb, r = m - c;
        const int *rows = a.ipool + d.seg;
        const double *L = a.arena + d.f_off;
        const bool use_sm = (m <= a.smem_doubles);
        if (!use_sm) { // cannot happen: host sizes smem for the largest front
            if (tid == 0)
                atomicCAS(err, 0, -(1 + s));
            break;
        }
        double *xs = sm;     // r
        double *w = sm + r;  // c

        if (d.parent >= 0) {
            if (tid == 0) {
                long long spins = 0;
                while (ld_volatile(&a.xdone[d.parent]) != a.epoch) {
                    __nanosleep(64);
                    if (++spins > a.spin_limit || ld_volatile(err) < 0) {
                        atomicCAS(err, 0, -(1 + s));
                        s_abort = 1;
                        break;
                    }
                }
                __threadfence();
            }
        }
        __syncthreads();
        if (s_abort)
            break;

        for (int i = tid; i < r; i += nt)
            xs[i] = __ldcg(&a.x[3 * (size_t) rows[d.cb + i / 3] + i % 3]);
        __syncthreads();
        // w_k = y_k - sum_i L[c+i, k] * xs[i]
        for (int k = warp; k < c; k += nwarps) {
            const double *lk = L + (size_t) k * m + c;
            double acc = 0.0;
            for (int i = lane; i < r; i += 32)
                acc += lk[i] * xs[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0)
                w[k] = a.y[3 * (size_t) d.first + k] - acc;
        }
        __syncthreads();
        // L11' x1 = w  (warp 0, column k descending)
        if (warp == 0) {
            for (int k = c - 1; k >= 0; --k) {
                const double *lk = L + (size_t) k * m;
                double acc = 0.0;
                for (int j = k + 1 + lane; j < c; j += 32)
                    acc += lk[j] * w[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    acc += __shfl_down_sync(0xffffffffu, acc, o);
                if (lane == 0)
                    w[k] = (w[k] - acc) / lk[k];
                __syncwarp();
            }
            for (int k = lane; k < c; k += 32)
                a.x[3 * (size_t) d.first + k] = w[k];
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicExch(&a.xdone[s], a.epoch);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// chi2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_chi2_partial(const int *f_type, const int *f_a, const int *f_b,
                                                      const double *f_z, const double *f_W, const double *st,
                                                      int n_factors, double *partial)
{
    __shared__ double red[256];
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (f < n_factors) {
        int type = f_type[f];
        int na = f_a[f];
        double z[3], W[9], r[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
            z[i] = f_z[3 * (size_t) f + i];
#pragma unroll
        for (int i = 0; i < 9; i++)
            W[i] = f_W[9 * (size_t) f + i];
        double scale;
        if (type == 1) { // xyt at `state`, weight 0.5   (april_graph.c:86-89)
            int nb = f_b[f];
            double pa[3], pb[3], Ja[9], Jb[9];
#pragma unroll
            for (int i = 0; i < 3; i++) { pa[i] = st[3 * (size_t) na + i]; pb[i] = st[3 * (size_t) nb + i]; }
            d_xyt_eval(pa, pb, z, Ja, Jb, r);
            scale = 0.5;
        } else { // weight 1.0   (april_graph.c:90-93)
            r[0] = z[0] - st[3 * (size_t) na + 0];
            r[1] = z[1] - st[3 * (size_t) na + 1];
            r[2] = d_mod2pi(z[2] - st[3 * (size_t) na + 2]);
            scale = 1.0;
        }
        double X[3];
        d_av(W, r, X);
        v = scale * (r[0] * X[0] + r[1] * X[1] + r[2] * X[2]);
    }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) k_chi2_final(const double *partial, int n, double *out)
{
    __shared__ double red[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256)
        v += partial[i];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        out[0] = red[0];
}

__global__ void k_apply_desc(asam_sn_desc_t *sn, const int *ids, const asam_sn_desc_t *desc, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        sn[ids[i]] = desc[i];
}

__global__ void k_clear_range(double *Adiag, double *Bq, double *Aoff, int q_first, int q_count, int s_first,
                              int s_count)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    size_t nd = 9 * (size_t) q_count, nb = 3 * (size_t) q_count, no = 9 * (size_t) s_count;
    if (i < nd)
        Adiag[9 * (size_t) q_first + i] = 0.0;
    else if (i < nd + nb)
        Bq[3 * (size_t) q_first + (i - nd)] = 0.0;
    else if (i < nd + nb + no)
        Aoff[9 * (size_t) s_first + (i - nd - nb)] = 0.0;
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
ASAM_EXPORT int asam_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

ASAM_EXPORT int asam_dev_create(asam_dev_t **out)
{
    *out = nullptr;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n <= 0)
        return set_err("no CUDA device");
    int dev = 0;
    const char *e = getenv("ASAM_DEVICE");
    if (!e)
        e = getenv("LOCAL_RANK");
    if (e)
        dev = atoi(e) % n;
    CK(cudaSetDevice(dev));
    asam_dev *d = new asam_dev();
    d->device = dev;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    d->n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    d->pin_cap = 8u << 20;
    CK(cudaMallocHost((void **) &d->pin, d->pin_cap));
    for (int i = 0; i < 6; i++)
        CK(cudaEventCreate(&d->ev[i]));
    if (buf_reserve(d, d->ctrl, 8 * sizeof(int), false, true))
        return 1;

    // launch geometry: k_factor keeps a whole front in shared memory when it fits
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int want = 200 * 1024; // fronts up to m = 159 stay on chip (M3500's largest is 147)
    const char *es = getenv("ASAM_FACTOR_SMEM_KB");
    if (es)
        want = atoi(es) * 1024;
    if (want > max_optin - 1024)
        want = max_optin - 1024;
    d->fac_smem = want;
    CK(cudaFuncSetAttribute(k_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, d->fac_smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_factor, d->fac_threads, d->fac_smem));
    if (occ < 1)
        return set_err("k_factor does not fit on an SM");
    d->fac_grid = occ * d->n_sm;

    d->bs_smem = 64 * 1024;
    CK(cudaFuncSetAttribute(k_backsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, d->bs_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_backsolve, d->bs_threads, d->bs_smem));
    if (occ < 1)
        return set_err("k_backsolve does not fit on an SM");
    d->bs_grid = occ * d->n_sm;
    *out = d;
    return 0;
}

ASAM_EXPORT void asam_dev_destroy(asam_dev_t *d)
{
    if (!d)
        return;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(d->stream);
    Buf *all[] = { &d->f_type, &d->f_a, &d->f_b, &d->f_z, &d->f_W, &d->f_slot, &d->lp, &d->st, &d->node2q, &d->q2node,
                   &d->Adiag, &d->Aoff, &d->Bq, &d->y, &d->x, &d->sn, &d->ipool, &d->arena, &d->arrive,
                   &d->xdone, &d->tasks_full, &d->nwait_full, &d->btasks_full, &d->tasks_tmp, &d->nwait_tmp,
                   &d->btasks_tmp, &d->ctrl, &d->partial, &d->patch_ids, &d->patch_desc, &d->pts, &d->flush };
    for (int i = 0; i < 2; i++)
        if (d->tev[i])
            cudaEventDestroy(d->tev[i]);
    for (Buf *b : all)
        if (b->p)
            cudaFree(b->p);
    if (d->pin)
        cudaFreeHost(d->pin);
    for (int i = 0; i < 6; i++)
        if (d->ev[i])
            cudaEventDestroy(d->ev[i]);
    cudaStreamDestroy(d->stream);
    delete d;
}

ASAM_EXPORT int asam_reserve(asam_dev_t *d, int n_nodes, int n_factors, int n_slots, int n_sn, int64_t ipool_ints,
                             int64_t arena_doubles)
{
    CK(cudaSetDevice(d->device));
    size_t N = (size_t) (n_nodes > 0 ? n_nodes : 0), Fn = (size_t) (n_factors > 0 ? n_factors : 0);
    size_t S = (size_t) (n_slots > 0 ? n_slots : 0), SN = (size_t) (n_sn > 0 ? n_sn : 0);
    int rc = 0;
    rc |= buf_reserve(d, d->f_type, Fn * sizeof(int), true, false);
    rc |= buf_reserve(d, d->f_a, Fn * sizeof(int), true, false);
    rc |= buf_reserve(d, d->f_b, Fn * sizeof(int), true, false);
    rc |= buf_reserve(d, d->f_slot, Fn * sizeof(int), true, false);
    rc |= buf_reserve(d, d->f_z, Fn * 3 * sizeof(double), true, false);
    rc |= buf_reserve(d, d->f_W, Fn * 9 * sizeof(double), true, false);
    rc |= buf_reserve(d, d->lp, N * 3 * sizeof(double), true, false);
    rc |= buf_reserve(d, d->st, N * 3 * sizeof(double), true, false);
    rc |= buf_reserve(d, d->node2q, N * sizeof(int), true, false);
    rc |= buf_reserve(d, d->q2node, N * sizeof(int), true, false);
    rc |= buf_reserve(d, d->Adiag, N * 9 * sizeof(double), true, true);
    rc |= buf_reserve(d, d->Bq, N * 3 * sizeof(double), true, true);
    rc |= buf_reserve(d, d->y, N * 3 * sizeof(double), true, true);
    rc |= buf_reserve(d, d->x, N * 3 * sizeof(double), true, true);
    rc |= buf_reserve(d, d->Aoff, S * 9 * sizeof(double), true, true);
    rc |= buf_reserve(d, d->sn, SN * sizeof(asam_sn_desc_t), true, false);
    rc |= buf_reserve(d, d->arrive, SN * sizeof(int), true, true);
    rc |= buf_reserve(d, d->xdone, SN * sizeof(int), true, true);
    rc |= buf_reserve(d, d->ipool, (size_t) ipool_ints * sizeof(int), true, false);
    rc |= buf_reserve(d, d->arena, (size_t) arena_doubles * sizeof(double), true, false);
    return rc;
}

ASAM_EXPORT int asam_upload_factors(asam_dev_t *d, int first, int count, const int32_t *type, const int32_t *na,
                                    const int32_t *nb, const double *z3, const double *W9)
{
    if (count <= 0)
        return 0;
    size_t need = (size_t) first + count;
    if (need * sizeof(int) > d->f_type.cap)
        return set_err("asam_upload_factors: capacity (call asam_reserve)");
    int rc = 0;
    rc |= upload(d, (int *) d->f_type.p + first, type, count * sizeof(int));
    rc |= upload(d, (int *) d->f_a.p + first, na, count * sizeof(int));
    rc |= upload(d, (int *) d->f_b.p + first, nb, count * sizeof(int));
    rc |= upload(d, (double *) d->f_z.p + 3 * (size_t) first, z3, count * 3 * sizeof(double));
    rc |= upload(d, (double *) d->f_W.p + 9 * (size_t) first, W9, count * 9 * sizeof(double));
    return rc;
}

ASAM_EXPORT int asam_upload_points(asam_dev_t *d, int which, int first, int count, const double *p3)
{
    if (count <= 0)
        return 0;
    Buf &b = which == 0 ? d->lp : d->st;
    if (((size_t) first + count) * 3 * sizeof(double) > b.cap)
        return set_err("asam_upload_points: capacity");
    return upload(d, (double *) b.p + 3 * (size_t) first, p3, (size_t) count * 3 * sizeof(double));
}

ASAM_EXPORT int asam_upload_node2q(asam_dev_t *d, int first, int count, const int32_t *node2q)
{
    if (count <= 0)
        return 0;
    if (((size_t) first + count) * sizeof(int) > d->node2q.cap)
        return set_err("asam_upload_node2q: capacity");
    return upload(d, (int *) d->node2q.p + first, node2q, (size_t) count * sizeof(int));
}

ASAM_EXPORT int asam_upload_q2node(asam_dev_t *d, int first, int count, const int32_t *q2node)
{
    if (count <= 0)
        return 0;
    if (((size_t) first + count) * sizeof(int) > d->q2node.cap)
        return set_err("asam_upload_q2node: capacity");
    return upload(d, (int *) d->q2node.p + first, q2node, (size_t) count * sizeof(int));
}

ASAM_EXPORT int asam_upload_fslot(asam_dev_t *d, int first, int count, const int32_t *fslot)
{
    if (count <= 0)
        return 0;
    if (((size_t) first + count) * sizeof(int) > d->f_slot.cap)
        return set_err("asam_upload_fslot: capacity");
    return upload(d, (int *) d->f_slot.p + first, fslot, (size_t) count * sizeof(int));
}

ASAM_EXPORT int asam_upload_ipool(asam_dev_t *d, int64_t first, int64_t count, const int32_t *data)
{
    if (count <= 0)
        return 0;
    if ((size_t) (first + count) * sizeof(int) > d->ipool.cap)
        return set_err("asam_upload_ipool: capacity");
    return upload(d, (int *) d->ipool.p + first, data, (size_t) count * sizeof(int));
}

ASAM_EXPORT int asam_upload_desc(asam_dev_t *d, int n, const int32_t *sn_ids, const asam_sn_desc_t *desc)
{
    if (n <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    if (buf_reserve(d, d->patch_ids, (size_t) n * sizeof(int), false, false))
        return 1;
    if (buf_reserve(d, d->patch_desc, (size_t) n * sizeof(asam_sn_desc_t), false, false))
        return 1;
    if (upload(d, d->patch_ids.p, sn_ids, (size_t) n * sizeof(int)))
        return 1;
    if (upload(d, d->patch_desc.p, desc, (size_t) n * sizeof(asam_sn_desc_t)))
        return 1;
    k_apply_desc<<<(n + 127) / 128, 128, 0, d->stream>>>((asam_sn_desc_t *) d->sn.p, (const int *) d->patch_ids.p,
                                                         (const asam_sn_desc_t *) d->patch_desc.p, n);
    d->n_launch++;
    CK(cudaGetLastError());
    return 0;
}

ASAM_EXPORT int asam_hessian_reset(asam_dev_t *d, int n_nodes, int n_slots, int n_lambda, double lambda)
{
    CK(cudaSetDevice(d->device));
    size_t total = 9 * (size_t) n_nodes + 9 * (size_t) n_slots + 3 * (size_t) n_nodes;
    if (total == 0)
        return 0;
    k_hessian_reset<<<(unsigned) ((total + 255) / 256), 256, 0, d->stream>>>(
        (double *) d->Adiag.p, (double *) d->Aoff.p, (double *) d->Bq.p, n_nodes, n_slots, n_lambda, lambda);
    d->n_launch++;
    CK(cudaGetLastError());
    return 0;
}

ASAM_EXPORT int asam_hessian_clear_range(asam_dev_t *d, int q_first, int q_count, int slot_first, int slot_count)
{
    CK(cudaSetDevice(d->device));
    size_t total = 12 * (size_t) q_count + 9 * (size_t) slot_count;
    if (total == 0)
        return 0;
    k_clear_range<<<(unsigned) ((total + 255) / 256), 256, 0, d->stream>>>(
        (double *) d->Adiag.p, (double *) d->Bq.p, (double *) d->Aoff.p, q_first, q_count, slot_first, slot_count);
    d->n_launch++;
    CK(cudaGetLastError());
    return 0;
}

ASAM_EXPORT int asam_linearize(asam_dev_t *d, int f_first, int f_count, const double *pts6)
{
    if (f_count <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    LinArgs a;
    a.f_type = (const int *) d->f_type.p;
    a.f_a = (const int *) d->f_a.p;
    a.f_b = (const int *) d->f_b.p;
    a.f_slot = (const int *) d->f_slot.p;
    a.f_z = (const double *) d->f_z.p;
    a.f_W = (const double *) d->f_W.p;
    a.lp = (const double *) d->lp.p;
    a.st = (const double *) d->st.p;
    a.pts = nullptr;
    if (pts6) {
        if (buf_reserve(d, d->pts, (size_t) f_count * 6 * sizeof(double), false, false))
            return 1;
        if (upload(d, d->pts.p, pts6, (size_t) f_count * 6 * sizeof(double)))
            return 1;
        a.pts = (const double *) d->pts.p;
    }
    a.node2q = (const int *) d->node2q.p;
    a.Adiag = (double *) d->Adiag.p;
    a.Aoff = (double *) d->Aoff.p;
    a.Bq = (double *) d->Bq.p;
    a.f_first = f_first;
    a.f_count = f_count;
    if (d->timing)
        CK(cudaEventRecord(d->ev[0], d->stream));
    k_linearize<<<(f_count + 127) / 128, 128, 0, d->stream>>>(a);
    d->n_launch++;
    CK(cudaGetLastError());
    if (d->timing) {
        CK(cudaEventRecord(d->ev[1], d->stream));
        d->ev_set[0] = 1;
    }
    return 0;
}

static int launch_factor(asam_dev *d, int ntasks, const int *tasks_dev, const int *nwait_dev)
{
    if (ntasks <= 0)
        return 0;
    CK(cudaMemsetAsync(d->ctrl.p, 0, 2 * sizeof(int), d->stream)); // ticket, err
    FacArgs a;
    a.sn = (const asam_sn_desc_t *) d->sn.p;
    a.ipool = (const int *) d->ipool.p;
    a.arena = (double *) d->arena.p;
    a.Adiag = (const double *) d->Adiag.p;
    a.Aoff = (const double *) d->Aoff.p;
    a.Bq = (const double *) d->Bq.p;
    a.q2node = (const int *) d->q2node.p;
    a.y = (double *) d->y.p;
    a.arrive = (int *) d->arrive.p;
    a.tasks = tasks_dev;
    a.nwait = nwait_dev;
    a.ntasks = ntasks;
    a.ctrl = (int *) d->ctrl.p;
    a.smem_doubles = d->fac_smem / (int) sizeof(double);
    a.spin_limit = 4000000LL; // a few seconds; a dependency bug must not hang the GPU
    int grid = d->fac_grid < ntasks ? d->fac_grid : ntasks;
    if (d->timing)
        CK(cudaEventRecord(d->ev[2], d->stream));
    k_factor<<<grid, d->fac_threads, d->fac_smem, d->stream>>>(a);
    d->n_launch++;
    CK(cudaGetLastError());
    if (d->timing) {
        CK(cudaEventRecord(d->ev[3], d->stream));
        d->ev_set[1] = 1;
    }
    return 0;
}

static int launch_backsolve(asam_dev *d, int ntasks, const int *btasks_dev)
{
    if (ntasks <= 0)
        return 0;
    CK(cudaMemsetAsync((int *) d->ctrl.p + 2, 0, sizeof(int), d->stream));
    d->epoch++;
    BsArgs a;
    a.sn = (const asam_sn_desc_t *) d->sn.p;
    a.ipool = (const int *) d->ipool.p;
    a.arena = (const double *) d->arena.p;
    a.y = (const double *) d->y.p;
    a.x = (double *) d->x.p;
    a.xdone = (int *) d->xdone.p;
    a.btasks = btasks_dev;
    a.ntasks = ntasks;
    a.ctrl = (int *) d->ctrl.p;
    a.epoch = d->epoch;
    a.smem_doubles = d->bs_smem / (int) sizeof(double);
    a.spin_limit = 4000000LL;
    int grid = d->bs_grid < ntasks ? d->bs_grid : ntasks;
    if (d->timing)
        CK(cudaEventRecord(d->ev[4], d->stream));
    k_backsolve<<<grid, d->bs_threads, d->bs_smem, d->stream>>>(a);
    d->n_launch++;
    CK(cudaGetLastError());
    if (d->timing) {
        CK(cudaEventRecord(d->ev[5], d->stream));
        d->ev_set[2] = 1;
    }
    return 0;
}

ASAM_EXPORT int asam_set_full_tasks(asam_dev_t *d, int ntasks, const int32_t *tasks, const int32_t *nwait,
                                    const int32_t *btasks)
{
    CK(cudaSetDevice(d->device));
    size_t b = (size_t) ntasks * sizeof(int);
    int headroom = ntasks / 2 + 1024; // room to prepend supernodes of poses appended later
    if (buf_reserve(d, d->tasks_full, b, false, false) || buf_reserve(d, d->nwait_full, b, false, false) ||
        buf_reserve(d, d->btasks_full, b + (size_t) headroom * sizeof(int), false, false))
        return 1;
    d->bt_cap = (int) (d->btasks_full.cap / sizeof(int));
    d->bt_start = d->bt_cap - ntasks;
    d->bt_count = ntasks;
    if (upload(d, d->tasks_full.p, tasks, b) || upload(d, d->nwait_full.p, nwait, b) ||
        upload(d, (int *) d->btasks_full.p + d->bt_start, btasks, b))
        return 1;
    d->ntasks_full = ntasks;
    return 0;
}

ASAM_EXPORT int asam_btasks_prepend(asam_dev_t *d, int n, const int32_t *ids)
{
    if (n <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    if (d->bt_start < n) { // out of head-room: move the list to the end of a larger buffer
        int newcap = 2 * (d->bt_count + n) + 1024;
        void *np = nullptr;
        CK(cudaMalloc(&np, (size_t) newcap * sizeof(int)));
        int newstart = newcap - d->bt_count;
        CK(cudaMemcpyAsync((int *) np + newstart, (int *) d->btasks_full.p + d->bt_start,
                           (size_t) d->bt_count * sizeof(int), cudaMemcpyDeviceToDevice, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        CK(cudaFree(d->btasks_full.p));
        d->btasks_full.p = np;
        d->btasks_full.cap = (size_t) newcap * sizeof(int);
        d->bt_cap = newcap;
        d->bt_start = newstart;
    }
    d->bt_start -= n;
    d->bt_count += n;
    return upload(d, (int *) d->btasks_full.p + d->bt_start, ids, (size_t) n * sizeof(int));
}

ASAM_EXPORT int asam_factor_full(asam_dev_t *d)
{
    CK(cudaSetDevice(d->device));
    return launch_factor(d, d->ntasks_full, (const int *) d->tasks_full.p, (const int *) d->nwait_full.p);
}

ASAM_EXPORT int asam_factor(asam_dev_t *d, int ntasks, const int32_t *tasks, const int32_t *nwait)
{
    if (ntasks <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    size_t b = (size_t) ntasks * sizeof(int);
    if (buf_reserve(d, d->tasks_tmp, b, false, false) || buf_reserve(d, d->nwait_tmp, b, false, false))
        return 1;
    if (upload(d, d->tasks_tmp.p, tasks, b) || upload(d, d->nwait_tmp.p, nwait, b))
        return 1;
    return launch_factor(d, ntasks, (const int *) d->tasks_tmp.p, (const int *) d->nwait_tmp.p);
}

ASAM_EXPORT int asam_backsolve_full(asam_dev_t *d)
{
    CK(cudaSetDevice(d->device));
    return launch_backsolve(d, d->bt_count, (const int *) d->btasks_full.p + d->bt_start);
}

ASAM_EXPORT int asam_backsolve(asam_dev_t *d, int ntasks, const int32_t *btasks)
{
    if (ntasks <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    size_t b = (size_t) ntasks * sizeof(int);
    if (buf_reserve(d, d->btasks_tmp, b, false, false))
        return 1;
    if (upload(d, d->btasks_tmp.p, btasks, b))
        return 1;
    return launch_backsolve(d, ntasks, (const int *) d->btasks_tmp.p);
}

ASAM_EXPORT int asam_download_x(asam_dev_t *d, int q_first, int q_count, double *x3)
{
    CK(cudaSetDevice(d->device));
    return download(d, x3, (const double *) d->x.p + 3 * (size_t) q_first, (size_t) q_count * 3 * sizeof(double));
}

ASAM_EXPORT int asam_download_y(asam_dev_t *d, int q_first, int q_count, double *y3)
{
    CK(cudaSetDevice(d->device));
    return download(d, y3, (const double *) d->y.p + 3 * (size_t) q_first, (size_t) q_count * 3 * sizeof(double));
}

ASAM_EXPORT int asam_chi2(asam_dev_t *d, int n_factors, double *chi2_out)
{
    *chi2_out = 0.0;
    if (n_factors <= 0)
        return 0;
    CK(cudaSetDevice(d->device));
    int nblk = (n_factors + 255) / 256;
    if (buf_reserve(d, d->partial, ((size_t) nblk + 1) * sizeof(double), false, false))
        return 1;
    double *partial = (double *) d->partial.p;
    k_chi2_partial<<<nblk, 256, 0, d->stream>>>((const int *) d->f_type.p, (const int *) d->f_a.p,
                                                 (const int *) d->f_b.p, (const double *) d->f_z.p,
                                                 (const double *) d->f_W.p, (const double *) d->st.p, n_factors,
                                                 partial + 1);
    k_chi2_final<<<1, 256, 0, d->stream>>>(partial + 1, nblk, partial);
    d->n_launch += 2;
    CK(cudaGetLastError());
    return download(d, chi2_out, partial, sizeof(double));
}

ASAM_EXPORT int asam_factor_status(asam_dev_t *d, int *status_out)
{
    CK(cudaSetDevice(d->device));
    int ctrl[2] = { 0, 0 };
    if (download(d, ctrl, d->ctrl.p, 2 * sizeof(int)))
        return 1;
    *status_out = ctrl[1];
    return 0;
}

ASAM_EXPORT int asam_debug_read_hessian(asam_dev_t *d, int n_nodes, int n_slots, double *Adiag9, double *Aoff9,
                                        double *Bq3)
{
    CK(cudaSetDevice(d->device));
    int rc = 0;
    if (Adiag9)
        rc |= download(d, Adiag9, d->Adiag.p, (size_t) n_nodes * 9 * sizeof(double));
    if (Aoff9)
        rc |= download(d, Aoff9, d->Aoff.p, (size_t) n_slots * 9 * sizeof(double));
    if (Bq3)
        rc |= download(d, Bq3, d->Bq.p, (size_t) n_nodes * 3 * sizeof(double));
    return rc;
}

ASAM_EXPORT int asam_debug_read_front(asam_dev_t *d, int64_t f_off, int64_t count, double *out)
{
    CK(cudaSetDevice(d->device));
    return download(d, out, (const double *) d->arena.p + f_off, (size_t) count * sizeof(double));
}

ASAM_EXPORT int asam_sync(asam_dev_t *d)
{
    CK(cudaSetDevice(d->device));
    CK(cudaStreamSynchronize(d->stream));
    d->pin_off = 0;
    return 0;
}

ASAM_EXPORT int asam_counters(asam_dev_t *d, int64_t *out3)
{
    out3[0] = d->n_launch;
    out3[1] = d->n_h2d;
    out3[2] = d->n_d2h;
    return 0;
}

// Generic device-side stopwatch on the library's stream (bench.py): asam_timer_start /
// asam_timer_stop bracket any sequence of asam_* calls; _stop synchronises and returns ms.
ASAM_EXPORT int asam_timer_start(asam_dev_t *d)
{
    CK(cudaSetDevice(d->device));
    if (!d->tev[0]) {
        CK(cudaEventCreate(&d->tev[0]));
        CK(cudaEventCreate(&d->tev[1]));
    }
    CK(cudaEventRecord(d->tev[0], d->stream));
    return 0;
}

ASAM_EXPORT int asam_timer_stop(asam_dev_t *d, float *ms)
{
    CK(cudaSetDevice(d->device));
    CK(cudaEventRecord(d->tev[1], d->stream));
    CK(cudaEventSynchronize(d->tev[1]));
    CK(cudaEventElapsedTime(ms, d->tev[0], d->tev[1]));
    return 0;
}

// Evict the working set from L2 between timed iterations: overwrite a buffer larger than L2.
ASAM_EXPORT int asam_l2_flush(asam_dev_t *d)
{
    CK(cudaSetDevice(d->device));
    const size_t bytes = (size_t) 384 << 20;
    if (buf_reserve(d, d->flush, bytes, false, false))
        return 1;
    d->flush_val ^= 0x5a;
    CK(cudaMemsetAsync(d->flush.p, d->flush_val, bytes, d->stream));
    return 0;
}

ASAM_EXPORT int asam_device_info(asam_dev_t *d, int *n_sm, int *fac_grid, int *fac_smem, int *bs_grid)
{
    *n_sm = d->n_sm;
    *fac_grid = d->fac_grid;
    *fac_smem = d->fac_smem;
    *bs_grid = d->bs_grid;
    return 0;
}

ASAM_EXPORT int asam_set_timing(asam_dev_t *d, int enabled)
{
    d->timing = enabled;
    d->ev_set[0] = d->ev_set[1] = d->ev_set[2] = 0;
    return 0;
}

ASAM_EXPORT int asam_last_kernel_ms(asam_dev_t *d, float *lin_ms, float *fac_ms, float *bs_ms)
{
    CK(cudaSetDevice(d->device));
    CK(cudaStreamSynchronize(d->stream));
    float *outs[3] = { lin_ms, fac_ms, bs_ms };
    for (int i = 0; i < 3; i++) {
        float ms = 0.f;
        if (d->ev_set[i])
            CK(cudaEventElapsedTime(&ms, d->ev[2 * i], d->ev[2 * i + 1]));
        if (outs[i])
            *outs[i] = ms;
    }
    return 0;
}
