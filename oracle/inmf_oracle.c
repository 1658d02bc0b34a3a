/*
 * TEST INFRASTRUCTURE ONLY (oracle/): FP64 C + OpenMP restatement of rliger's iNMF / UINMF ANLS loop.
 *
 * Used (i) as a second, independent checker in tests/ (it must agree with oracle/inmf_oracle.py, which is
 * pinned to the reference's golden fixture data/pbmcPlot.rda) and (ii) as the CPU baseline that bench.py
 * times on the GPU box's host cores ("CPU restatement of RcppPlanc inmf", never "RcppPlanc").
 * Nothing under liger_b200/ links or calls this file.
 *
 * It follows (citations into /root/reference):
 *   update equations      R/cINMF.R:477-503  (inmf_solveH / inmf_solveV / inmf_solveW)
 *   objective             src/objErr.cpp:13-31
 *   iteration order       H_i for all i, V_i for all i (old W), W (new V_i)   [pinned by data/pbmcPlot.rda]
 *   UINMF objective       R/integration.R:1109-1113 (stacked rows [E_i; P_i], [W;0] + [V_i; U_i])
 * The NNLS arithmetic (RcppPlanc::bppnnls_prod; RcppPlanc >= 2.0.0 is not in the tree) is restated from
 * Kim & Park 2011 (block principal pivoting, cold start, backup rule).
 *
 * Layout: every dense matrix is column-major double (R layout).  X_i is CSC with stacked rows (m + u_i).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXK 64

/* Solve G_FF x_F = b_F by Cholesky; idx[0..p) lists F ascending; fac is p(p+1)/2 scratch. */
static void chol_solve(int p, const int* idx, const double* G, int k, const double* b, double* x, double* fac) {
    for (int j = 0; j < p; ++j) {
        const int gj = idx[j];
        double* rowj = fac + j * (j + 1) / 2;
        double d = G[gj * k + gj];
        for (int t = 0; t < j; ++t) d -= rowj[t] * rowj[t];
        const double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
        rowj[j] = inv;
        for (int i = j + 1; i < p; ++i) {
            double* rowi = fac + i * (i + 1) / 2;
            double s = G[idx[i] * k + gj];
            for (int t = 0; t < j; ++t) s -= rowi[t] * rowj[t];
            rowi[j] = s * inv;
        }
    }
    for (int j = 0; j < p; ++j) {
        const double* rowj = fac + j * (j + 1) / 2;
        double s = b[idx[j]];
        for (int t = 0; t < j; ++t) s -= rowj[t] * x[t];
        x[j] = s * rowj[j];
    }
    for (int j = p - 1; j >= 0; --j) {
        double s = x[j];
        for (int i = j + 1; i < p; ++i) s -= fac[i * (i + 1) / 2 + j] * x[i];
        x[j] = s * fac[j * (j + 1) / 2 + j];
    }
}

/* One NNLS column: min ||Cx - b||, x >= 0, from G = C^T C (k x k) and g = C^T b.  Returns pivots. */
static int bpp_column(int k, const double* G, const double* b, double* xout, uint64_t* passive) {
    uint64_t F = 0;
    int alpha = 3, beta = k + 1, pivots = 0;
    int idx[ORC_MAXK];
    double x[ORC_MAXK], fac[ORC_MAXK * (ORC_MAXK + 1) / 2];
    const uint64_t kmask = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
    for (int iter = 0; iter < 5 * k + 50; ++iter) {
        int p = 0;
        for (int f = 0; f < k; ++f)
            if (F >> f & 1ull) idx[p++] = f;
        if (p > 0) chol_solve(p, idx, G, k, b, x, fac);
        uint64_t infeas = 0;
        for (int t = 0; t < p; ++t)
            if (x[t] < 0.0) infeas |= 1ull << idx[t];
        for (int j = 0; j < k; ++j) {
            if (F >> j & 1ull) continue;
            double y = -b[j];
            for (int t = 0; t < p; ++t) y += G[j * k + idx[t]] * x[t];
            if (y < 0.0) infeas |= 1ull << j;
        }
        if (!infeas) {
            for (int f = 0; f < k; ++f) xout[f] = 0.0;
            for (int t = 0; t < p; ++t) xout[idx[t]] = x[t] > 0.0 ? x[t] : 0.0;
            if (passive) *passive = F;
            return pivots;
        }
        const int ninf = __builtin_popcountll(infeas);
        uint64_t flip;
        if (ninf < beta) {
            beta = ninf;
            alpha = 3;
            flip = infeas;
        } else if (alpha >= 1) {
            --alpha;
            flip = infeas;
        } else {
            flip = 1ull << (63 - __builtin_clzll(infeas));
        }
        F = (F ^ flip) & kmask;
        ++pivots;
    }
    for (int f = 0; f < k; ++f) xout[f] = 0.0;
    if (passive) *passive = F;
    return -1;
}

/* X (k x c, column-major) = argmin_{X>=0} ||C X - B||  from CtC (k x k) and CtB (k x c). */
int orc_bppnnls_prod(int k, int64_t c, const double* CtC, const double* CtB, double* X, int nthreads) {
    if (k < 1 || k > ORC_MAXK) return 1;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad)
    for (int64_t j = 0; j < c; ++j)
        if (bpp_column(k, CtC, CtB + j * k, X + j * k, NULL) < 0) ++bad;
    return bad ? 2 : 0;
}

static void gram(const double* A, int64_t rows, int k, double* out /* k x k */) {
    /* out = A^T A, A column-major rows x k */
    for (int a = 0; a < k; ++a)
        for (int b = 0; b <= a; ++b) {
            double s = 0.0;
            const double* ca = A + (int64_t)a * rows;
            const double* cb = A + (int64_t)b * rows;
            for (int64_t r = 0; r < rows; ++r) s += ca[r] * cb[r];
            out[a * k + b] = out[b * k + a] = s;
        }
}

/*
 * d datasets; X_i has rows_i = m + u[i] stacked rows and n[i] columns (CSC, double values).
 * W: m x k (in: init, out: result).  V[i]: rows_i x k = [V_i; U_i] (in/out).  H[i]: n_i x k (out; used as
 * given when niter == 0).  traj: niter objective values or NULL.  Returns 0 on success.
 */
int orc_inmf(int d, int k, int niter, int m, const int* n, const int* u, const int* const* colptr,
             const int* const* rowidx, const double* const* val, const double* lambda, double* W, double* const* V,
             double* const* H, double* traj, double* objErr, int nthreads, int64_t* pivots_out) {
    if (k < 1 || k > ORC_MAXK || d < 1) return 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    const int nt = omp_get_max_threads();
#else
    const int nt = 1;
#endif
    int maxrows = 0;
    for (int i = 0; i < d; ++i)
        if (m + u[i] > maxrows) maxrows = m + u[i];
    double* sq = (double*)calloc((size_t)d, sizeof(double));
    double** Lr = (double**)calloc((size_t)d, sizeof(double*)); /* L~ row-major rows x k (gather friendly) */
    double** R = (double**)calloc((size_t)d, sizeof(double*));  /* R_i = X_i H_i, row-major rows x k */
    double* G = (double*)calloc((size_t)d * k * k, sizeof(double));
    double* Rt = (double*)calloc((size_t)nt * maxrows * k, sizeof(double)); /* per-thread partial R */
    double* Gt = (double*)calloc((size_t)nt * k * k, sizeof(double));
    double* CtB = (double*)calloc((size_t)maxrows * k, sizeof(double));
    double* Xs = (double*)calloc((size_t)maxrows * k, sizeof(double));
    int64_t pivots = 0;
    int status = 0;
    for (int i = 0; i < d; ++i) {
        const int rows = m + u[i];
        Lr[i] = (double*)calloc((size_t)rows * k, sizeof(double));
        R[i] = (double*)calloc((size_t)rows * k, sizeof(double));
        double s = 0.0;
        const int64_t nnz = colptr[i][n[i]];
#pragma omp parallel for reduction(+ : s)
        for (int64_t e = 0; e < nnz; ++e) s += val[i][e] * val[i][e];
        sq[i] = s;
    }

    for (int it = 0; it <= niter; ++it) {
        /* objective of the factors after iteration it-1 (it == niter: final value) */
        const int want_obj = (it == niter) || (traj != NULL && it > 0);
        if (want_obj && !(it == 0 && niter > 0)) {
            double total = 0.0;
            for (int i = 0; i < d; ++i) {
                const int rows = m + u[i];
                if (it == 0) { /* niter == 0: statistics of the given H */
                    memset(R[i], 0, (size_t)rows * k * sizeof(double));
                    for (int64_t c = 0; c < n[i]; ++c)
                        for (int e = colptr[i][c]; e < colptr[i][c + 1]; ++e)
                            for (int f = 0; f < k; ++f)
                                R[i][(int64_t)rowidx[i][e] * k + f] += val[i][e] * H[i][(int64_t)f * n[i] + c];
                    gram(H[i], n[i], k, G + (size_t)i * k * k);
                }
                /* L~ = [W+V; U], LtL, VtV */
                double *LtL = (double*)calloc((size_t)k * k, sizeof(double)), *VtV = (double*)calloc((size_t)k * k, sizeof(double));
                double dot = 0.0;
                for (int r = 0; r < rows; ++r)
                    for (int a = 0; a < k; ++a) {
                        const double va = V[i][(int64_t)a * rows + r];
                        const double la = va + (r < m ? W[(int64_t)a * m + r] : 0.0);
                        dot += la * R[i][(int64_t)r * k + a];
                        for (int b = 0; b < k; ++b) {
                            const double vb = V[i][(int64_t)b * rows + r];
                            const double lb = vb + (r < m ? W[(int64_t)b * m + r] : 0.0);
                            LtL[a * k + b] += la * lb;
                            VtV[a * k + b] += va * vb;
                        }
                    }
                double tr1 = 0.0, tr2 = 0.0;
                const double* Gi = G + (size_t)i * k * k;
                for (int a = 0; a < k; ++a)
                    for (int b = 0; b < k; ++b) {
                        tr1 += LtL[a * k + b] * Gi[b * k + a];
                        tr2 += VtV[a * k + b] * Gi[b * k + a];
                    }
                total += sq[i] - 2.0 * dot + tr1 + lambda[i] * tr2;
                free(LtL);
                free(VtV);
            }
            if (it == niter) *objErr = total;
            if (traj && it > 0) traj[it - 1] = total;
        }
        if (it == niter) break;

        /* ---------------- H solves: CtC = L~^T L~ + lambda V~^T V~, CtB = L~^T X  (R/cINMF.R:477-483) ----- */
        for (int i = 0; i < d; ++i) {
            const int rows = m + u[i];
            double CtC[ORC_MAXK * ORC_MAXK];
            memset(CtC, 0, sizeof(CtC));
            for (int r = 0; r < rows; ++r) {
                double* lr = Lr[i] + (int64_t)r * k;
                for (int a = 0; a < k; ++a) lr[a] = V[i][(int64_t)a * rows + r] + (r < m ? W[(int64_t)a * m + r] : 0.0);
            }
            for (int a = 0; a < k; ++a)
                for (int b = 0; b <= a; ++b) {
                    double s = 0.0;
                    for (int r = 0; r < rows; ++r)
                        s += Lr[i][(int64_t)r * k + a] * Lr[i][(int64_t)r * k + b] +
                             lambda[i] * V[i][(int64_t)a * rows + r] * V[i][(int64_t)b * rows + r];
                    CtC[a * k + b] = CtC[b * k + a] = s;
                }
            memset(Rt, 0, (size_t)nt * maxrows * k * sizeof(double));
            memset(Gt, 0, (size_t)nt * k * k * sizeof(double));
            int64_t piv = 0;
            int bad = 0;
#pragma omp parallel reduction(+ : piv, bad)
            {
#ifdef _OPENMP
                const int tid = omp_get_thread_num();
#else
                const int tid = 0;
#endif
                double* Rp = Rt + (size_t)tid * maxrows * k;
                double* Gp = Gt + (size_t)tid * k * k;
                double b[ORC_MAXK], x[ORC_MAXK];
#pragma omp for schedule(dynamic, 256)
                for (int64_t c = 0; c < n[i]; ++c) {
                    for (int f = 0; f < k; ++f) b[f] = 0.0;
                    const int e0 = colptr[i][c], e1 = colptr[i][c + 1];
                    for (int e = e0; e < e1; ++e) {
                        const double v = val[i][e];
                        const double* lr = Lr[i] + (int64_t)rowidx[i][e] * k;
                        for (int f = 0; f < k; ++f) b[f] += v * lr[f];
                    }
                    const int pv = bpp_column(k, CtC, b, x, NULL);
                    if (pv < 0) ++bad; else piv += pv;
                    for (int f = 0; f < k; ++f) H[i][(int64_t)f * n[i] + c] = x[f];
                    /* sufficient statistics of the V / W solves: R_i += x_c h_c^T, G_i += h_c h_c^T */
                    for (int e = e0; e < e1; ++e) {
                        const double v = val[i][e];
                        double* rr = Rp + (int64_t)rowidx[i][e] * k;
                        for (int f = 0; f < k; ++f) rr[f] += v * x[f];
                    }
                    for (int a = 0; a < k; ++a)
                        if (x[a] != 0.0)
                            for (int bb = 0; bb < k; ++bb) Gp[a * k + bb] += x[a] * x[bb];
                }
            }
            pivots += piv;
            if (bad) status = 2;
            double* Gi = G + (size_t)i * k * k;
            memset(Gi, 0, (size_t)k * k * sizeof(double));
            memset(R[i], 0, (size_t)rows * k * sizeof(double));
            for (int t = 0; t < nt; ++t) {
                for (int e = 0; e < k * k; ++e) Gi[e] += Gt[(size_t)t * k * k + e];
                const double* Rp = Rt + (size_t)t * maxrows * k;
#pragma omp parallel for
                for (int64_t e = 0; e < (int64_t)rows * k; ++e) R[i][e] += Rp[e];
            }
        }
        /* ---------------- V_i / U_i solves (old W): CtC = (1+lambda) G_i, CtB = R_i^T - G_i W^T  (R/cINMF.R:485-491) -- */
        for (int i = 0; i < d; ++i) {
            const int rows = m + u[i];
            double CtC[ORC_MAXK * ORC_MAXK];
            const double* Gi = G + (size_t)i * k * k;
            for (int e = 0; e < k * k; ++e) CtC[e] = (1.0 + lambda[i]) * Gi[e];
#pragma omp parallel for
            for (int r = 0; r < rows; ++r)
                for (int f = 0; f < k; ++f) {
                    double v = R[i][(int64_t)r * k + f];
                    if (r < m)
                        for (int f2 = 0; f2 < k; ++f2) v -= Gi[f * k + f2] * W[(int64_t)f2 * m + r];
                    CtB[(int64_t)r * k + f] = v;
                }
            if (orc_bppnnls_prod(k, rows, CtC, CtB, Xs, nthreads)) status = 2;
            for (int r = 0; r < rows; ++r)
                for (int f = 0; f < k; ++f) V[i][(int64_t)f * rows + r] = Xs[(int64_t)r * k + f];
        }
        /* ---------------- W solve (new V_i): CtC = sum G_i, CtB = sum (R_i^T - G_i V_i^T)  (R/cINMF.R:493-503) ---- */
        {
            double CtC[ORC_MAXK * ORC_MAXK];
            memset(CtC, 0, sizeof(CtC));
            for (int i = 0; i < d; ++i)
                for (int e = 0; e < k * k; ++e) CtC[e] += G[(size_t)i * k * k + e];
#pragma omp parallel for
            for (int r = 0; r < m; ++r)
                for (int f = 0; f < k; ++f) {
                    double v = 0.0;
                    for (int i = 0; i < d; ++i) {
                        const int rows = m + u[i];
                        const double* Gi = G + (size_t)i * k * k;
                        double acc = R[i][(int64_t)r * k + f];
                        for (int f2 = 0; f2 < k; ++f2) acc -= Gi[f * k + f2] * V[i][(int64_t)f2 * rows + r];
                        v += acc;
                    }
                    CtB[(int64_t)r * k + f] = v;
                }
            if (orc_bppnnls_prod(k, m, CtC, CtB, Xs, nthreads)) status = 2;
            for (int r = 0; r < m; ++r)
                for (int f = 0; f < k; ++f) W[(int64_t)f * m + r] = Xs[(int64_t)r * k + f];
        }
    }
    if (pivots_out) *pivots_out = pivots;
    for (int i = 0; i < d; ++i) {
        free(Lr[i]);
        free(R[i]);
    }
    free(sq);
    free(Lr);
    free(R);
    free(G);
    free(Rt);
    free(Gt);
    free(CtB);
    free(Xs);
    return status;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
