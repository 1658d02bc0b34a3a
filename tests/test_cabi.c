/* Plain-C caller of the liger_b200 ABI (include/liger_b200.h), compiled with gcc by tests/test_abi.py.
 *
 *   test_cabi nogpu   argument errors, the R RNG known answer, and -- without a device -- LGR_ERR_NO_DEVICE
 *   test_cabi gpu     a small iNMF (explicit inits, then in-library inits), lgr_bppnnls_prod and lgr_objerr_i, online iNMF from
 *                     arrays and through a column-chunk reader
 *
 * The numbers checked here need no oracle: closed forms (diagonal Gram), invariants (non-negativity, identical
 * results of two identical calls up to fp32 reduction order, objective decreasing) and R's documented
 * set.seed(1); runif(3) = 0.2655087 0.3721239 0.5728534.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "liger_b200.h"

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #cond, lgr_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

/* deterministic toy matrix: m x n, ~30 % dense, values in (0, 2) */
static void toy_csc(int m, int n, unsigned seed, int32_t** p, int32_t** i, double** x) {
    *p = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t));
    *i = (int32_t*)malloc((size_t)m * n * sizeof(int32_t));
    *x = (double*)malloc((size_t)m * n * sizeof(double));
    unsigned s = seed;
    int nnz = 0;
    for (int c = 0; c < n; ++c) {
        (*p)[c] = nnz;
        for (int r = 0; r < m; ++r) {
            s = s * 1664525u + 1013904223u;
            if ((s >> 24) < 77) {
                (*i)[nnz] = r;
                (*x)[nnz] = 0.05 + 1.9 * (double)((s >> 8) & 0xffff) / 65536.0;
                ++nnz;
            }
        }
    }
    (*p)[n] = nnz;
}

/* a column-chunk reader over in-memory CSC arrays (stands in for an HDF5 reader): see lgr_chunk_source */
typedef struct toy_source {
    int32_t* const* p;
    int32_t* const* ri;
    double* const* x;
    int calls;
} toy_source;
static int64_t toy_read_cols(void* user, int32_t ds, int64_t col0, int64_t ncol, int32_t* colptr, int32_t* rowidx, double* values,
                             int64_t capacity) {
    toy_source* S = (toy_source*)user;
    const int32_t b = S->p[ds][col0], e = S->p[ds][col0 + ncol];
    S->calls += 1;
    if (e - b > capacity) return -(int64_t)(e - b);
    for (int64_t j = 0; j <= ncol; ++j) colptr[j] = S->p[ds][col0 + j] - b;
    memcpy(rowidx, S->ri[ds] + b, (size_t)(e - b) * sizeof(int32_t));
    memcpy(values, S->x[ds] + b, (size_t)(e - b) * sizeof(double));
    return e - b;
}

static int run_nogpu(void) {
    CHECK(lgr_version() == LGR_VERSION_MAJOR * 1000 + LGR_VERSION_MINOR);
    lgr_rng* r = lgr_r_set_seed(1);
    double u[3];
    lgr_r_runif(r, 3, 0.0, 1.0, u);
    lgr_r_free(r);
    CHECK(fabs(u[0] - 0.2655087) < 1e-7 && fabs(u[1] - 0.3721239) < 1e-7 && fabs(u[2] - 0.5728534) < 1e-7);
    CHECK(lgr_inmf(NULL, NULL) == LGR_ERR_INVALID);
    CHECK(strlen(lgr_last_error()) > 0);
    double G[4] = {1, 0, 0, 1}, B[2] = {1, -1}, X[2];
    if (lgr_device_count() == 0) {
        CHECK(lgr_bppnnls_prod(2, 1, G, B, X, -1) == LGR_ERR_NO_DEVICE);   /* no CPU fallback */
        lgr_ctx* ctx = NULL;
        lgr_inmf_args a;
        memset(&a, 0, sizeof(a));
        CHECK(lgr_ctx_create(&a, &ctx) != LGR_OK && ctx == NULL);
    }
    CHECK(lgr_bppnnls_prod(0, 1, G, B, X, -1) != LGR_OK);     /* k out of range */
    printf("cabi nogpu ok\n");
    return 0;
}

static int run_gpu(void) {
    CHECK(lgr_device_count() > 0);
    /* bppnnls_prod on a diagonal Gram: x = max(b / g, 0) */
    {
        double G[9] = {2, 0, 0, 0, 4, 0, 0, 0, 0.5}, B[6] = {1, -1, 2, -3, 8, 0.25}, X[6];
        CHECK(lgr_bppnnls_prod(3, 2, G, B, X, -1) == LGR_OK);
        const double want[6] = {0.5, 0, 4, 0, 2, 0.5};
        for (int j = 0; j < 6; ++j) CHECK(fabs(X[j] - want[j]) < 1e-12);
    }
    enum { D = 2, M = 40, K = 5, NITER = 6 };
    const int n[D] = {90, 61};
    int32_t *p[D], *ri[D];
    double* x[D];
    lgr_csc E[D];
    memset(E, 0, sizeof(E));   /* optional fields (row_scale / col_scale) = NULL */
    for (int i = 0; i < D; ++i) {
        toy_csc(M, n[i], 17u + (unsigned)i, &p[i], &ri[i], &x[i]);
        E[i].nrow = M;
        E[i].ncol = n[i];
        E[i].colptr = p[i];
        E[i].rowidx = ri[i];
        E[i].values = x[i];
        E[i].value_dtype = LGR_F64;
        E[i].location = 0;
    }
    double lam[D] = {5.0, 5.0};
    /* explicit inits drawn with the library's R stream, in .runINMF.list's order */
    lgr_rng* rng = lgr_r_set_seed(7);
    double W0[M * K], V0a[M * K], V0b[M * K];
    lgr_r_runif(rng, M * K, 0.0, 2.0, W0);
    lgr_r_runif(rng, M * K, 0.0, 2.0, V0a);
    lgr_r_runif(rng, M * K, 0.0, 2.0, V0b);
    lgr_r_free(rng);
    const double* V0[D] = {V0a, V0b};
    lgr_inmf_args a;
    memset(&a, 0, sizeof(a));
    a.d = D;
    a.k = K;
    a.niter = NITER;
    a.flags = LGR_FLAG_TRAJECTORY;
    a.m = M;
    a.E = E;
    a.lambda = lam;
    a.Winit = W0;
    a.Vinit = V0;
    a.device = -1;
    a.world = 1;
    double W[M * K], Va[M * K], Vb[M * K], traj[NITER];
    double* Ha = (double*)calloc((size_t)n[0] * K, sizeof(double));
    double* Hb = (double*)calloc((size_t)n[1] * K, sizeof(double));
    double* V[D] = {Va, Vb};
    double* H[D] = {Ha, Hb};
    lgr_inmf_out o;
    memset(&o, 0, sizeof(o));
    o.W = W;
    o.V = V;
    o.H = H;
    o.obj_traj = traj;
    CHECK(lgr_inmf(&a, &o) == LGR_OK);
    for (int t = 1; t < NITER; ++t) CHECK(traj[t] < traj[t - 1]);
    CHECK(fabs(o.objErr - traj[NITER - 1]) < 1e-6 * traj[NITER - 1]);
    for (int j = 0; j < M * K; ++j) CHECK(W[j] >= 0 && Va[j] >= 0 && isfinite(W[j]));
    for (int j = 0; j < n[0] * K; ++j) CHECK(Ha[j] >= 0 && isfinite(Ha[j]));
    CHECK(o.timings.nnls_unconverged == 0 && o.timings.kernel_launches > 0);
    /* objErr_i of the returned factors sums to the returned objective (H goes in k x n, src/objErr.cpp:13) */
    {
        double total = 0;
        for (int i = 0; i < D; ++i) {
            double* Ht = (double*)malloc((size_t)n[i] * K * sizeof(double));
            for (int c = 0; c < n[i]; ++c)
                for (int f = 0; f < K; ++f) Ht[(size_t)c * K + f] = H[i][(size_t)f * n[i] + c];
            double e = 0;
            CHECK(lgr_objerr_i(K, Ht, W, V[i], &E[i], lam[i], -1, &e) == LGR_OK);
            total += e;
            free(Ht);
        }
        CHECK(fabs(total - o.objErr) < 1e-5 * o.objErr);
    }
    /* the in-library initialisation (Winit = Vinit = NULL, RcppPlanc's NULL-init call of R/suggestK.R:78-88) with
       init_seed = 7 draws the same W, V_i as above: same factors */
    {
        lgr_inmf_args b = a;
        b.Winit = NULL;
        b.Vinit = NULL;
        b.init_seed = 7;
        b.flags = 0;
        double W2[M * K];
        lgr_inmf_out o2;
        memset(&o2, 0, sizeof(o2));
        o2.W = W2;
        CHECK(lgr_inmf(&b, &o2) == LGR_OK);
        double num = 0, den = 0;
        for (int j = 0; j < M * K; ++j) {
            num += (W2[j] - W[j]) * (W2[j] - W[j]);
            den += W[j] * W[j];
        }
        CHECK(sqrt(num / den) < 1e-5);
        CHECK(fabs(o2.objErr - o.objErr) < 1e-6 * o.objErr);
    }
    /* a malformed dgCMatrix (row index out of range) is rejected, not written through */
    {
        const int32_t saved = ri[0][3];
        ri[0][3] = M + 5;
        CHECK(lgr_inmf(&a, &o) == LGR_ERR_INVALID);
        ri[0][3] = -1;
        CHECK(lgr_inmf(&a, &o) == LGR_ERR_INVALID);
        ri[0][3] = saved;
        CHECK(lgr_inmf(&a, &o) == LGR_OK);
    }
    /* online iNMF from the arrays and through a reader of column chunks (lgr_online_inmf_chunked): same factors */
    {
        lgr_online_args oa;
        memset(&oa, 0, sizeof(oa));
        oa.d = D;
        oa.k = K;
        oa.m = M;
        oa.E = E;
        oa.lambda = 5.0;
        oa.max_epochs = 3;
        oa.minibatch_size = 30;
        oa.hals_iter = 1;
        oa.seed = 4;
        oa.device = -1;
        double Wa[M * K], Wb[M * K], Vx[D][M * K];
        double* Vp[D] = {Vx[0], Vx[1]};
        double* Hp[D] = {Ha, Hb};
        lgr_online_out oo;
        memset(&oo, 0, sizeof(oo));
        oo.W = Wa;
        oo.V = Vp;
        oo.H = Hp;
        CHECK(lgr_online_inmf(&oa, &oo) == LGR_OK);
        CHECK(oo.iterations == (n[0] + n[1]) * 3 / 30 && oo.objErr > 0);
        const double obj_a = oo.objErr;
        lgr_csc shapes[D];
        memset(shapes, 0, sizeof(shapes));
        int64_t nnz[D];
        for (int i = 0; i < D; ++i) {
            shapes[i].nrow = M;
            shapes[i].ncol = n[i];
            nnz[i] = p[i][n[i]];
        }
        toy_source ts = {p, ri, x, 0};
        lgr_chunk_source src = {toy_read_cols, &ts, nnz, 16, 8};   /* 16 columns per chunk; capacity 8: the reader asks for more */
        oa.E = shapes;
        oo.W = Wb;
        CHECK(lgr_online_inmf_chunked(&oa, &src, &oo) == LGR_OK);
        CHECK(ts.calls >= (n[0] + 15) / 16 + (n[1] + 15) / 16 + 2 * K);
        double num = 0, den = 0;
        for (int j = 0; j < M * K; ++j) {
            num += (Wb[j] - Wa[j]) * (Wb[j] - Wa[j]);
            den += Wa[j] * Wa[j];
        }
        CHECK(sqrt(num / den) < 1e-6);
        CHECK(fabs(oo.objErr - obj_a) < 1e-6 * obj_a);
        nnz[1] += 3;   /* announced more than the reader delivers */
        CHECK(lgr_online_inmf_chunked(&oa, &src, &oo) == LGR_ERR_INVALID);
        CHECK(lgr_online_inmf_chunked(&oa, NULL, &oo) == LGR_ERR_INVALID);
    }
    printf("cabi gpu ok objErr=%.6f\n", o.objErr);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) return run_gpu();
    return run_nogpu();
}
