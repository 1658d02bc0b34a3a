// liger_b200 -- warp-per-column block-principal-pivoting NNLS in FP64 for SMALL batches (k <= 32).
// Same problem and pivoting rule as lgr_nnls.cu (reference: RcppPlanc::bppnnls_prod, call sites R/cINMF.R:489,501: the
// V_i / U_i and W solves have only m ~ 2000 columns per Gram).  With one thread per column such a batch occupies a
// fraction of a warp per SM and every dependent instruction pays its full latency (measured: 0.06 + 0.09 ms per
// iteration at C3 for ~1.2 pivot rounds).  Here a whole warp works on one column:
//
//   * lane f owns factor f: b_f, u_f, x_f live in one register each; the passive set F is warp-uniform;
//   * each round solves the KKT system from the smaller side (primal G_FF or dual P_AA, P = G^-1 from k_gram_inv), so
//     the system has s = min(|F|, k - |F|) <= 16 unknowns; lane t holds row t of the s x s block and the elimination is
//     a Gauss-Jordan sweep whose pivot-row entries travel by shuffle: s steps of (s - j) shuffle + FMA, all rows in
//     parallel, and the solution drops out without substitution passes;
//   * the k-vector products against the solution read matrix ROWS f (stride 33 doubles: conflict free) while the
//     solution entries travel by shuffle; feasibility tests are warp ballots, so the exchange rule is uniform.
//
// Groups whose Gram is numerically singular (*ok == 0) are skipped: launch_nnls(..., flags | 16) mops them up.
#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int WKP = 32;
constexpr int WMST = 33;   // row stride (doubles) of the staged matrices
constexpr double WEPS = 2.220446049250313e-16;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// Solves M_SS z = r_S for the index set S (s <= 16 members, ascending); lane t < s returns z_t (0 for pinned pivots).
// `rhs` is this lane's entry of the factor-indexed right-hand side.  ok = false when a pivot collapsed.
__device__ __forceinline__ double solve_block(const double* __restrict__ M, const uint32_t S, const int s, const double rhs,
                                              const int lane, bool& ok) {
    const int my = (lane < s) ? (int)__fns(S, 0, lane + 1) : 0;   // lane t: idx_t
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j >= s) break;   // warp-uniform
        const int gj = __shfl_sync(0xffffffffu, my, j);
        a[j] = M[my * WMST + gj];
    }
    double r = shfl_d(rhs, my);
    double myinv = 0.0;
    ok = true;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j >= s) break;   // warp-uniform
        {
            const double pj = shfl_d(a[j], j);
            const double orig = M[(int)__shfl_sync(0xffffffffu, my, j) * (WMST + 1)];
            double inv = 0.0;
            if (pj > orig * (WEPS * 4.0))
                inv = 1.0 / pj;
            else
                ok = false;
            if (lane == j) myinv = inv;
            const double f = (lane == j) ? 0.0 : a[j] * inv;   // multiplier of the pivot row for my row
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c > j && c < s) {   // warp-uniform
                    const double rc = shfl_d(a[c], j);
                    a[c] = fma(-f, rc, a[c]);
                }
            }
            const double rr = shfl_d(r, j);
            r = fma(-f, rr, r);
        }
    }
    return (lane < s) ? r * myinv : 0.0;
}

__global__ void __launch_bounds__(256) k_nnls_warp(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols, int k,
                                                   Counters* cnt, int flags) {
    __shared__ double Gs[WKP * WMST];
    __shared__ double Ps[WKP * WMST];
    __shared__ int next_col;
    const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
    const bool cold_start = (flags & 1) != 0;
    const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    const uint32_t mybit = 1u << lane;
    const int MAXIT = 4 * k + 16;

    long long per = (total_cols + gridDim.x - 1) / gridDim.x;
    long long pos = (long long)blockIdx.x * per;
    const long long stop = pos + per < total_cols ? pos + per : total_cols;
    unsigned long long pivots_local = 0;
    while (pos < stop) {
        int g = 0;
        while (g + 1 < ngroups && pos >= groups[g + 1].col_off) ++g;
        const NnlsGroup grp = groups[g];
        const long long gend = grp.col_off + grp.ncols;
        const int c0 = (int)(pos - grp.col_off);
        const int ccols = (int)((stop < gend ? stop : gend) - pos);
        pos += ccols;
        if (*grp.ok == 0) continue;   // singular Gram: k_nnls takes this group (CTA-uniform)
        __syncthreads();
        for (int e = tid; e < WKP * WKP; e += blockDim.x) {
            const int a = e / WKP, b = e % WKP;
            Gs[a * WMST + b] = grp.gram[e];
            Ps[a * WMST + b] = grp.ginv[e];
        }
        if (tid == 0) next_col = 0;
        __syncthreads();
        const int ld = grp.ld;
        const double* __restrict__ rhs = reinterpret_cast<const double*>(grp.rhs) + (size_t)c0 * ld;
        double* __restrict__ xout = reinterpret_cast<double*>(grp.x) + (size_t)c0 * ld;
        unsigned long long* __restrict__ masks = grp.mask + c0;
        (void)nw;
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&next_col, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= ccols) break;
            const double b = (lane < k) ? rhs[(size_t)c * ld + lane] : 0.0;
            uint32_t F = cold_start ? 0u : ((uint32_t)masks[c] & kmask);
            const double tol = warp_max(fabs(b)) * (WEPS * 8.0);
            double u = 0.0, x = 0.0;
            bool have_u = false;
            int alpha = 3, beta = k + 1, iter = 0;
            for (;;) {
                const int p = __popc(F);
                const bool dual = (k - p) < p;
                const uint32_t S = dual ? ((~F) & kmask) : F;
                const int s = dual ? (k - p) : p;
                if (dual && !have_u) {   // u = P b
                    double acc = 0.0;
                    for (int t = 0; t < k; ++t) acc = fma(Ps[lane * WMST + t], shfl_d(b, t), acc);
                    u = acc;
                    have_u = true;
                }
                const double* M = dual ? Ps : Gs;
                bool ok = true;
                const double z = solve_block(M, S, s, dual ? u : b, lane, ok);
                if (!ok && lane == 0) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
                // expand by factor id, then the product against the solved side: acc_f = sum_{g in S} M[f][g] z_g
                const int rank = __popc(S & (mybit - 1u));
                const double zt = shfl_d(z, rank);
                const double zf = (S & mybit) ? zt : 0.0;
                double acc = 0.0;
                {
                    uint32_t rem = S;
                    int t = 0;
                    while (rem) {   // warp-uniform
                        const int gidx = __ffs((int)rem) - 1;
                        rem &= rem - 1;
                        acc = fma(M[lane * WMST + gidx], shfl_d(z, t), acc);
                        ++t;
                    }
                }
                uint32_t infeas;
                if (dual) {
                    // z solves P_AA z = u_A: y_A = -z >= 0 required; x = u - P_:A z, x_F >= 0 required
                    x = u - acc;
                    const double tolx = warp_max((F & mybit) ? fabs(x) : 0.0) * (WEPS * 8.0);
                    infeas = __ballot_sync(0xffffffffu, (S & mybit) && zf > tol) |
                             __ballot_sync(0xffffffffu, (F & mybit) && x < -tolx);
                } else {
                    // x_F = z >= 0 required; y_A = G_AF x_F - b_A >= 0 required
                    const double y = acc - b;
                    const double tolx = warp_max(fabs(zf)) * (WEPS * 8.0);
                    infeas = __ballot_sync(0xffffffffu, (S & mybit) && zf < -tolx) |
                             __ballot_sync(0xffffffffu, (((~F) & kmask) & mybit) && y < -tol);
                    x = zf;
                }
                if (infeas != 0u && iter < MAXIT) {
                    const int ninf = __popc(infeas);
                    uint32_t flip;
                    if (ninf < beta) {
                        beta = ninf;
                        alpha = 3;
                        flip = infeas;
                    } else if (alpha >= 1) {
                        --alpha;
                        flip = infeas;
                    } else {
                        flip = 1u << (31 - __clz((int)infeas));
                    }
                    F ^= flip;
                    ++iter;
                    if (lane == 0) ++pivots_local;
                    continue;
                }
                if (infeas != 0u && lane == 0) atomicAdd(&cnt->nnls_unconverged, 1ull);
                break;
            }
            const double xo = ((F & mybit) && x > 0.0) ? x : 0.0;
            if (lane < k || ld == WKP) xout[(size_t)c * ld + lane] = (lane < k) ? xo : 0.0;
            if (lane == 0) masks[c] = (unsigned long long)F;
        }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

}  // namespace

bool nnls_warp_supported(int k, int kp, long long total_cols, int num_sms) {
    // small batches only: beyond ~64 columns per warp slot the thread-per-column kernel has the better throughput
    return kp == 32 && k <= 32 && total_cols <= (long long)num_sms * 16 * 64;
}

void launch_nnls_warp(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, int flags,
                      cudaStream_t st) {
    if (total_cols <= 0) return;
    const int threads = 256;
    long long want = (total_cols + 7) / 8;          // one column per warp when there are that few
    const long long cap = (long long)num_sms * 4;
    const int grid = (int)(want < cap ? want : cap);
    k_nnls_warp<<<grid, threads, 0, st>>>(groups, ngroups, total_cols, k, cnt, flags);
    LGR_CUDA(cudaGetLastError());
}

}  // namespace lgr
