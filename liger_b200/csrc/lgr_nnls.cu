// liger_b200 -- batched NNLS by block principal pivoting (Kim & Park 2011) on a shared Gram matrix.
// Replaces RcppPlanc::bppnnls_prod (reference call sites R/cINMF.R:481,489,501) and the H / V / W solves
// inside RcppPlanc::inmf / uinmf.
//
//   min_x>=0  1/2 x^T G x - b^T x     for every column b of a group that shares the k x k Gram G.
//
// Design (one thread per column, column state in shared memory):
//   * passive set F as a 64-bit mask; warm-started from the previous ANLS iteration.
//   * every pivot round solves the KKT system exactly for the current F.  Because G is shared by all
//     columns of the group, its inverse P = G^-1 (fp64, k_gram_inv) is available, and the system can be
//     solved from whichever side is smaller:
//        primal (|F| = p <= k - p):  x_F = G_FF^-1 b_F,             y_A = G_AF x_F - b_A
//        dual   (|A| = q <  p)    :  y_A = -P_AA^-1 u_A,  u = P b,  x_F = u_F + P_FA y_A
//     so the per-column Cholesky never exceeds k/2 (work ~ s^3/6 with s = min(p, q)); u = P b is formed
//     once per column with register accumulators and broadcast shared-memory reads of P.
//   * the per-column Cholesky factor lives in a shared-memory POOL, allocated per warp and per round as
//     [element][lane] (bank-conflict free).  Slots are counting-sorted by s every round so that the lanes
//     of a warp do similar work, and finished slots are compacted away (SIMT efficiency).
//   * if G is numerically singular (dead factor) the group falls back to primal-only solves.
#include <type_traits>

#include "lgr_common.cuh"

namespace lgr {

template <typename T>
struct Eps;
template <>
struct Eps<float> {
    static __device__ __forceinline__ float v() { return 1.1920929e-7f; }
};
template <>
struct Eps<double> {
    static __device__ __forceinline__ double v() { return 2.220446049250313e-16; }
};
__device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsq(double x) { return 1.0 / sqrt(x); }

// ------------------------------------------------------------------------------------------------
// k_gram_inv: one CTA per group.  gram = cA*A + cB*B (fp64), ginv = gram^-1 by Gauss-Jordan (SPD: no
// pivoting), ok = 0 when a pivot collapses (then the NNLS kernel uses primal solves only).
// ------------------------------------------------------------------------------------------------
template <int KP, int THREADS>
__global__ void __launch_bounds__(THREADS) k_gram_inv(const GramSpec* __restrict__ specs, int k) {
    const GramSpec sp = specs[blockIdx.x];
    __shared__ double M[KP][KP + 1];
    __shared__ double colbuf[KP];
    __shared__ double rowbuf[KP];
    __shared__ int ok_s;
    const int tid = threadIdx.x;
    constexpr int EPT = KP * KP / THREADS;   // elements per thread (1 for KP = 32, 4 for KP = 64)
    for (int e = tid; e < KP * KP; e += blockDim.x) {
        const int a = e / KP, b = e % KP;
        double v = 0.0;
        if (a < k && b < k) {
            v = sp.cA * sp.A[e];
            if (sp.B) v += sp.cB * sp.B[e];
        }
        sp.gram[e] = v;
        M[a][b] = v;
    }
    if (tid == 0) ok_s = 1;
    __syncthreads();
    double maxdiag = 0.0;
    for (int a = 0; a < k; ++a) maxdiag = fmax(maxdiag, M[a][a]);
    for (int c = 0; c < k; ++c) {  // in-place Gauss-Jordan, SPD => no pivoting; two barriers per step
        __syncthreads();   // the previous step's update is complete
        const double piv = M[c][c];
        if (!(piv > 1e-11 * maxdiag)) {
            if (tid == 0) ok_s = 0;
            break;   // uniform: every thread reads the same pivot
        }
        const double ip = 1.0 / piv;
        if (tid < KP) {
            colbuf[tid] = M[tid][c];                              // column c before the step
            rowbuf[tid] = (tid == c) ? ip : M[c][tid] * ip;       // row c after the step
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int e = tid + j * THREADS;
            const int r = e / KP, col = e % KP;
            if (r < k && col < k) {
                if (r == c) {
                    M[r][col] = rowbuf[col];
                } else {
                    const double base = (col == c) ? 0.0 : M[r][col];
                    M[r][col] = base - colbuf[r] * rowbuf[col];
                }
            }
        }
    }
    __syncthreads();
    const int ok = ok_s;
    for (int e = tid; e < KP * KP; e += blockDim.x) {
        const int a = e / KP, b = e % KP;
        double v = 0.0;
        if (ok && a < k && b < k) v = 0.5 * (M[a][b] + M[b][a]);
        sp.ginv[e] = v;
    }
    if (tid == 0) *sp.ok = ok;
}

void launch_gram_inv(const GramSpec* specs, int nspecs, int k, int kp, cudaStream_t st) {
    if (nspecs <= 0) return;
    if (kp == 32)
        k_gram_inv<32, 1024><<<nspecs, 1024, 0, st>>>(specs, k);
    else
        k_gram_inv<64, 1024><<<nspecs, 1024, 0, st>>>(specs, k);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Solve M_II z = r_I by Cholesky (left-looking, packed lower-triangular factor in the warp's pool).
//   idx[t*32]   t-th member of the index set I (ascending), one byte per entry, [t][lane]
//   M           k x k symmetric matrix in shared memory (row stride MST)
//   r[f*33]     right-hand side indexed by factor id, [f][lane] tile (stride 33)
//   fac[e*FS]   packed factor, e = i(i+1)/2 + j;  the diagonal stores 1/L_jj
//   z[t*FS]     solution in compact order (lives in the pool right after the factor)
// A non-positive pivot pins that coordinate to zero (numerically singular principal block).
// ------------------------------------------------------------------------------------------------
template <typename T, int MST>
__device__ __forceinline__ bool chol_solve(int n, const unsigned char* idx, const T* __restrict__ M, const T* r, T* fac,
                                           T* z, const int FS) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        const int gj = idx[j * 32];
        const T* Mj = M + gj * MST;
        T* rowj = fac + (j * (j + 1) / 2) * FS;
        T d = Mj[gj];
#pragma unroll 4
        for (int t = 0; t < j; ++t) {
            const T l = rowj[t * FS];
            d -= l * l;
        }
        T invd;
        if (d > Mj[gj] * (Eps<T>::v() * (T)4)) {
            invd = rsq(d);
        } else {
            invd = (T)0;
            ok = false;
        }
        rowj[j * FS] = invd;
        int i = j + 1;
        for (; i + 1 < n; i += 2) {  // two rows at a time: shares row j, doubles the ILP of the FMA chain
            T* r0 = fac + (i * (i + 1) / 2) * FS;
            T* r1 = r0 + (i + 1) * FS;
            T s0 = Mj[idx[i * 32]], s1 = Mj[idx[(i + 1) * 32]];
#pragma unroll 2
            for (int t = 0; t < j; ++t) {
                const T l = rowj[t * FS];
                s0 -= r0[t * FS] * l;
                s1 -= r1[t * FS] * l;
            }
            r0[j * FS] = s0 * invd;
            r1[j * FS] = s1 * invd;
        }
        if (i < n) {
            T* r0 = fac + (i * (i + 1) / 2) * FS;
            T s0 = Mj[idx[i * 32]];
#pragma unroll 4
            for (int t = 0; t < j; ++t) s0 -= r0[t * FS] * rowj[t * FS];
            r0[j * FS] = s0 * invd;
        }
    }
    for (int j = 0; j < n; ++j) {  // forward substitution
        const T* rowj = fac + (j * (j + 1) / 2) * FS;
        T s = r[idx[j * 32] * 33];
#pragma unroll 4
        for (int t = 0; t < j; ++t) s -= rowj[t * FS] * z[t * FS];
        z[j * FS] = s * rowj[j * FS];
    }
    for (int j = n - 1; j >= 0; --j) {  // back substitution
        T s = z[j * FS];
        const T* col = fac + ((j + 1) * (j + 2) / 2 + j) * FS;  // element (j+1, j)
        for (int i = j + 1; i < n; ++i) {
            s -= col[0] * z[i * FS];
            col += (i + 1) * FS;
        }
        z[j * FS] = s * fac[((j * (j + 1) / 2) + j) * FS];
    }
    return ok;
}

// shared-memory carve-up (host and device must agree)
struct NnlsSmem {
    size_t G, P, warp0, per_warp, b, u, pool, idx, total;
};
__host__ __device__ inline NnlsSmem nnls_layout(int kp, int nwarps, int pool_elems, size_t ts) {
    NnlsSmem L;
    const size_t mst = (size_t)kp + 4;  // row stride of G / P: = 4 (mod 32) words -> conflict-free 128-bit row reads
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += (bytes + 15) & ~(size_t)15;
        return at;
    };
    L.G = take((size_t)kp * mst * ts);
    L.P = take((size_t)kp * mst * ts);
    L.warp0 = o;
    size_t w = 0;
    auto wtake = [&](size_t bytes) {
        const size_t at = w;
        w += (bytes + 15) & ~(size_t)15;
        return at;
    };
    L.b = wtake((size_t)kp * 33 * ts);
    L.u = wtake((size_t)kp * 33 * ts);
    L.pool = wtake((size_t)pool_elems * ts);
    L.idx = wtake((size_t)kp * 32);
    L.per_warp = w;
    L.total = o + (size_t)nwarps * w;
    return L;
}

// One pivot round of one column (executed by the lane that owns it).  Returns true when the column stays active.
template <typename T, int KP>
__device__ __forceinline__ bool nnls_round(const int k, const unsigned long long kmask, const int dual_ok, const int MAXIT,
                                           const T* __restrict__ Gs, const T* __restrict__ Ps, T* bcol, const T* ucol,
                                           unsigned char* idx, T* fac, const int FS, unsigned long long& F, int& state,
                                           const T tol, Counters* cnt, unsigned long long& pivots, const int hist_mode) {
    constexpr int MST = KP + 4;
    const int p = __popcll(F);
    const bool dual = dual_ok && (k - p) < p;
    const int s = dual ? (k - p) : p;
    {
        unsigned long long r = dual ? ((~F) & kmask) : F;
        int t = 0;
        while (r) {
            idx[t * 32] = (unsigned char)(__ffsll((long long)r) - 1);
            r &= r - 1;
            ++t;
        }
    }
    T* z = fac + (s * (s + 1) / 2) * FS;
    bool ok = true;
    if (s > 0) ok = chol_solve<T, MST>(s, idx, dual ? Ps : Gs, dual ? ucol : bcol, fac, z, FS);
    if (!ok) atomicAdd(&cnt->nnls_bad_pivot, 1ull);

    // all k entries of  x = u - P_:A z  (dual)  or  y = G_:F x_F - b  (primal)  with register accumulators
    T acc[KP];
    const T* init = dual ? ucol : bcol;
#pragma unroll
    for (int f = 0; f < KP; ++f) acc[f] = init[f * 33];
    if (!dual) {
#pragma unroll
        for (int f = 0; f < KP; ++f) acc[f] = -acc[f];
    }
    const T* Mat = dual ? Ps : Gs;
    const T sgn = dual ? (T)-1 : (T)1;
    for (int t = 0; t < s; ++t) {
        const T zt = sgn * z[t * FS];
        const T* row = Mat + idx[t * 32] * MST;
#pragma unroll
        for (int f = 0; f < KP; ++f) acc[f] = fma(zt, row[f], acc[f]);
    }
    unsigned long long infeas = 0ull;
    if (dual) {
        // y_A = -z >= 0 required ; x_F = acc_F >= 0 required
        for (int t = 0; t < s; ++t)
            if (z[t * FS] > tol) infeas |= 1ull << idx[t * 32];
        T xmax = (T)0;
#pragma unroll
        for (int f = 0; f < KP; ++f)
            if ((F >> f) & 1ull) xmax = fmax(xmax, fabs(acc[f]));
        const T tolx = xmax * Eps<T>::v() * (T)8;
#pragma unroll
        for (int f = 0; f < KP; ++f)
            if (((F >> f) & 1ull) && acc[f] < -tolx) infeas |= 1ull << f;
    } else {
        T xmax = (T)0;
        for (int t = 0; t < s; ++t) xmax = fmax(xmax, fabs(z[t * FS]));
        const T tolx = xmax * Eps<T>::v() * (T)8;
        for (int t = 0; t < s; ++t)
            if (z[t * FS] < -tolx) infeas |= 1ull << idx[t * 32];
        const unsigned long long A = (~F) & kmask;
#pragma unroll
        for (int f = 0; f < KP; ++f)
            if (((A >> f) & 1ull) && acc[f] < -tol) infeas |= 1ull << f;
    }
    int alpha = state & 0xff, beta = (state >> 8) & 0xff, iter = state >> 16;
    if (infeas != 0ull && iter < MAXIT) {
        const int ninf = __popcll(infeas);
        unsigned long long flip;
        if (ninf < beta) {
            beta = ninf;
            alpha = 3;
            flip = infeas;
        } else if (alpha >= 1) {
            --alpha;
            flip = infeas;
        } else {
            flip = 1ull << (63 - __clzll((long long)infeas));
        }
        F ^= flip;
        state = alpha | (beta << 8) | ((iter + 1) << 16);
        ++pivots;
        return true;
    }
    if (infeas != 0ull) atomicAdd(&cnt->nnls_unconverged, 1ull);
    if (hist_mode) atomicAdd(&cnt->hist[hist_mode == 2 ? (p < 39 ? p : 39) : (iter < 39 ? iter : 39)], 1ull);
    // final x, expanded by factor id, over b's column of the tile (b is no longer needed)
    if (dual) {
#pragma unroll
        for (int f = 0; f < KP; ++f) {
            T v = acc[f];
            if (!((F >> f) & 1ull) || !(v > (T)0)) v = (T)0;
            bcol[f * 33] = v;
        }
    } else {
#pragma unroll
        for (int f = 0; f < KP; ++f) bcol[f * 33] = (T)0;
        for (int t = 0; t < s; ++t) {
            const T v = z[t * FS];
            bcol[idx[t * 32] * 33] = v > (T)0 ? v : (T)0;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// k_nnls: persistent CTAs.  The columns of all groups form one flat range that is split evenly over the CTAs; a CTA
// walks its range group by group (G / P are loaded once per group, so a CTA reloads them at most a few times).
// Inside a group's chunk every LANE owns one column from load to store and claims the next column from a shared
// counter as soon as its current one is solved ("lane refill"): there is no barrier of any kind between columns,
// so a column that needs several pivot rounds delays nobody, and the warp stays full until the chunk runs dry.
// ------------------------------------------------------------------------------------------------
template <typename T, int KP>
__device__ __forceinline__ void load_row(const T* __restrict__ r, T* bcol, int k, int ld) {
    if (ld == KP) {  // padded row, 16-byte aligned: vector loads
        constexpr int VEC = 16 / sizeof(T);
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
        const V* rv = reinterpret_cast<const V*>(r);
#pragma unroll
        for (int i = 0; i < KP / VEC; ++i) {
            const V v = rv[i];
            const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) bcol[(i * VEC + j) * 33] = e[j];
        }
    } else {
        for (int f = 0; f < KP; ++f) bcol[f * 33] = f < k ? r[f] : (T)0;
    }
}
template <typename T, int KP>
__device__ __forceinline__ void store_row(T* __restrict__ x, const T* bcol, int k, int ld) {
    if (ld == KP) {
        constexpr int VEC = 16 / sizeof(T);
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
        V* xv = reinterpret_cast<V*>(x);
#pragma unroll
        for (int i = 0; i < KP / VEC; ++i) {
            V v;
            T* e = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) e[j] = bcol[(i * VEC + j) * 33];
            xv[i] = v;
        }
    } else {
        for (int f = 0; f < k; ++f) x[f] = bcol[f * 33];
    }
}

template <typename T, int KP>
__global__ void __launch_bounds__(256, 1) k_nnls(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols,
                                                 int k, int pool_elems, Counters* cnt, int flags) {
    constexpr int MST = KP + 4;
    const int NW = blockDim.x >> 5;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NnlsSmem L = nnls_layout(KP, NW, pool_elems, sizeof(T));
    T* Gs = reinterpret_cast<T*>(smem_raw + L.G);
    T* Ps = reinterpret_cast<T*>(smem_raw + L.P);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* wbase = smem_raw + L.warp0 + (size_t)warp * L.per_warp;
    T* bS = reinterpret_cast<T*>(wbase + L.b);     // [f][lane], stride 33
    T* uS = reinterpret_cast<T*>(wbase + L.u);
    T* pool = reinterpret_cast<T*>(wbase + L.pool);
    unsigned char* idxS = wbase + L.idx;           // [t][lane]
    __shared__ int next_col;

    const bool cold_start = (flags & 1) != 0;
    const int hist_mode = (flags & 2) ? ((flags & 4) ? 2 : 1) : 0;
    const bool no_dual = (flags & 8) != 0;
    const unsigned long long kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    const int MAXIT = 4 * k + 16;
    const int share = pool_elems / 32;  // pool elements per lane in the regular [e][lane] layout
    T* bcol = bS + lane;
    T* ucol = uS + lane;

    // this CTA's slice of the flat column range
    long long per = (total_cols + gridDim.x - 1) / gridDim.x;
    per = (per + 31) & ~31LL;
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
        // mop-up launch after k_nnls_reg (CTA-uniform): flag 16 = only the groups it skipped (singular Gram); with flag 64
        // additionally, in the other groups, only the columns it flagged in their warm-start mask (bit 63: system too big
        // for its register templates), continuing from the passive set it left there
        const bool only_deferred = (flags & 64) && *grp.ok != 0;
        if ((flags & 16) && *grp.ok != 0 && !only_deferred) continue;
        __syncthreads();  // every warp is done with the previous chunk (G / P / counter change)
        const int dual_ok = no_dual ? 0 : *grp.ok;
        for (int e = tid; e < KP * KP; e += blockDim.x) {
            const int a = e / KP, b = e % KP;
            Gs[a * MST + b] = (T)grp.gram[e];
            Ps[a * MST + b] = (T)grp.ginv[e];
        }
        if (tid == 0) next_col = 0;
        __syncthreads();
        const int ld = grp.ld;
        const T* __restrict__ rhs = reinterpret_cast<const T*>(grp.rhs) + (size_t)c0 * ld;
        T* __restrict__ xout = reinterpret_cast<T*>(grp.x) + (size_t)c0 * ld;
        unsigned long long* __restrict__ masks = grp.mask + c0;
        const T* __restrict__ upre = grp.u ? reinterpret_cast<const T*>(grp.u) + (size_t)c0 * ld : nullptr;

        bool active = false, exhausted = false;
        int col = 0;
        unsigned long long F = 0ull;
        int state = 0;
        T tol = (T)0;
        for (;;) {
            // ---- refill: idle lanes claim the next columns of the chunk ----
            const unsigned want = __ballot_sync(0xffffffffu, !active && !exhausted);
            if (want) {
                const int leader = __ffs((int)want) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&next_col, __popc(want));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!active && !exhausted) {
                    const int c = base + __popc(want & ((1u << lane) - 1u));
                    if (c >= ccols) {
                        exhausted = true;
                    } else if (only_deferred && !(masks[c] >> 63)) {
                        // solved by k_nnls_reg: not ours; the lane claims again on the next trip
                    } else {
                        col = c;
                        load_row<T, KP>(rhs + (size_t)c * ld, bcol, k, ld);
                        F = (cold_start && !only_deferred) ? 0ull : (masks[c] & kmask);
                        state = 3 | ((k + 1) << 8);
                        T bm = (T)0;
#pragma unroll
                        for (int f = 0; f < KP; ++f) bm = fmax(bm, fabs(bcol[f * 33]));
                        tol = bm * Eps<T>::v() * (T)8;
                        if (dual_ok && upre) {   // u = P b was formed on the tensor cores (k_pb_tc)
                            load_row<T, KP>(upre + (size_t)c * ld, ucol, k, ld);
                        } else if (dual_ok) {  // u = P b : register accumulators, broadcast shared-memory reads of P
                            T acc[KP];
#pragma unroll
                            for (int f = 0; f < KP; ++f) acc[f] = (T)0;
                            for (int t = 0; t < k; ++t) {
                                const T bt = bcol[t * 33];
                                const T* Pt = Ps + t * MST;
#pragma unroll
                                for (int f = 0; f < KP; ++f) acc[f] = fma(Pt[f], bt, acc[f]);
                            }
#pragma unroll
                            for (int f = 0; f < KP; ++f) ucol[f * 33] = acc[f];
                        }
                        active = true;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, active)) {
                if (__all_sync(0xffffffffu, exhausted)) break;
                continue;   // lanes that skipped columns claim again
            }
            // ---- one pivot round for every active lane ----
            int need = 0;
            if (active) {
                const int p = __popcll(F);
                const int s = (dual_ok && (k - p) < p) ? (k - p) : p;
                need = s * (s + 1) / 2 + s;
            }
            bool still = active;
            const bool regular = active && need <= share;
            if (regular)
                still = nnls_round<T, KP>(k, kmask, dual_ok, MAXIT, Gs, Ps, bcol, ucol, idxS + lane, pool + lane, 32, F, state,
                                          tol, cnt, pivots_local, hist_mode);
            // columns whose factor does not fit the per-lane share (only without the dual side, or k > 32):
            // a few lanes at a time over the whole warp pool
            unsigned big = __ballot_sync(0xffffffffu, active && !regular);
            while (big) {
                __syncwarp();
                int maxneed = ((big >> lane) & 1u) ? need : 0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) maxneed = max(maxneed, __shfl_xor_sync(0xffffffffu, maxneed, off));
                int fit = pool_elems / maxneed;
                if (fit < 1) fit = 1;
                if (fit > 32) fit = 32;
                const int rank = __popc(big & ((1u << lane) - 1u));
                const bool mine = ((big >> lane) & 1u) && rank < fit;
                if (mine)
                    still = nnls_round<T, KP>(k, kmask, dual_ok, MAXIT, Gs, Ps, bcol, ucol, idxS + lane, pool + rank, fit, F,
                                              state, tol, cnt, pivots_local, hist_mode);
                big &= ~__ballot_sync(0xffffffffu, mine);
            }
            __syncwarp();
            // ---- solved columns leave: the lane writes its own row (full 128-byte lines) and its passive set ----
            if (active && !still) {
                store_row<T, KP>(xout + (size_t)col * ld, bcol, k, ld);
                masks[col] = F;
            }
            active = still;
        }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

NnlsConfig nnls_config(int k, int kp, bool is_double, int num_sms) {
    NnlsConfig c{};
    const size_t ts = is_double ? 8 : 4;
    const int worst = k * (k + 1) / 2 + k;                 // primal-only fallback: s = k
    // per-lane share sized for systems of `cls` unknowns (larger ones run a few lanes at a time over the whole pool)
    int cls = k / 2 + 1;
    if (const char* e = getenv("LGR_NNLS_POOL_S")) cls = atoi(e);
    if (cls < 2) cls = 2;
    if (cls > k) cls = k;
    int pool = 32 * (cls * (cls + 1) / 2 + cls);
    if (pool < worst) pool = worst;
    int nw = 8;
    if (const char* e = getenv("LGR_NNLS_WARPS")) nw = atoi(e);
    if (nw < 1) nw = 1;
    if (nw > 8) nw = 8;
    const size_t limit = 226 * 1024;
    while (nw > 1 && nnls_layout(kp, nw, pool, ts).total > limit) --nw;
    while (nnls_layout(kp, nw, pool, ts).total > limit && pool > worst) pool -= 32;
    c.threads = nw * 32;
    c.pool_elems = pool;
    c.smem = nnls_layout(kp, nw, pool, ts).total;
    int per_sm = (int)(limit / (c.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * c.threads > 512) per_sm = 512 / c.threads > 0 ? 512 / c.threads : 1;   // register budget
    c.grid_cap = num_sms * per_sm;
    c.ctas_per_sm = per_sm;
    c.ct_cols = 1024;
    return c;
}

void launch_nnls(bool is_double, const NnlsGroup* groups, int ngroups, long long total_cols, int k, int kp,
                 const NnlsConfig& cfg, Counters* cnt, int flags, cudaStream_t st) {
    if (total_cols <= 0) return;
    // small launches (the V_i / W solves): fewer warps per CTA so that the columns spread over all SMs
    const int num_sms = cfg.grid_cap / (cfg.ctas_per_sm > 0 ? cfg.ctas_per_sm : 1);
    int nw = cfg.threads / 32;
    while (nw > 1 && total_cols < (long long)num_sms * nw * 32) nw >>= 1;
    const int threads = nw * 32;
    const size_t smem = nnls_layout(kp, nw, cfg.pool_elems, is_double ? 8 : 4).total;
    long long want = (total_cols + threads - 1) / threads;   // at least one column per lane
    int grid = (int)(want < cfg.grid_cap ? want : cfg.grid_cap);
    // mop-up launches (flag 16: only what the fast kernels left behind, normally nothing) pay for their empty pass: one CTA
    // per SM at most, so that the pass is a single wave of CTAs that exit at once
    if ((flags & 16) && grid > num_sms) grid = num_sms;
    if (is_double) {
        if (kp == 32)
            k_nnls<double, 32><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cfg.pool_elems, cnt, flags);
        else
            k_nnls<double, 64><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cfg.pool_elems, cnt, flags);
    } else {
        if (kp == 32)
            k_nnls<float, 32><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cfg.pool_elems, cnt, flags);
        else
            k_nnls<float, 64><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cfg.pool_elems, cnt, flags);
    }
    LGR_CUDA(cudaGetLastError());
}

void nnls_init() {
    const int maxs = 227 * 1024 - 1024;
    LGR_CUDA(cudaFuncSetAttribute(k_nnls<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls<float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls<double, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls<double, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
}

}  // namespace lgr
