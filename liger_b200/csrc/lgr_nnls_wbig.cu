// liger_b200 -- warp-per-column fp32 NNLS for the H-solve columns whose KKT system is too big for the register templates
// of lgr_nnls_reg.cu (more than 16 unknowns; only possible for 32 < k <= 62, while |F| is near k/2, i.e. during the first
// ANLS iterations of a k = 50 run).  k_nnls_reg flags those columns in their warm-start mask (bit 63) and this kernel
// continues them from the passive set they had reached.  Same problem and exchange rule as everywhere else (reference:
// RcppPlanc::bppnnls_prod, R/cINMF.R:481).
//
// A warp owns one column: lane f holds factors f and f + 32 (b, u, x in two registers each); the passive set is a
// warp-uniform 64-bit mask.  Each round solves from the smaller side (s = min(|F|, k - |F|) <= 31 unknowns): lane t holds
// row t of the s x s block and a Gauss-Jordan sweep moves the pivot row by shuffle; the k-vector products read matrix
// rows f and f + 32 (stride 65 words: conflict free) while the solution travels by shuffle.
#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int BKP = 64;
constexpr int BMST = 65;
constexpr unsigned long long BDEFER = 1ull << 63;
template <typename T> __device__ __forceinline__ T beps();
template <> __device__ __forceinline__ float beps<float>() { return 1.1920929e-7f; }
template <> __device__ __forceinline__ double beps<double>() { return 2.220446049250313e-16; }
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double tmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float tabs(float a) { return fabsf(a); }
__device__ __forceinline__ double tabs(double a) { return fabs(a); }
__device__ __forceinline__ float tfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double tfma(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
__device__ __forceinline__ T wmaxf(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = tmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
// index of the (t+1)-th set bit of a 64-bit mask (t < popcount)
__device__ __forceinline__ int nth_set64(unsigned long long S, int t) {
    const uint32_t lo = (uint32_t)S, hi = (uint32_t)(S >> 32);
    const int nlo = __popc(lo);
    return t < nlo ? (int)__fns(lo, 0, t + 1) : 32 + (int)__fns(hi, 0, t - nlo + 1);
}
// value of factor `idx` of a vector held as (v0: factors 0..31, v1: factors 32..63) across the warp
template <typename T>
__device__ __forceinline__ T pick(T v0, T v1, int idx) {
    const T a = __shfl_sync(0xffffffffu, v0, idx & 31), b = __shfl_sync(0xffffffffu, v1, idx & 31);
    return idx < 32 ? a : b;
}

// M_SS z = r_S, s <= W unknowns; lane t < s returns z_t (0 for pinned pivots)
template <int W, typename T>
__device__ __forceinline__ T solve_big(const T* __restrict__ M, const unsigned long long S, const int s, const T r0, const T r1,
                                       const int lane, bool& ok) {
    const int my = (lane < s) ? nth_set64(S, lane) : 0;
    T a[W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = (T)0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j >= s) break;   // warp-uniform
        const int gj = __shfl_sync(0xffffffffu, my, j);
        a[j] = M[my * BMST + gj];
    }
    T r = pick(r0, r1, my);
    T myinv = (T)0;
    ok = true;
    // Gauss-Jordan with a ROTATING register window: after step j the surviving columns shift down by one, so the pivot
    // column is always a[0] and the step body has static register indices inside a dynamic loop (a fully unrolled
    // 32 x 32 sweep is ~48 KB of code and starved the instruction cache: measured 2x slower)
    for (int j = 0; j < s; ++j) {
        const T pj = __shfl_sync(0xffffffffu, a[0], j);
        const T orig = M[(int)__shfl_sync(0xffffffffu, my, j) * (BMST + 1)];
        T inv = (T)0;
        if (pj > orig * (beps<T>() * (T)4))
            inv = (T)1 / pj;
        else
            ok = false;
        if (lane == j) myinv = inv;
        const T f = (lane == j) ? (T)0 : a[0] * inv;   // multiplier of the pivot row for my row
#pragma unroll
        for (int c = 1; c < W; ++c) {
            const T rc = __shfl_sync(0xffffffffu, a[c], j);
            a[c - 1] = tfma(-f, rc, a[c]);
        }
        a[W - 1] = (T)0;
        const T rr = __shfl_sync(0xffffffffu, r, j);
        r = tfma(-f, rr, r);
    }
    return (lane < s) ? r * myinv : (T)0;
}

// ALL = false: only the columns flagged by k_nnls_reg<64> (fp32 H solves).  ALL = true: every column of the launch, from
// its warm-start set or cold (flags & 1) -- the FP64 V_i / U_i / W solves at k > 32.
template <typename T, bool ALL>
__global__ void __launch_bounds__(256) k_nnls_wbig(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols, int k,
                                                   Counters* cnt, int flags) {
    extern __shared__ __align__(16) unsigned char bsm_raw[];
    T* Gs = reinterpret_cast<T*>(bsm_raw);
    T* Ps = Gs + BKP * BMST;
    const bool cold_start = ALL && (flags & 1) != 0;
    __shared__ int next_col;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long kmask = (1ull << k) - 1ull;   // k <= 62
    const unsigned long long bit0 = 1ull << lane, bit1 = 1ull << (lane + 32);
    const int MAXIT = 4 * k + 16;

    long long per = (total_cols + gridDim.x - 1) / gridDim.x;
    if (!ALL) per = (per + 31) & ~31LL;
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
        if (*grp.ok == 0) continue;   // singular Gram: k_nnls solves the whole group
        __syncthreads();
        for (int e = tid; e < BKP * BKP; e += blockDim.x) {
            const int a = e / BKP, b = e % BKP;
            Gs[a * BMST + b] = (T)grp.gram[e];
            Ps[a * BMST + b] = (T)grp.ginv[e];
        }
        if (tid == 0) next_col = 0;
        __syncthreads();
        const T* __restrict__ rhs = reinterpret_cast<const T*>(grp.rhs) + (size_t)c0 * BKP;
        T* __restrict__ xout = reinterpret_cast<T*>(grp.x) + (size_t)c0 * BKP;
        unsigned long long* __restrict__ masks = grp.mask + c0;
        for (;;) {
            // a warp claims 32 columns at a time and looks for the flagged ones among them
            // (ALL: one column per claim -- a small batch must spread over as many warps as possible)
            constexpr int CLAIM = ALL ? 1 : 32;
            int base = 0;
            if (lane == 0) base = atomicAdd(&next_col, CLAIM);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= ccols) break;
            const unsigned long long mine = (lane < CLAIM && base + lane < ccols) ? masks[base + lane] : 0ull;
            unsigned todo = __ballot_sync(0xffffffffu, ALL ? (lane == 0) : ((mine & BDEFER) != 0ull));
            while (todo) {
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const int c = base + src;
                unsigned long long F = cold_start ? 0ull : (__shfl_sync(0xffffffffu, mine, src) & kmask);
                const T b0 = rhs[(size_t)c * BKP + lane], b1 = rhs[(size_t)c * BKP + 32 + lane];
                const T tol = wmaxf(tmax(tabs(b0), tabs(b1))) * (beps<T>() * (T)8);
                T u0 = (T)0, u1 = (T)0, x0 = (T)0, x1 = (T)0;
                bool have_u = false;
                int alpha = 3, beta = k + 1, iter = 0;
                for (;;) {
                    const int p = __popcll(F);
                    const bool dual = (k - p) < p;
                    const unsigned long long S = dual ? ((~F) & kmask) : F;
                    const int s = dual ? (k - p) : p;
                    if (dual && !have_u) {   // u = P b
                        T a0 = (T)0, a1 = (T)0;
                        for (int t = 0; t < k; ++t) {
                            const T bt = pick(b0, b1, t);
                            a0 = tfma(Ps[lane * BMST + t], bt, a0);
                            a1 = tfma(Ps[(lane + 32) * BMST + t], bt, a1);
                        }
                        u0 = a0;
                        u1 = a1;
                        have_u = true;
                    }
                    const T* M = dual ? Ps : Gs;
                    bool ok = true;
                    const T rr0 = dual ? u0 : b0, rr1 = dual ? u1 : b1;
                    const T z = s <= 16   ? solve_big<16, T>(M, S, s, rr0, rr1, lane, ok)
                                : s <= 24 ? solve_big<24, T>(M, S, s, rr0, rr1, lane, ok)
                                          : solve_big<32, T>(M, S, s, rr0, rr1, lane, ok);
                    if (!ok && lane == 0) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
                    const uint32_t Slo = (uint32_t)S, Shi = (uint32_t)(S >> 32);
                    const int nlo = __popc(Slo);
                    const T zt0 = __shfl_sync(0xffffffffu, z, __popc(Slo & ((1u << lane) - 1u)));
                    const T zt1 = __shfl_sync(0xffffffffu, z, (nlo + __popc(Shi & ((1u << lane) - 1u))) & 31);
                    const T zf0 = (S & bit0) ? zt0 : (T)0, zf1 = (S & bit1) ? zt1 : (T)0;
                    T a0 = (T)0, a1 = (T)0;
                    {
                        unsigned long long rem = S;
                        int t = 0;
                        while (rem) {   // warp-uniform
                            const int gi = __ffsll((long long)rem) - 1;
                            rem &= rem - 1;
                            const T zt = __shfl_sync(0xffffffffu, z, t);
                            a0 = tfma(M[lane * BMST + gi], zt, a0);
                            a1 = tfma(M[(lane + 32) * BMST + gi], zt, a1);
                            ++t;
                        }
                    }
                    unsigned long long infeas;
                    if (dual) {
                        x0 = u0 - a0;
                        x1 = u1 - a1;
                        const T tolx = wmaxf(tmax((F & bit0) ? tabs(x0) : (T)0, (F & bit1) ? tabs(x1) : (T)0)) * (beps<T>() * (T)8);
                        const uint32_t ilo = __ballot_sync(0xffffffffu, ((S & bit0) && zf0 > tol) || ((F & bit0) && x0 < -tolx));
                        const uint32_t ihi = __ballot_sync(0xffffffffu, ((S & bit1) && zf1 > tol) || ((F & bit1) && x1 < -tolx));
                        infeas = (unsigned long long)ilo | ((unsigned long long)ihi << 32);
                    } else {
                        const T y0 = a0 - b0, y1 = a1 - b1;
                        const T tolx = wmaxf(tmax(tabs(zf0), tabs(zf1))) * (beps<T>() * (T)8);
                        const unsigned long long A = (~F) & kmask;
                        const uint32_t ilo = __ballot_sync(0xffffffffu, ((S & bit0) && zf0 < -tolx) || ((A & bit0) && y0 < -tol));
                        const uint32_t ihi = __ballot_sync(0xffffffffu, ((S & bit1) && zf1 < -tolx) || ((A & bit1) && y1 < -tol));
                        infeas = (unsigned long long)ilo | ((unsigned long long)ihi << 32);
                        x0 = zf0;
                        x1 = zf1;
                    }
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
                        ++iter;
                        if (lane == 0) ++pivots_local;
                        continue;
                    }
                    if (infeas != 0ull && lane == 0) atomicAdd(&cnt->nnls_unconverged, 1ull);
                    break;
                }
                xout[(size_t)c * BKP + lane] = ((F & bit0) && x0 > (T)0) ? x0 : (T)0;
                xout[(size_t)c * BKP + 32 + lane] = ((F & bit1) && x1 > (T)0) ? x1 : (T)0;
                if (lane == 0) masks[c] = F;   // clears the flag
            }
        }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

}  // namespace

template <typename T, bool ALL>
static void launch_wbig_t(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, int flags,
                          cudaStream_t st) {
    if (total_cols <= 0) return;
    const size_t smem = (size_t)2 * BKP * BMST * sizeof(T);
    long long want = (total_cols + 255) / 256;
    if (ALL) want = (total_cols + 7) / 8;           // small batches: one column per warp when there are that few
    const long long cap = (long long)num_sms * (sizeof(T) == 4 ? 4 : 3);
    const int grid = (int)(want < cap ? want : cap);
    k_nnls_wbig<T, ALL><<<grid, 256, smem, st>>>(groups, ngroups, total_cols, k, cnt, flags);
    LGR_CUDA(cudaGetLastError());
}

// continues the columns k_nnls_reg<64> flagged (bit 63 of their warm-start mask); all other columns are left alone
void launch_nnls_wbig(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt,
                      cudaStream_t st) {
    launch_wbig_t<float, false>(groups, ngroups, total_cols, k, num_sms, cnt, 0, st);
}

// FP64, 32 < k <= 62, every column of a small batch (V_i / U_i / W solves); singular groups are skipped (flag 16 mop-up)
bool nnls_wbig64_supported(int k, int kp, long long total_cols, int num_sms) {
    return kp == 64 && k <= 62 && total_cols <= (long long)num_sms * 16 * 64;
}
void launch_nnls_wbig64(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, int flags,
                        cudaStream_t st) {
    launch_wbig_t<double, true>(groups, ngroups, total_cols, k, num_sms, cnt, flags, st);
}


// opt-in shared-memory sizes (per device; called from kernels_init)
void nnls_wbig_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_wbig<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((size_t)2 * BKP * BMST * sizeof(float))));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_wbig<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((size_t)2 * BKP * BMST * sizeof(double))));
}

}  // namespace lgr
