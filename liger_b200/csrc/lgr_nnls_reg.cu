// liger_b200 -- register-resident block-principal-pivoting NNLS for the H solves (fp32, k <= 32).
// Same problem and pivoting rule as lgr_nnls.cu (reference: RcppPlanc::bppnnls_prod, call site R/cINMF.R:481);
// different execution plan:
//
//   * one thread per column, a lane refills itself as soon as its column is solved (no barrier between columns);
//   * every pivot round solves the KKT system from the smaller side (primal G_FF or dual P_AA, P = G^-1 from
//     k_gram_inv), so the per-column system has s = min(|F|, k-|F|) <= 16 unknowns;
//   * the s x s principal block is gathered into REGISTERS, padded to a compile-time size SMAX in {8,12,16} with
//     identity rows (index 32 addresses a padding row/column of the staged matrix whose only non-zero is a unit
//     diagonal), and factorised by a fully unrolled Cholesky: no shared-memory traffic, no data-dependent loop
//     bounds, no divergence inside a warp.  SMAX is chosen per warp and per round from the largest s of its lanes;
//   * the k-vector products (x = u - P_:A y_A, y = G_:F x_F - b) are dense 32x32 products against the solution
//     expanded by factor id: the matrix rows are then warp-uniform, i.e. broadcast 128-bit shared-memory reads
//     (one wavefront per four FMAs' worth of operands) instead of per-lane row gathers (bank conflicts);
//   * u = P b is formed lazily, only for columns that actually take the dual side.
//
// Groups whose Gram is numerically singular (no inverse: *ok == 0) are left to k_nnls (lgr_nnls.cu), which is
// launched right after this kernel restricted to exactly those groups.
#include <type_traits>

#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int RSUB = 4096;       // columns ordered (by system size) and solved per pass of a CTA (KP = 32)
constexpr float REPS = 1.1920929e-7f;
constexpr unsigned long long DEFER_BIT = 1ull << 63;   // warm-start mask flag: "left to k_nnls" (needs k <= 62)

// Sizes for a padded factor count KP (32 or 64): the staged matrices have KP + 1 rows (row KP = padding row whose only
// non-zero is a unit diagonal) of stride KP + 4 words; the b / u staging tiles are [16-byte chunk][lane] with KP/4
// chunks + a zero chunk (rhs of the padding unknowns); the z tile is [row][lane] with KP + 1 rows.
template <int KP>
struct RT {
    static constexpr int MST = KP + 4;
    static constexpr int ROWS = KP + 1;
    static constexpr int COL = (KP + 1) * 32;
    static constexpr int BUF = (KP / 4 + 1) * 32 * 4;
    static constexpr int SUB = KP == 32 ? RSUB : 1024;
    using Mask = typename std::conditional<KP == 32, uint32_t, unsigned long long>::type;
};
__device__ __forceinline__ int popc_m(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popc_m(unsigned long long x) { return __popcll(x); }
__device__ __forceinline__ int ffs_m(uint32_t x) { return __ffs((int)x); }
__device__ __forceinline__ int ffs_m(unsigned long long x) { return __ffsll((long long)x); }
__device__ __forceinline__ int top_m(uint32_t x) { return 31 - __clz((int)x); }
__device__ __forceinline__ int top_m(unsigned long long x) { return 63 - __clzll((long long)x); }

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
// word offset of element f of a lane's row in a [chunk][lane] staging tile (relative to the lane's chunk 0)
__device__ __forceinline__ int bw(int f) { return (f >> 2) * 128 + (f & 3); }
// packed fp32 x2 FMA (Blackwell): d.lo = a.lo * b.lo + c.lo, d.hi = a.hi * b.hi + c.hi in one instruction
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b) {
    unsigned long long d, a, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RegSmem {
    size_t G, P, fmask, warp0, per_warp, total;
};
template <int KP>
__host__ __device__ inline RegSmem reg_layout(int nwarps) {
    RegSmem L;
    L.G = 0;
    L.P = (size_t)RT<KP>::ROWS * RT<KP>::MST * 4;
    L.fmask = 2 * L.P;
    L.warp0 = L.fmask + (size_t)RT<KP>::SUB * sizeof(typename RT<KP>::Mask);
    L.per_warp = (size_t)(4 * RT<KP>::BUF + RT<KP>::COL) * 4;   // 2 x (b, u) staging tiles (double buffered), z tile
    L.total = L.warp0 + (size_t)nwarps * L.per_warp;
    return L;
}

// One pivot round of one column with the principal block padded to SMAX.  Returns true while the column is active.
template <int SMAX, int KP>
__device__ __forceinline__ bool round_reg(const int k, const typename RT<KP>::Mask kmask, const bool dual, const int s,
                                          const int smax, const typename RT<KP>::Mask S,
                                          const float* __restrict__ Mat, const float* bcol, const float* ucol, float* zcol,
                                          typename RT<KP>::Mask& F, int& state, const float tol, const int MAXIT, Counters* cnt,
                                          unsigned long long& pivots, float4* __restrict__ xrow) {
    using Mask = typename RT<KP>::Mask;
    constexpr int RKP = KP, RMST = RT<KP>::MST;
    constexpr Mask ONE = 1;
    constexpr int N = SMAX * (SMAX + 1) / 2;
    int idx[SMAX];
    {
        Mask r = S;
#pragma unroll
        for (int t = 0; t < SMAX; ++t) {
            idx[t] = r ? (ffs_m(r) - 1) : RKP;
            r &= r - 1;
        }
    }
    float A[N];
#pragma unroll
    for (int i = 0; i < SMAX; ++i) {
        const float* rowp = Mat + idx[i] * RMST;
#pragma unroll
        for (int j = 0; j < SMAX; ++j)
            if (j <= i) A[tri(i, j)] = rowp[idx[j]];
    }
    // left-looking Cholesky, factor in place, diagonal holds 1/L_jj; a collapsed pivot pins the coordinate to zero
    bool ok = true;
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
        const float orig = A[tri(j, j)];
        float d = orig;
#pragma unroll
        for (int t = 0; t < SMAX; ++t)
            if (t < j) d = fmaf(-A[tri(j, t)], A[tri(j, t)], d);
        float invd = 0.f;
        if (d > orig * (REPS * 4.f))
            invd = rsqrtf(d);
        else
            ok = false;
        A[tri(j, j)] = invd;
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
            if (i > j) {
                float a = A[tri(i, j)];
#pragma unroll
                for (int t = 0; t < SMAX; ++t)
                    if (t < j) a = fmaf(-A[tri(i, t)], A[tri(j, t)], a);
                A[tri(i, j)] = a * invd;
            }
        }
    }
    if (!ok) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
    const float* rhs = dual ? ucol : bcol;
    float z[SMAX];
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
        float v = rhs[bw(idx[j])];
#pragma unroll
        for (int t = 0; t < SMAX; ++t)
            if (t < j) v = fmaf(-A[tri(j, t)], z[t], v);
        z[j] = v * A[tri(j, j)];
    }
#pragma unroll
    for (int j = SMAX - 1; j >= 0; --j) {
        float v = z[j];
#pragma unroll
        for (int i = 0; i < SMAX; ++i)
            if (i > j) v = fmaf(-A[tri(i, j)], z[i], v);
        z[j] = v * A[tri(j, j)];
    }
    // infeasible coordinates of the solved side; expand the (signed) solution by factor id
    Mask infeas = 0;
    float zmax = 0.f;
#pragma unroll
    for (int t = 0; t < SMAX; ++t) zmax = fmaxf(zmax, fabsf(z[t]));
    const float tolz = dual ? tol : zmax * (REPS * 8.f);
    const float sgn = dual ? -1.f : 1.f;   // dual: y_A = -z must be >= 0 ; primal: x_F = z must be >= 0
#pragma unroll
    for (int g = 0; g < RKP; ++g) zcol[g * 32] = 0.f;
#pragma unroll
    for (int t = 0; t < SMAX; ++t) {
        if (t < s) {
            if (sgn * z[t] < -tolz) infeas |= ONE << idx[t];
            zcol[idx[t] * 32] = sgn * z[t];
        }
    }
    // dense product against the expanded vector: acc = u - P zs (dual, zs = -(-z) ... see sign below) or G zs - b
    float acc[RKP];
    {
        const float4* src = reinterpret_cast<const float4*>(rhs);
        const float sg0 = dual ? 1.f : -1.f;
#pragma unroll
        for (int c = 0; c < RKP / 4; ++c) {
            const float4 v = src[c * 32];
            acc[4 * c + 0] = sg0 * v.x;
            acc[4 * c + 1] = sg0 * v.y;
            acc[4 * c + 2] = sg0 * v.z;
            acc[4 * c + 3] = sg0 * v.w;
        }
    }
    // dual:   x = u - P_:A z = u + P_:A (sgn z)      primal: y = G_:F z - b = -b + G_:F (sgn z)
    if (SMAX <= 8) {
        // few unknowns: gather the s rows of the matrix (per-lane rows, some bank conflicts, but only s of them)
#pragma unroll
        for (int t = 0; t < SMAX; ++t) {
            if (t < smax) {   // warp-uniform
                const float zt = sgn * z[t];
                const float4* row = reinterpret_cast<const float4*>(Mat + idx[t] * RMST);
#pragma unroll
                for (int c = 0; c < RKP / 4; ++c) {
                    const float4 mrow = row[c];
                    ffma2(acc[4 * c + 0], acc[4 * c + 1], mrow.x, mrow.y, zt);
                    ffma2(acc[4 * c + 2], acc[4 * c + 3], mrow.z, mrow.w, zt);
                }
            }
        }
    } else {
        for (int g = 0; g < k; ++g) {
            const float zg = zcol[g * 32];
            const float4* row = reinterpret_cast<const float4*>(Mat + g * RMST);
#pragma unroll
            for (int c = 0; c < RKP / 4; ++c) {
                const float4 mrow = row[c];
                ffma2(acc[4 * c + 0], acc[4 * c + 1], mrow.x, mrow.y, zg);
                ffma2(acc[4 * c + 2], acc[4 * c + 3], mrow.z, mrow.w, zg);
            }
        }
    }
    // sign tests over all 32 coordinates with four independent chains, masked by the set afterwards.  In the dual case
    // acc is x on every coordinate (x_A = u_A - P_AA z = 0 up to rounding), so the unmasked maximum scales the tolerance.
    {
        float thr = tol;
        if (dual) {
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
            for (int f = 0; f < RKP; f += 4) {
                m0 = fmaxf(m0, fabsf(acc[f]));
                m1 = fmaxf(m1, fabsf(acc[f + 1]));
                m2 = fmaxf(m2, fabsf(acc[f + 2]));
                m3 = fmaxf(m3, fabsf(acc[f + 3]));
            }
            thr = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * (REPS * 8.f);
        }
        Mask n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
        for (int f = 0; f < RKP; f += 4) {
            n0 |= (acc[f] < -thr) ? (ONE << f) : (Mask)0;
            n1 |= (acc[f + 1] < -thr) ? (ONE << (f + 1)) : (Mask)0;
            n2 |= (acc[f + 2] < -thr) ? (ONE << (f + 2)) : (Mask)0;
            n3 |= (acc[f + 3] < -thr) ? (ONE << (f + 3)) : (Mask)0;
        }
        infeas |= ((n0 | n1) | (n2 | n3)) & (dual ? F : ((~F) & kmask));
    }
    int alpha = state & 0xff, beta = (state >> 8) & 0xff, iter = state >> 16;
    if (infeas != 0 && iter < MAXIT) {
        const int ninf = popc_m(infeas);
        Mask flip;
        if (ninf < beta) {
            beta = ninf;
            alpha = 3;
            flip = infeas;
        } else if (alpha >= 1) {
            --alpha;
            flip = infeas;
        } else {
            flip = ONE << top_m(infeas);
        }
        F ^= flip;
        state = alpha | (beta << 8) | ((iter + 1) << 16);
        ++pivots;
        return true;
    }
    if (infeas != 0) atomicAdd(&cnt->nnls_unconverged, 1ull);
    // final x straight from registers: full 128-byte row
    if (dual) {
#pragma unroll
        for (int f = 0; f < RKP; ++f)
            if (!((F >> f) & ONE) || !(acc[f] > 0.f)) acc[f] = 0.f;
    } else {
#pragma unroll
        for (int f = 0; f < RKP; ++f) acc[f] = fmaxf(zcol[f * 32], 0.f);
    }
#pragma unroll
    for (int c = 0; c < RKP / 4; ++c) xrow[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    return false;
}

template <int KP>
__global__ void __launch_bounds__(256, 1) k_nnls_reg(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols,
                                                     int k, Counters* cnt, int flags) {
    using Mask = typename RT<KP>::Mask;
    constexpr int RKP = KP, RMST = RT<KP>::MST, RROWS = RT<KP>::ROWS, RBUF = RT<KP>::BUF, RSUBK = RT<KP>::SUB;
    constexpr Mask ONE = 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NW = blockDim.x >> 5;
    const RegSmem L = reg_layout<KP>(NW);
    float* Gs = reinterpret_cast<float*>(smem_raw + L.G);
    float* Ps = reinterpret_cast<float*>(smem_raw + L.P);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* wbase = reinterpret_cast<float*>(smem_raw + L.warp0 + (size_t)warp * L.per_warp);
    // staging tiles: buffer q of b at wbase + q*2*RBUF, of u at wbase + (q*2+1)*RBUF; this lane's chunk 0 at + lane*4
    float* stage = wbase + lane * 4;
    float* zcol = wbase + 4 * RBUF + lane;
    __shared__ int next_col;
    __shared__ int hist[36];   // bins (side, system size): 17 primal, 17 dual; [34] = systems above 8 unknowns
    __shared__ unsigned short order[RSUBK];
    Mask* fmask = reinterpret_cast<Mask*>(smem_raw + L.fmask);   // warm-start passive sets of the pass

    const bool cold_start = (flags & 1) != 0;
    const bool defer_all = KP > 32 && (flags & 128) != 0;   // developer knob LGR_NNLS_PAIR_ALL: every column goes to the continuation kernel
    const bool sync_rounds = (flags & 32) != 0;   // CTA-wide round barrier: the warps walk the unrolled code together (I-cache)
    const Mask kmask = (k >= (int)(8 * sizeof(Mask))) ? ~(Mask)0 : ((ONE << k) - ONE);
    const int MAXIT = 4 * k + 16;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(stage + q * RBUF + (RKP / 4) * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
    zcol[RKP * 32] = 0.f;

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
        if (*grp.ok == 0) continue;   // singular Gram: k_nnls takes this group (CTA-uniform)
        __syncthreads();
        for (int e = tid; e < RROWS * RMST; e += blockDim.x) {
            const int a = e / RMST, b = e % RMST;
            float gv = 0.f, pv = 0.f;
            if (a < RKP && b < RKP) {
                gv = (float)grp.gram[a * RKP + b];
                pv = (float)grp.ginv[a * RKP + b];
            } else if (a == RKP && b == RKP) {
                gv = 1.f;
                pv = 1.f;
            }
            Gs[e] = gv;
            Ps[e] = pv;
        }
        __syncthreads();
        const float* __restrict__ rhs_g = reinterpret_cast<const float*>(grp.rhs) + (size_t)c0 * RKP;
        float* __restrict__ xout_g = reinterpret_cast<float*>(grp.x) + (size_t)c0 * RKP;
        unsigned long long* __restrict__ masks_g = grp.mask + c0;
        const float* __restrict__ upre_g = grp.u ? reinterpret_cast<const float*>(grp.u) + (size_t)c0 * RKP : nullptr;
      for (int sub0 = 0; sub0 < ccols; sub0 += RSUBK) {
        // ---- order this sub-chunk's columns by the side and the size of the system their warm start implies (primal
        //      systems first, largest first): the lanes of a warp then solve systems of similar size (the small templates
        //      are actually used) from the same matrix (G or P: the 128-bit row reads of the dense products are then
        //      one broadcast per warp instead of two) ----
        const int nsub = min(RSUBK, ccols - sub0);
        const float* __restrict__ rhs = rhs_g + (size_t)sub0 * RKP;
        float* __restrict__ xout = xout_g + (size_t)sub0 * RKP;
        unsigned long long* __restrict__ masks = masks_g + sub0;
        const float* __restrict__ upre = upre_g ? upre_g + (size_t)sub0 * RKP : nullptr;
        __syncthreads();   // previous sub-chunk fully drained (order[] / counters are reused)
        for (int e = tid; e < 36; e += blockDim.x) hist[e] = 0;   // (a small batch runs with a single warp)
        __syncthreads();
        const bool by_side = (flags & 64) == 0;   // (developer knob LGR_NNLS_REG_NOSIDE: order by size only)
        auto bin_of = [&](int p) { return ((by_side && (k - p) < p) ? 17 : 0) + 16 - min(min(p, k - p), 16); };
        for (int c = tid; c < nsub; c += blockDim.x) {
            const Mask f0 = cold_start ? (Mask)0 : ((Mask)masks[c] & kmask);
            fmask[c] = f0;
            const int p0 = popc_m(f0);
            atomicAdd(&hist[bin_of(p0)], 1);
            if (KP > 32 && min(p0, k - p0) > 16) atomicAdd(&hist[35], 1);   // systems this kernel hands over anyway
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0, big = 0;
            for (int b = 0; b < 34; ++b) {
                const int h = hist[b];
                if (b % 17 < 8) big += h;   // systems of more than 8 unknowns
                hist[b] = run;
                run += h;
            }
            hist[34] = big;
            next_col = 0;
        }
        __syncthreads();
        for (int c = tid; c < nsub; c += blockDim.x) order[atomicAdd(&hist[bin_of(popc_m(fmask[c]))], 1)] = (unsigned short)c;
        __syncthreads();

        // round barrier only while the big unrolled templates dominate (they do not fit the instruction cache: the
        // warps then walk the code together); a pass of mostly small systems runs barrier-free
        const bool sync_pass = sync_rounds && hist[34] * 4 > nsub;
        // k > 32: when most columns of the pass go to the continuation kernel anyway, all of them do -- it runs eight warps
        // per SM against this kernel's four (measured on a C4 shard: -0.2 to -0.3 ms per iteration while the systems are big,
        // +0.3 to +0.6 ms if applied when they are small)
        const bool defer_pass = defer_all || (KP > 32 && (flags & 256) == 0 && hist[35] * 2 > nsub);
        bool active = false, exhausted = false, have_u = false;
        int col = 0, state = 0, nxt = -1;
        Mask F = 0;
        float tol = 0.f;
        // every lane keeps ONE column reserved ahead of the one it is solving, and prefetches that column's b (and u)
        // row -- exactly one 128-byte line each -- into L1 while it works: the refill then finds its operands on chip
        int cur = 0;   // staging buffer of the column being solved; the reserved column streams into cur ^ 1
        auto claim = [&](bool wants, int into) {
            const unsigned w = __ballot_sync(0xffffffffu, wants);
            int got = -1;
            if (w) {
                const int leader = __ffs((int)w) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&next_col, __popc(w));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (wants) {
                    const int ci = base + __popc(w & ((1u << lane) - 1u));
                    if (ci < nsub) {
                        got = order[ci];
                        const uint32_t db = smem_addr(stage + into * 2 * RBUF);
                        const float* gb = rhs + (size_t)got * RKP;
#pragma unroll
                        for (int i = 0; i < RKP / 4; ++i)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(db + i * 512), "l"(gb + 4 * i) : "memory");
                        if (upre) {
                            const float* gu = upre + (size_t)got * RKP;
#pragma unroll
                            for (int i = 0; i < RKP / 4; ++i)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(db + RBUF * 4 + i * 512), "l"(gu + 4 * i)
                                             : "memory");
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return got;
        };
        nxt = claim(true, 0);
        cur = 1;   // the first refill flips to buffer 0
        for (;;) {
            const bool refill = !active && !exhausted;
            int c = -1;
            if (refill) {
                c = nxt;
                if (c < 0) exhausted = true;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");   // the reserved column's rows have landed (issued a round ago)
            if (c >= 0) cur ^= 1;
            {
                const bool wants = refill && c >= 0;
                const int got = claim(wants, cur ^ 1);
                if (wants) nxt = got;
            }
            const float* bcol = stage + cur * 2 * RBUF;
            float* ucol = stage + cur * 2 * RBUF + RBUF;
            if (c >= 0) {
                col = c;
                const float4* b4 = reinterpret_cast<const float4*>(bcol);
                float bm = 0.f;
#pragma unroll
                for (int i = 0; i < RKP / 4; ++i) {
                    const float4 v = b4[i * 32];
                    bm = fmaxf(bm, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                }
                F = fmask[c];
                state = 3 | ((k + 1) << 8);
                tol = bm * (REPS * 8.f);
                have_u = upre != nullptr;   // u = P b from the tensor cores arrived with b
                active = true;
            }
            if (sync_pass ? !__syncthreads_or(active) : !__any_sync(0xffffffffu, active)) break;
            int s = 0;
            bool dual = false;
            Mask S = 0;
            if (active) {
                const int p = popc_m(F);
                dual = (k - p) < p;
                S = dual ? ((~F) & kmask) : F;
                s = dual ? (k - p) : p;
                // u = P b, once per column, broadcast reads of P.  A fresh column without a passive set (first iteration, cold start)
                // every fresh column takes it first and starts from F = {u > 0}, the support of the clipped unconstrained
                // solution, instead of from the empty set (which reaches the same point two rounds later)
                const bool seed = state == (3 | ((k + 1) << 8)) && F == 0 && !have_u;
                // (a column that is about to be handed over -- more than 16 unknowns, k > 32 -- does not need u here: the
                //  continuation kernel forms its own)
                const bool hand_over = KP > 32 && (s > 16 || defer_pass) && !seed;
                if ((dual || seed) && !have_u && !hand_over) {
                    float acc[RKP];
#pragma unroll
                    for (int f = 0; f < RKP; ++f) acc[f] = 0.f;
                    for (int t = 0; t < k; ++t) {
                        const float bt = bcol[bw(t)];
                        const float4* row = reinterpret_cast<const float4*>(Ps + t * RMST);
#pragma unroll
                        for (int c = 0; c < RKP / 4; ++c) {
                            const float4 mrow = row[c];
                            ffma2(acc[4 * c + 0], acc[4 * c + 1], mrow.x, mrow.y, bt);
                            ffma2(acc[4 * c + 2], acc[4 * c + 3], mrow.z, mrow.w, bt);
                        }
                    }
#pragma unroll
                    for (int c4 = 0; c4 < RKP / 4; ++c4)
                        reinterpret_cast<float4*>(ucol)[c4 * 32] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
                    have_u = true;
                    if (seed) {
                        Mask f0 = 0;
#pragma unroll
                        for (int f = 0; f < RKP; ++f) f0 |= (acc[f] > 0.f) ? (ONE << f) : (Mask)0;
                        F = f0 & kmask;
                        const int p2 = popc_m(F);
                        dual = (k - p2) < p2;
                        S = dual ? ((~F) & kmask) : F;
                        s = dual ? (k - p2) : p2;
                    }
                }
            }
            // systems with more than 16 unknowns (possible only for k > 32) do not fit the register templates: the
            // column is handed to k_nnls with its current passive set (flag bit in the warm-start mask)
            if (KP > 32 && active && (s > 16 || defer_pass)) {
                masks[col] = (unsigned long long)F | DEFER_BIT;
                active = false;
                s = 0;
            }
            const int smax = __reduce_max_sync(0xffffffffu, s);
            bool still = false;
            if (active) {
                const float* Mat = dual ? Ps : Gs;
                float4* xrow = reinterpret_cast<float4*>(xout + (size_t)col * RKP);
                if (smax <= 8)
                    still = round_reg<8, KP>(k, kmask, dual, s, smax, S, Mat, bcol, ucol, zcol, F, state, tol, MAXIT, cnt, pivots_local, xrow);
                else if (smax <= 12)
                    still = round_reg<12, KP>(k, kmask, dual, s, smax, S, Mat, bcol, ucol, zcol, F, state, tol, MAXIT, cnt, pivots_local, xrow);
                else
                    still = round_reg<16, KP>(k, kmask, dual, s, smax, S, Mat, bcol, ucol, zcol, F, state, tol, MAXIT, cnt, pivots_local, xrow);
                if (!still) masks[col] = (unsigned long long)F;
            }
            active = still;
        }
      }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

}  // namespace

bool nnls_reg_supported(int k, int kp) { return (kp == 32 && k <= 32) || (kp == 64 && k <= 62); }

void launch_nnls_reg(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int kp, int num_sms, Counters* cnt,
                     int flags, cudaStream_t st) {
    if (total_cols <= 0) return;
    int nw = kp == 32 ? 8 : 4;   // per-warp staging tiles double in size at KP = 64
    if (const char* e = getenv("LGR_NNLS_REG_WARPS")) nw = atoi(e);
    if (nw < 1) nw = 1;
    if (nw > (kp == 32 ? 8 : 4)) nw = kp == 32 ? 8 : 4;
    while (nw > 1 && total_cols < (long long)num_sms * nw * 32) nw >>= 1;
    const int threads = nw * 32;
    const size_t smem = kp == 32 ? reg_layout<32>(nw).total : reg_layout<64>(nw).total;
    long long want = (total_cols + threads - 1) / threads;
    const int grid = (int)(want < num_sms ? want : num_sms);
    static const bool sync_rounds = getenv("LGR_NNLS_REG_NOSYNC") == nullptr;
    static const bool no_side = getenv("LGR_NNLS_REG_NOSIDE") != nullptr;
    static const bool pair_all = getenv("LGR_NNLS_PAIR_ALL") != nullptr, pair_min = getenv("LGR_NNLS_PAIR_MIN") != nullptr;
    const int fl = flags | (sync_rounds ? 32 : 0) | (no_side ? 64 : 0) | (pair_all ? 128 : 0) | (pair_min ? 256 : 0);
    if (kp == 32)
        k_nnls_reg<32><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cnt, fl);
    else
        k_nnls_reg<64><<<grid, threads, smem, st>>>(groups, ngroups, total_cols, k, cnt, fl);
    LGR_CUDA(cudaGetLastError());
}

void nnls_reg_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_reg<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reg_layout<32>(8).total));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_reg<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reg_layout<64>(4).total));
}

}  // namespace lgr
