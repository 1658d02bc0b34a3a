// liger_b200 -- register-resident block-principal-pivoting NNLS for KKT systems of 17..26 unknowns: TWO lanes per column.
//
// The H solves of a k = 50 run (config C4) start from passive sets of ~25 of the 50 factors, so nearly every column of the
// first ANLS iterations has a system of 17..25 unknowns on its smaller side -- too big for the one-thread-per-column
// register templates of lgr_nnls_reg.cu (16 unknowns = 136 registers), which flags those columns (bit 63 of the warm-start
// mask).  This kernel continues them.  Same problem, exchange rule and tolerances as lgr_nnls_reg.cu (reference:
// RcppPlanc::bppnnls_prod, call site R/cINMF.R:481); what changes is where the s x s principal block lives:
//
//   * a PAIR of adjacent lanes owns one column; lane parity p holds the rows i = 2 m + p of the lower triangle in registers
//     (182 registers for 26 unknowns), fully unrolled code, no shared-memory traffic inside the factorisation;
//   * right-looking Cholesky: at step j the owner of row j publishes 1 / L_jj, both lanes scale their part of column j,
//     exchange it by shuffles (one per entry) and update their own rows -- every update is an independent FMA on registers;
//   * forward substitution runs on the owner of each row, backward substitution sums the two lanes' partial columns; the
//     solution is replicated in both lanes by shuffles;
//   * the k-vector products x = u - P_:A y_A, y = G_:F x_F - b are split by coordinate: lane p owns the 16-byte chunks
//     2 c + p of a factor row (factors 8 c + 4 p .. + 3, c = 0 .. 7), so the two parities read neighbouring chunks of a
//     matrix row -- with a contiguous split (factors 32 p ..) they were 128 bytes apart, the same banks: ncu counted 8
//     shared-memory wavefronts per 128-bit read instead of 4, and these reads are 86 % of the kernel's wavefronts;
//   * u = P b (once per column, when it first takes the dual side) is computed by the WHOLE warp for one column at a time
//     (lane f: u_f and u_f+32, columns of the symmetric P read conflict free): a round in which one pair needs u no longer
//     makes all sixteen pairs run the 400 row reads of a product;
//   * b, u and the expanded solution of the pair's column sit in per-warp shared-memory tiles [factor][column of the warp].
//
// Measured on one GPU's shard of C4 (500 k columns, k = 50, first iterations): the warp-per-column Gauss-Jordan kernel
// (lgr_nnls_wbig.cu) needs 4.8-6.0 ms, a thread per column with the block in shared memory (lgr_nnls_scol.cu) 4.1-4.8 ms
// (one warp per scheduler, every FMA waits for shared-memory operands), this kernel 1.8-2.2 ms.
// Columns whose system exceeds 26 unknowns (k > 52 only) keep their flag and are taken by k_nnls_scol afterwards.
#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int PKP = 64;             // padded factor count
constexpr int PMST = 68;            // row stride of the staged matrices (16-byte aligned rows)
constexpr int PROWS = 65;           // + the padding row (unit diagonal)
constexpr int PSUB = 2048;          // columns scanned / ordered per pass
constexpr int PWARPS = 8;           // 16 columns per warp
constexpr int PCW = 16;             // columns of a warp
constexpr int PTILE = PROWS * PCW;  // words of one [factor][column] tile (row 64 = zero: rhs of the padding unknowns)
constexpr int PSMAX = 26;           // largest system of the register templates
constexpr float PEPS = 1.1920929e-7f;
constexpr unsigned long long PDEFER = 1ull << 63;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ constexpr int poff(int m) { return m * (m + 1); }   // first register of local row m (capacity 2 m + 2)
// factor owned by lane parity p at local position q = 4 c + e: chunk 2 c + p of the row, element e
__device__ __forceinline__ int fown(int q, int p) { return ((q >> 2) << 3) + 4 * p + (q & 3); }

__device__ __forceinline__ void pffma2(float& d0, float& d1, float a0, float a1, float b) {
    unsigned long long d, a, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

struct PairSmem {
    size_t G, P, order, warp0, per_warp, total;
};
__host__ __device__ inline PairSmem pair_layout() {
    PairSmem L;
    L.G = 0;
    L.P = (size_t)PROWS * PMST * 4;
    L.order = 2 * L.P;
    L.warp0 = L.order + (size_t)PSUB * 2;
    L.per_warp = (size_t)3 * PTILE * 4;   // b, u, expanded solution
    L.total = L.warp0 + (size_t)PWARPS * L.per_warp;
    return L;
}

// The steps of the factorisation and of the two substitutions are template recursions: the step index must be a compile-time
// constant for the block to stay in registers, and a `#pragma unroll` over bodies of this size is only a hint.
template <int J, int SMAX>
__device__ __forceinline__ void chol_steps(float (&A)[poff(SMAX / 2)], const uint32_t (&ipk)[(SMAX + 3) / 4], const float* __restrict__ Mat,
                                           bool& ok, const int p, const int lane) {
    constexpr int M = SMAX / 2;
    constexpr int mj = J >> 1, pj = J & 1;
    const int gj = (int)((ipk[J >> 2] >> (8 * (J & 3))) & 0xffu);
    const float orig = Mat[gj * PMST + gj];
    const float dm = A[poff(mj) + J];
    float invm = 0.f;
    bool okm = true;
    if (dm > orig * (PEPS * 4.f))
        invm = rsqrtf(dm);
    else
        okm = false;
    const float invd = __shfl_sync(FULL, invm, (lane & ~1) | pj);
    if (p == pj) {
        A[poff(mj) + J] = invd;
        ok = ok && okm;
    }
    // column J below the diagonal: row J + 1 is local row mj of the odd lane when J is even
    if (!(J & 1)) {
        const float v = A[poff(mj) + J] * invd;
        if (p) A[poff(mj) + J] = v;
    }
#pragma unroll
    for (int m = mj + 1; m < M; ++m) A[poff(m) + J] *= invd;
    float col[SMAX];
#pragma unroll
    for (int c = J + 1; c < SMAX; ++c) col[c] = __shfl_sync(FULL, A[poff(c >> 1) + J], (lane & ~1) | (c & 1));
    if (!(J & 1)) {   // row J + 1 (odd lane); the even lane updates an unused padding register
        const float lij = A[poff(mj) + J];
        A[poff(mj) + J + 1] = fmaf(-lij, col[J + 1], A[poff(mj) + J + 1]);
    }
#pragma unroll
    for (int m = mj + 1; m < M; ++m) {
        const float lij = A[poff(m) + J];
#pragma unroll
        for (int c = J + 1; c <= 2 * m + 1; ++c) A[poff(m) + c] = fmaf(-lij, col[c], A[poff(m) + c]);
    }
    if constexpr (J + 1 < SMAX) chol_steps<J + 1, SMAX>(A, ipk, Mat, ok, p, lane);
}
template <int J, int SMAX>
__device__ __forceinline__ void fwd_steps(const float (&A)[poff(SMAX / 2)], float (&z)[SMAX], const int lane) {
    constexpr int mj = J >> 1;
    float v = z[J];
#pragma unroll
    for (int t = 0; t < J; ++t) v = fmaf(-A[poff(mj) + t], z[t], v);
    z[J] = __shfl_sync(FULL, v * A[poff(mj) + J], (lane & ~1) | (J & 1));
    if constexpr (J + 1 < SMAX) fwd_steps<J + 1, SMAX>(A, z, lane);
}
template <int J, int SMAX>
__device__ __forceinline__ void bwd_steps(const float (&A)[poff(SMAX / 2)], float (&z)[SMAX], float (&zm)[SMAX / 2], const int p,
                                          const int lane) {
    constexpr int M = SMAX / 2;
    constexpr int mj = J >> 1, pj = J & 1;
    float part = 0.f;
    if (!(J & 1)) {
        const float t = A[poff(mj) + J] * zm[mj];
        part = p ? t : 0.f;
    }
#pragma unroll
    for (int m = mj + 1; m < M; ++m) part = fmaf(A[poff(m) + J], zm[m], part);
    part += __shfl_xor_sync(FULL, part, 1);
    const float w = (z[J] - part) * A[poff(mj) + J];
    z[J] = __shfl_sync(FULL, w, (lane & ~1) | pj);
    if (p == pj) zm[mj] = z[J];
    if constexpr (J > 0) bwd_steps<J - 1, SMAX>(A, z, zm, p, lane);
}

// One pivot round of the pair's column with the principal block padded to SMAX (even).  Both lanes of the pair call it with
// identical arguments except p (lane parity); returns true while the column is active.
template <int SMAX>
__device__ __forceinline__ bool round_pair(const int k, const unsigned long long kmask, const bool dual, const unsigned long long S,
                                           const float* __restrict__ Mat, const float* bt, const float* ut, float* zc,
                                           unsigned long long& F, int& state, const float tol, const int MAXIT, Counters* cnt,
                                           unsigned long long& pivots, float* __restrict__ xrow, const bool live, const int p,
                                           const int lane) {
    constexpr int M = SMAX / 2;
    constexpr int NR = poff(M);
    constexpr int NPK = (SMAX + 3) / 4;
    const float* rhs = dual ? ut : bt;
    float A[NR], z[SMAX];
    uint32_t ipk[NPK];
    {
        int idx[SMAX];
        unsigned long long r = S;
#pragma unroll
        for (int t = 0; t < SMAX; ++t) {
            idx[t] = r ? (__ffsll((long long)r) - 1) : PKP;
            r &= r - 1;
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float* mrow = Mat + (p ? idx[2 * m + 1] : idx[2 * m]) * PMST;
#pragma unroll
            for (int c = 0; c <= 2 * m + 1; ++c) A[poff(m) + c] = mrow[idx[c]];
        }
#pragma unroll
        for (int w = 0; w < NPK; ++w) {
            uint32_t pk = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * w + e < SMAX) pk |= (uint32_t)idx[4 * w + e] << (8 * e);
            ipk[w] = pk;
        }
    }
    auto idx_of = [&](int t) { return (int)((ipk[t >> 2] >> (8 * (t & 3))) & 0xffu); };
    // ---- right-looking Cholesky; the diagonal slot of the owner keeps 1 / L_jj; a collapsed pivot pins its coordinate to zero ----
    bool ok = true;
    chol_steps<0, SMAX>(A, ipk, Mat, ok, p, lane);
    if (!__shfl_xor_sync(FULL, ok, 1) || !ok) {
        if (live && !p) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
    }
    // right-hand side of the solved side (0 for the padding unknowns), replicated in both lanes
#pragma unroll
    for (int t = 0; t < SMAX; ++t) z[t] = rhs[idx_of(t) * PCW];
    // ---- forward substitution on the owner of each row, backward substitution summed over the two lanes ----
    fwd_steps<0, SMAX>(A, z, lane);
    {
        float zm[M];
#pragma unroll
        for (int m = 0; m < M; ++m) zm[m] = 0.f;
        bwd_steps<SMAX - 1, SMAX>(A, z, zm, p, lane);
    }
    // ---- infeasible coordinates of the solved side; the (signed) solution expanded by factor id into the pair's tile column ----
    unsigned long long infeas = 0;
    float zmax = 0.f;
#pragma unroll
    for (int t = 0; t < SMAX; ++t) zmax = fmaxf(zmax, fabsf(z[t]));
    const float tolz = dual ? tol : zmax * (PEPS * 8.f);
    const float sgn = dual ? -1.f : 1.f;   // dual: y_A = -z must be >= 0 ; primal: x_F = z must be >= 0
    __syncwarp();   // the partner has finished reading the tile column (products of the previous round)
#pragma unroll
    for (int f = 0; f < 32; ++f) zc[(32 * p + f) * PCW] = 0.f;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < SMAX; ++t) {
        const int gi = idx_of(t);
        const float zt = sgn * z[t];
        if (gi < PKP) {
            if (zt < -tolz) infeas |= 1ull << gi;
            if ((t & 1) == p) zc[gi * PCW] = zt;
        }
    }
    __syncwarp();
    // ---- the k-vector of the other side, this lane's 32 coordinates (fown):  dual  x = u + P (sgn z),  primal  y = -b + G (sgn z) ----
    float acc[32];
    {
        const float sg0 = dual ? 1.f : -1.f;
#pragma unroll
        for (int f = 0; f < 32; ++f) acc[f] = sg0 * rhs[fown(f, p) * PCW];
    }
    for (int g = 0; g < k; ++g) {
        const float zg = zc[g * PCW];
        const float4* row = reinterpret_cast<const float4*>(Mat + g * PMST) + p;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 m4 = row[2 * c];
            pffma2(acc[4 * c + 0], acc[4 * c + 1], m4.x, m4.y, zg);
            pffma2(acc[4 * c + 2], acc[4 * c + 3], m4.z, m4.w, zg);
        }
    }
    {
        // dual: the vector is x on every coordinate (x_A = 0 up to rounding), its maximum scales the tolerance.  (Computed by
        // every lane: the shuffle must not sit in a branch that only the dual pairs of a warp take.)
        float m0 = 0.f, m1 = 0.f;
#pragma unroll
        for (int f = 0; f < 32; f += 2) {
            m0 = fmaxf(m0, fabsf(acc[f]));
            m1 = fmaxf(m1, fabsf(acc[f + 1]));
        }
        m0 = fmaxf(m0, m1);
        m0 = fmaxf(m0, __shfl_xor_sync(FULL, m0, 1));
        const float thr = dual ? m0 * (PEPS * 8.f) : tol;
        uint32_t neg = 0;
#pragma unroll
        for (int f = 0; f < 32; ++f) neg |= (acc[f] < -thr) ? (1u << f) : 0u;
        // local bit 4 c + e is factor 8 c + 4 p + e: spread the nibbles, mask with the other side's set, merge the two lanes
        unsigned long long mine = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) mine |= (unsigned long long)((neg >> (4 * c)) & 0xFu) << (8 * c + 4 * p);
        mine &= dual ? F : ((~F) & kmask);
        const uint32_t plo = __shfl_xor_sync(FULL, (uint32_t)mine, 1), phi = __shfl_xor_sync(FULL, (uint32_t)(mine >> 32), 1);
        infeas |= mine | (unsigned long long)plo | ((unsigned long long)phi << 32);
    }
    int alpha = state & 0xff, beta = (state >> 8) & 0xff, iter = state >> 16;
    if (!live) return false;
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
        if (!p) ++pivots;
        return true;
    }
    if (infeas != 0ull && !p) atomicAdd(&cnt->nnls_unconverged, 1ull);
    // ---- solved: this lane's chunks of the x row ----
    if (dual) {
#pragma unroll
        for (int f = 0; f < 32; ++f)
            if (!((F >> fown(f, p)) & 1ull) || !(acc[f] > 0.f)) acc[f] = 0.f;
    } else {
#pragma unroll
        for (int f = 0; f < 32; ++f) acc[f] = fmaxf(zc[fown(f, p) * PCW], 0.f);
    }
    float4* x4 = reinterpret_cast<float4*>(xrow) + p;   // chunks 2 c + p of the 64-float row
#pragma unroll
    for (int c = 0; c < 8; ++c) x4[2 * c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    return false;
}

__global__ void __launch_bounds__(PWARPS * 32, 1) k_nnls_pair(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols,
                                                              int k, Counters* cnt) {
    extern __shared__ __align__(16) unsigned char psm_raw[];
    const PairSmem L = pair_layout();
    float* Gs = reinterpret_cast<float*>(psm_raw + L.G);
    float* Ps = reinterpret_cast<float*>(psm_raw + L.P);
    unsigned short* order = reinterpret_cast<unsigned short*>(psm_raw + L.order);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, p = lane & 1, cw = lane >> 1;
    float* wbase = reinterpret_cast<float*>(psm_raw + L.warp0 + (size_t)warp * L.per_warp);
    float* bt = wbase + cw;               // b of the pair's column: factor f at bt[f * PCW]
    float* ut = wbase + PTILE + cw;       // u = P b
    float* zc = wbase + 2 * PTILE + cw;   // expanded solution
    __shared__ int next_col;
    __shared__ int hist[68];   // bins (side, system size): 33 primal, 33 dual; [66] = flagged columns of the pass
    const unsigned long long kmask = (1ull << k) - 1ull;   // k <= 62
    const int MAXIT = 4 * k + 16;
    if (!p) {   // rhs of the padding unknowns
        bt[PKP * PCW] = 0.f;
        ut[PKP * PCW] = 0.f;
        zc[PKP * PCW] = 0.f;
    }

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
        if (*grp.ok == 0) continue;   // singular Gram: k_nnls solves the whole group (CTA-uniform)
        __syncthreads();
        for (int e = tid; e < PROWS * PMST; e += blockDim.x) {
            const int a = e / PMST, b = e % PMST;
            float gv = 0.f, pv = 0.f;
            if (a < PKP && b < PKP) {
                gv = (float)grp.gram[a * PKP + b];
                pv = (float)grp.ginv[a * PKP + b];
            } else if (a == PKP && b == PKP) {
                gv = 1.f;
                pv = 1.f;
            }
            Gs[e] = gv;
            Ps[e] = pv;
        }
        const float* __restrict__ rhs_g = reinterpret_cast<const float*>(grp.rhs) + (size_t)c0 * PKP;
        float* __restrict__ xout_g = reinterpret_cast<float*>(grp.x) + (size_t)c0 * PKP;
        unsigned long long* __restrict__ masks_g = grp.mask + c0;
        for (int sub0 = 0; sub0 < ccols; sub0 += PSUB) {
            const int nsub = min(PSUB, ccols - sub0);
            const float* __restrict__ rhs = rhs_g + (size_t)sub0 * PKP;
            float* __restrict__ xout = xout_g + (size_t)sub0 * PKP;
            unsigned long long* __restrict__ masks = masks_g + sub0;
            // ---- the flagged columns of this pass, ordered by the side and the size of their system (primal first, largest
            //      first): the pairs of a warp then read the rows of ONE matrix in the products (two distinct 16-byte
            //      addresses per 128-bit read -- the two parities -- instead of four) ----
            __syncthreads();
            if (tid < 68) hist[tid] = 0;
            __syncthreads();
            auto bin_of = [&](int pc) { return ((k - pc) < pc ? 33 : 0) + 32 - min(min(pc, k - pc), 32); };
            for (int c = tid; c < nsub; c += blockDim.x) {
                const unsigned long long mk = masks[c];
                if (mk & PDEFER) atomicAdd(&hist[bin_of(__popcll(mk & kmask))], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int b = 0; b < 66; ++b) {
                    const int h = hist[b];
                    hist[b] = run;
                    run += h;
                }
                hist[66] = run;   // flagged columns of the pass
                next_col = 0;
            }
            __syncthreads();
            for (int c = tid; c < nsub; c += blockDim.x) {
                const unsigned long long mk = masks[c];
                if (mk & PDEFER) order[atomicAdd(&hist[bin_of(__popcll(mk & kmask))], 1)] = (unsigned short)c;
            }
            __syncthreads();
            const int nflag = hist[66];
            if (nflag == 0) continue;   // CTA-uniform

            bool active = false, exhausted = false, have_u = false;
            int col = 0, state = 0;
            unsigned long long F = 0;
            float tol = 0.f;
            for (;;) {
                // ---- refill: a pair whose column is solved takes the next one of the ordered list ----
                {
                    const bool wants = !active && !exhausted;
                    const unsigned w = __ballot_sync(FULL, wants && !p);
                    if (w) {
                        const int leader = __ffs((int)w) - 1;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(&next_col, __popc(w));
                        base = __shfl_sync(FULL, base, leader);
                        const int ci = base + __popc(w & ((1u << (lane & ~1)) - 1u));   // same value on both lanes of a pair
                        float bm = 0.f;
                        if (wants) {
                            if (ci < nflag) {
                                col = order[ci];
                                const float4* b4 = reinterpret_cast<const float4*>(rhs + (size_t)col * PKP) + p;
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 v = __ldg(b4 + 2 * i);
                                    float* dst = bt + (8 * i + 4 * p) * PCW;
                                    dst[0] = v.x;
                                    dst[PCW] = v.y;
                                    dst[2 * PCW] = v.z;
                                    dst[3 * PCW] = v.w;
                                    bm = fmaxf(bm, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                                }
                                F = masks[col] & kmask;
                                state = 3 | ((k + 1) << 8);
                                have_u = false;
                                active = true;
                            } else {
                                exhausted = true;
                            }
                        }
                        bm = fmaxf(bm, __shfl_xor_sync(FULL, bm, 1));
                        if (wants && active) tol = bm * (PEPS * 8.f);
                        __syncwarp();   // both halves of b are in the tile
                    }
                }
                // CTA-wide round barrier: the unrolled templates are far larger than the instruction cache, so the warps of a
                // CTA walk them together (one instruction stream per SM instead of eight)
                if (!__syncthreads_or(active)) break;
                int s = 0;
                bool dual = false;
                unsigned long long S = 0;
                if (active) {
                    const int pc = __popcll(F);
                    dual = (k - pc) < pc;
                    S = dual ? ((~F) & kmask) : F;
                    s = dual ? (k - pc) : pc;
                    if (s > PSMAX) {   // k > 52 only: stays flagged for k_nnls_scol, with the passive set it has reached
                        if (!p) masks[col] = F | PDEFER;
                        active = false;
                        s = 0;
                        S = 0;
                        dual = false;
                    }
                }
                // ---- u = P b, once per column, when it first takes the dual side.  The whole warp works on one column at a time:
                //      lane f forms u_f and u_f+32 from columns f, f + 32 of the symmetric P (consecutive words: conflict free)
                //      and b broadcast from the tile ----
                {
                    unsigned need = __ballot_sync(FULL, active && dual && !have_u && !p);
                    if (need) {
                        while (need) {
                            const int cwn = (__ffs((int)need) - 1) >> 1;
                            need &= need - 1;
                            const float* bcol = wbase + cwn;
                            float a0 = 0.f, a1 = 0.f;
                            for (int t = 0; t < k; ++t) {
                                const float bv = bcol[t * PCW];
                                a0 = fmaf(Ps[t * PMST + lane], bv, a0);
                                a1 = fmaf(Ps[t * PMST + 32 + lane], bv, a1);
                            }
                            float* ucol = wbase + PTILE + cwn;
                            ucol[lane * PCW] = a0;
                            ucol[(32 + lane) * PCW] = a1;
                        }
                        if (active && dual) have_u = true;
                        __syncwarp();
                    }
                }
                const bool big = __syncthreads_or(s > 22) != 0;
                const float* Mat = dual ? Ps : Gs;
                float* xrow = xout + (size_t)col * PKP;
                // every lane runs the round (an idle pair solves an identity system into its own tile column and registers):
                // the shuffles and warp barriers inside are unconditional
                bool still;
                if (!big)
                    still = round_pair<22>(k, kmask, dual, S, Mat, bt, ut, zc, F, state, tol, MAXIT, cnt, pivots_local, xrow, active, p, lane);
                else
                    still = round_pair<26>(k, kmask, dual, S, Mat, bt, ut, zc, F, state, tol, MAXIT, cnt, pivots_local, xrow, active, p, lane);
                if (active) {
                    if (!still) {
                        if (!p) masks[col] = F;   // clears the flag
                        active = false;
                    }
                }
            }
        }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

}  // namespace

// continues the columns k_nnls_reg<64> flagged (bit 63 of their warm-start mask) whose system has at most 26 unknowns
void launch_nnls_pair(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, cudaStream_t st) {
    if (total_cols <= 0) return;
    const size_t smem = pair_layout().total;
    long long want = (total_cols + PWARPS * PCW - 1) / (PWARPS * PCW);
    const int grid = (int)(want < num_sms ? want : num_sms);
    k_nnls_pair<<<grid, PWARPS * 32, smem, st>>>(groups, ngroups, total_cols, k, cnt);
    LGR_CUDA(cudaGetLastError());
}

void nnls_pair_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_layout().total));
}

}  // namespace lgr
