// liger_b200 -- thread-per-column fp32 NNLS with the principal block in SHARED MEMORY, for the H-solve columns whose KKT
// system is too big for the register templates of lgr_nnls_reg.cu (more than 16 unknowns: 32 < k <= 62 while |F| is near
// k / 2, i.e. most columns of the first ANLS iterations of a k = 50 run).  k_nnls_reg<64> flags those columns in their
// warm-start mask (bit 63); this kernel continues them from the passive set they had reached.  Same problem, same exchange
// rule and the same tolerances as lgr_nnls_reg.cu (reference: RcppPlanc::bppnnls_prod, call site R/cINMF.R:481).
//
// Role: systems of up to 26 unknowns go to k_nnls_pair (lgr_nnls_pair.cu: two lanes per column, block in registers, ~2 x
// faster than this kernel); what is left for this one are the systems of 27..31 unknowns (k > 52) and, with
// LGR_NNLS_NO_PAIR, everything k_nnls_reg<64> flags.
//
// Why not a warp per column (lgr_nnls_wbig.cu, the round-1 continuation kernel): a Gauss-Jordan sweep that moves its
// pivot row by shuffles spends ~2500 warp instructions per column and round at s = 25; a thread that owns its column
// spends s^3 / 6 FMAs, ~300 warp instructions per column.  The price is storage: the lower triangle of an s x s block does
// not fit the register file above s = 16, so it lives in shared memory, [word][thread] (conflict free), and the
// factorisation is row by row (Cholesky-Banachiewicz, two columns at a time: three shared-memory reads per two FMAs).
// 183 KB of triangles leave one warp per scheduler, so every FMA waits for its shared-memory operands: measured 4.1-4.8 ms
// on a C4 shard against 4.8-6.0 ms for the warp-per-column kernel (and 1.8-2.2 ms for k_nnls_pair).
// b and u = P b stay in registers (the matrix no longer needs them).  All loops run to the warp's largest system, columns
// are ordered by system size, smaller systems are padded with identity rows (index 64 of the staged matrix), so a warp
// never diverges inside a round.
#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int SKP = 64;             // padded factor count
constexpr int SMST = 68;            // row stride of the staged matrices (16-byte aligned rows)
constexpr int SROWS = 65;           // + the padding row (unit diagonal)
constexpr int SSUB = 2048;          // columns scanned / ordered per pass
constexpr float SEPS = 1.1920929e-7f;
constexpr unsigned long long SDEFER = 1ull << 63;

__host__ __device__ constexpr int stri(int i) { return i * (i + 1) / 2; }
// developer counters (LGR_SCOL_DBG): [0, 40) rounds per column, [40] longest CTA (cycles), [41] sum of CTA cycles,
// [42] columns, [43] warp rounds, [44] sum of the warps' system sizes over their rounds
__device__ unsigned long long g_scol_dbg[56];   // [48, 54): cycles of CTA 0 / warp 0 per phase

__device__ __forceinline__ void sffma2(float& d0, float& d1, float a0, float a1, float b) {
    unsigned long long d, a, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

struct ScolSmem {
    size_t G, P, order, scratch, total;
    int words;   // per-thread scratch words: lower triangle of the largest system + the solution vector + 32 packed factor ids
};
__host__ __device__ inline ScolSmem scol_layout(int k, int nt) {
    ScolSmem L;
    const int smaxk = k / 2;   // s = min(|F|, k - |F|)
    L.words = stri(smaxk) + 32 + 8;
    L.G = 0;
    L.P = (size_t)SROWS * SMST * 4;
    L.order = 2 * L.P;
    L.scratch = L.order + (size_t)SSUB * 2;
    L.total = L.scratch + (size_t)L.words * nt * 4;
    return L;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_nnls_scol(const NnlsGroup* __restrict__ groups, int ngroups, long long total_cols, int k,
                                                     Counters* cnt, int dbg) {
    extern __shared__ __align__(16) unsigned char ssm_raw[];
    const long long t_start = dbg ? clock64() : 0;
    const ScolSmem L = scol_layout(k, NT);
    float* Gs = reinterpret_cast<float*>(ssm_raw + L.G);
    float* Ps = reinterpret_cast<float*>(ssm_raw + L.P);
    unsigned short* order = reinterpret_cast<unsigned short*>(ssm_raw + L.order);
    const int tid = threadIdx.x, lane = tid & 31;
    // this thread's scratch: word w at sc[w * NT].  Words [0, tri(smaxk)): the lower triangle (row i at tri(i)), later the
    // expanded solution (65 words) and the k-vector of the round (64 words); then 32 words: the solution z; then 8 words: the
    // factor ids of the unknowns (64 = padding), so that the loops below index them instead of walking the mask bit by bit.
    float* sc = reinterpret_cast<float*>(ssm_raw + L.scratch) + tid;
    float* zs = sc + (size_t)(L.words - 40) * NT;
    uint32_t* iws = reinterpret_cast<uint32_t*>(sc + (size_t)(L.words - 8) * NT);   // factor id of unknown t: byte t % 4 of word t / 4
    float* ze = sc;                  // expanded solution by factor id, [0, 64] (64 = dump of the padding unknowns)
    float* xs = sc + 65 * NT;        // k-vector of the round (x or y), [0, 64)
    __shared__ int next_col;
    __shared__ int hist[34];
    const unsigned long long kmask = (1ull << k) - 1ull;   // k <= 62
    const int MAXIT = 4 * k + 16;

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
        for (int e = tid; e < SROWS * SMST; e += NT) {
            const int a = e / SMST, b = e % SMST;
            float gv = 0.f, pv = 0.f;
            if (a < SKP && b < SKP) {
                gv = (float)grp.gram[a * SKP + b];
                pv = (float)grp.ginv[a * SKP + b];
            } else if (a == SKP && b == SKP) {
                gv = 1.f;
                pv = 1.f;
            }
            Gs[e] = gv;
            Ps[e] = pv;
        }
        const float* __restrict__ rhs_g = reinterpret_cast<const float*>(grp.rhs) + (size_t)c0 * SKP;
        float* __restrict__ xout_g = reinterpret_cast<float*>(grp.x) + (size_t)c0 * SKP;
        unsigned long long* __restrict__ masks_g = grp.mask + c0;
        for (int sub0 = 0; sub0 < ccols; sub0 += SSUB) {
            const int nsub = min(SSUB, ccols - sub0);
            const float* __restrict__ rhs = rhs_g + (size_t)sub0 * SKP;
            float* __restrict__ xout = xout_g + (size_t)sub0 * SKP;
            unsigned long long* __restrict__ masks = masks_g + sub0;
            // ---- the flagged columns of this pass, ordered by the size of their system (largest first) ----
            __syncthreads();
            if (tid < 34) hist[tid] = 0;
            __syncthreads();
            for (int c = tid; c < nsub; c += NT) {
                const unsigned long long mk = masks[c];
                if (mk & SDEFER) {
                    const int p = __popcll(mk & kmask);
                    atomicAdd(&hist[32 - min(min(p, k - p), 32)], 1);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int b = 0; b < 33; ++b) {
                    const int h = hist[b];
                    hist[b] = run;
                    run += h;
                }
                hist[33] = run;   // flagged columns of the pass
                next_col = 0;
            }
            __syncthreads();
            for (int c = tid; c < nsub; c += NT) {
                const unsigned long long mk = masks[c];
                if (mk & SDEFER) {
                    const int p = __popcll(mk & kmask);
                    order[atomicAdd(&hist[32 - min(min(p, k - p), 32)], 1)] = (unsigned short)c;
                }
            }
            __syncthreads();
            const int nflag = hist[33];
            if (nflag == 0) continue;   // CTA-uniform

            bool active = false, exhausted = false, have_u = false;
            int col = 0, state = 0;
            unsigned long long F = 0;
            float tol = 0.f;
            float b[SKP], u[SKP];
#pragma unroll
            for (int f = 0; f < SKP; ++f) b[f] = u[f] = 0.f;
            for (;;) {
                // ---- refill: a lane whose column is solved takes the next one of the ordered list ----
                {
                    const bool wants = !active && !exhausted;
                    const unsigned w = __ballot_sync(0xffffffffu, wants);
                    if (w) {
                        const int leader = __ffs((int)w) - 1;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(&next_col, __popc(w));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (wants) {
                            const int ci = base + __popc(w & ((1u << lane) - 1u));
                            if (ci < nflag) {
                                col = order[ci];
                                const float4* b4 = reinterpret_cast<const float4*>(rhs + (size_t)col * SKP);
                                float bm = 0.f;
#pragma unroll
                                for (int i = 0; i < SKP / 4; ++i) {
                                    const float4 v = __ldg(b4 + i);
                                    b[4 * i] = v.x;
                                    b[4 * i + 1] = v.y;
                                    b[4 * i + 2] = v.z;
                                    b[4 * i + 3] = v.w;
                                    bm = fmaxf(bm, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                                }
                                F = masks[col] & kmask;
                                state = 3 | ((k + 1) << 8);
                                tol = bm * (SEPS * 8.f);
                                have_u = false;
                                active = true;
                            } else {
                                exhausted = true;
                            }
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, active)) break;
                int s = 0;
                bool dual = false;
                unsigned long long S = 0;
                if (active) {
                    const int p = __popcll(F);
                    dual = (k - p) < p;
                    S = dual ? ((~F) & kmask) : F;
                    s = dual ? (k - p) : p;
                }
                // ---- u = P b, once per column, when it first takes the dual side (b travels through the scratch so that the
                //      loop over the factors is a dynamic one; the rows of P are broadcast reads) ----
                if (__any_sync(0xffffffffu, active && dual && !have_u)) {
                    const bool need = active && dual && !have_u;
#pragma unroll
                    for (int f = 0; f < SKP; ++f) {
                        sc[f * NT] = b[f];
                        if (need) u[f] = 0.f;
                    }
#pragma unroll 2
                    for (int t = 0; t < k; ++t) {
                        const float bt = need ? sc[t * NT] : 0.f;
                        const float4* row = reinterpret_cast<const float4*>(Ps + t * SMST);
#pragma unroll
                        for (int c = 0; c < SKP / 4; ++c) {
                            const float4 m4 = row[c];
                            sffma2(u[4 * c + 0], u[4 * c + 1], m4.x, m4.y, bt);
                            sffma2(u[4 * c + 2], u[4 * c + 3], m4.z, m4.w, bt);
                        }
                    }
                    if (need) have_u = true;
                }
                const int smax = __reduce_max_sync(0xffffffffu, s);
                const bool prof = dbg && blockIdx.x == 0 && tid == 0;
                long long pc = prof ? clock64() : 0;
                auto phase = [&](int id) {
                    if (prof) {
                        const long long now = clock64();
                        g_scol_dbg[48 + id] += (unsigned long long)(now - pc);
                        pc = now;
                    }
                };
                if (dbg && lane == 0) {
                    atomicAdd(&g_scol_dbg[43], 1ull);
                    atomicAdd(&g_scol_dbg[44], (unsigned long long)smax);
                }
                const float* Mat = dual ? Ps : Gs;
                // ---- right-hand side of the solved side: z_t = rhs[idx_t] (through the scratch), 0 for the padding unknowns ----
#pragma unroll
                for (int f = 0; f < SKP; ++f) sc[f * NT] = dual ? u[f] : b[f];
                sc[SKP * NT] = 0.f;
                {   // factor ids of the unknowns, packed four to a word (static loop: no dynamic register indexing)
                    unsigned long long r = S;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        uint32_t pk = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t gi = r ? (uint32_t)(__ffsll((long long)r) - 1) : (uint32_t)SKP;
                            r &= r - 1;
                            pk |= gi << (8 * e);
                        }
                        iws[w * NT] = pk;
                    }
                }
                auto idx_of = [&](int t) { return (int)((iws[(t >> 2) * NT] >> (8 * (t & 3))) & 0xffu); };
#pragma unroll 4
                for (int t = 0; t < smax; ++t) zs[t * NT] = sc[idx_of(t) * NT];
                phase(0);
                // ---- gather the principal block (lower triangle; padding index 64 = identity row), four columns at a time ----
                {
                    float* arow = sc;
                    for (int i = 0; i < smax; ++i) {
                        const float* mrow = Mat + idx_of(i) * SMST;
                        for (int j0 = 0; j0 <= i; j0 += 4) {
                            const uint32_t pk = iws[(j0 >> 2) * NT];
                            const float v0 = mrow[pk & 0xffu], v1 = mrow[(pk >> 8) & 0xffu], v2 = mrow[(pk >> 16) & 0xffu], v3 = mrow[pk >> 24];
                            arow[j0 * NT] = v0;
                            if (j0 + 1 <= i) arow[(j0 + 1) * NT] = v1;
                            if (j0 + 2 <= i) arow[(j0 + 2) * NT] = v2;
                            if (j0 + 3 <= i) arow[(j0 + 3) * NT] = v3;
                        }
                        arow += (i + 1) * NT;
                    }
                }
                phase(1);
                // ---- Cholesky, row by row; the diagonal slot holds 1 / L_ii; a collapsed pivot pins its coordinate to zero ----
                bool ok = true;
                {
                    float* rowi = sc;
                    for (int i = 0; i < smax; ++i) {
                        int j = 0;
                        const float* rowj = sc;
                        for (; j + 1 < i; j += 2) {   // columns j and j + 1, both below the diagonal
                            const float* rowj1 = rowj + (j + 1) * NT;
                            float s0 = rowi[j * NT], s1 = rowi[(j + 1) * NT];
#pragma unroll 4
                            for (int t = 0; t < j; ++t) {
                                const float li = rowi[t * NT];
                                s0 = fmaf(-li, rowj[t * NT], s0);
                                s1 = fmaf(-li, rowj1[t * NT], s1);
                            }
                            const float l0 = s0 * rowj[j * NT];
                            rowi[j * NT] = l0;
                            s1 = fmaf(-l0, rowj1[j * NT], s1);
                            rowi[(j + 1) * NT] = s1 * rowj1[(j + 1) * NT];
                            rowj = rowj1 + (j + 2) * NT;
                        }
                        if (j < i) {   // one more column below the diagonal
                            float s0 = rowi[j * NT];
#pragma unroll 4
                            for (int t = 0; t < j; ++t) s0 = fmaf(-rowi[t * NT], rowj[t * NT], s0);
                            rowi[j * NT] = s0 * rowj[j * NT];
                        }
                        const float orig = rowi[i * NT];
                        float d0 = orig, d1 = 0.f;
                        int t = 0;
#pragma unroll 2
                        for (; t + 1 < i; t += 2) {
                            const float a0 = rowi[t * NT], a1 = rowi[(t + 1) * NT];
                            d0 = fmaf(-a0, a0, d0);
                            d1 = fmaf(-a1, a1, d1);
                        }
                        if (t < i) {
                            const float a0 = rowi[t * NT];
                            d0 = fmaf(-a0, a0, d0);
                        }
                        const float d = d0 + d1;
                        float invd = 0.f;
                        if (d > orig * (SEPS * 4.f))
                            invd = rsqrtf(d);
                        else
                            ok = false;
                        rowi[i * NT] = invd;
                        rowi += (i + 1) * NT;
                    }
                }
                if (active && !ok) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
                phase(2);
                // ---- forward and backward substitution ----
                {
                    const float* rowj = sc;
                    for (int j = 0; j < smax; ++j) {
                        float v0 = zs[j * NT], v1 = 0.f;
                        int t = 0;
#pragma unroll 2
                        for (; t + 1 < j; t += 2) {
                            v0 = fmaf(-rowj[t * NT], zs[t * NT], v0);
                            v1 = fmaf(-rowj[(t + 1) * NT], zs[(t + 1) * NT], v1);
                        }
                        if (t < j) v0 = fmaf(-rowj[t * NT], zs[t * NT], v0);
                        zs[j * NT] = (v0 + v1) * rowj[j * NT];
                        rowj += (j + 1) * NT;
                    }
                    for (int j = smax - 1; j >= 0; --j) {
                        float v0 = zs[j * NT], v1 = 0.f;
                        const float* col = sc + (size_t)(stri(j + 1) + j) * NT;   // element (j + 1, j); (i + 1, j) is i + 2 words further
                        int i = j + 1;
#pragma unroll 2
                        for (; i + 1 < smax; i += 2) {
                            const float* col1 = col + (i + 1) * NT;
                            v0 = fmaf(-col[0], zs[i * NT], v0);
                            v1 = fmaf(-col1[0], zs[(i + 1) * NT], v1);
                            col = col1 + (i + 2) * NT;
                        }
                        if (i < smax) v0 = fmaf(-col[0], zs[i * NT], v0);
                        zs[j * NT] = (v0 + v1) * sc[(size_t)(stri(j) + j) * NT];
                    }
                }
                phase(3);
                // ---- infeasible coordinates of the solved side; the (signed) solution expanded by factor id ----
                unsigned long long infeas = 0;
                float zmax = 0.f;
#pragma unroll 4
                for (int t = 0; t < smax; ++t) zmax = fmaxf(zmax, fabsf(zs[t * NT]));
                const float tolz = dual ? tol : zmax * (SEPS * 8.f);
                const float sgn = dual ? -1.f : 1.f;   // dual: y_A = -z must be >= 0 ; primal: x_F = z must be >= 0
#pragma unroll
                for (int f = 0; f <= SKP; ++f) ze[f * NT] = 0.f;
                {   // (the factor ids move to registers first: ze overwrites nothing of iws / zs, but keeps the loop free of the lambda)
#pragma unroll 4
                    for (int t = 0; t < smax; ++t) {
                        const int gi = idx_of(t);
                        const float zt = sgn * zs[t * NT];
                        if (gi < SKP && zt < -tolz) infeas |= 1ull << gi;
                        ze[gi * NT] = gi < SKP ? zt : 0.f;
                    }
                }
                // ---- the k-vector of the other side:  dual  x = u + P (sgn z),  primal  y = -b + G (sgn z);  32 coordinates at a time ----
                float amax = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float acc[32];
#pragma unroll
                    for (int f = 0; f < 32; ++f) acc[f] = dual ? u[32 * h + f] : -b[32 * h + f];
#pragma unroll 2
                    for (int gg = 0; gg < k; ++gg) {
                        const float zg = ze[gg * NT];
                        const float4* row = reinterpret_cast<const float4*>(Mat + gg * SMST + 32 * h);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 m4 = row[c];
                            sffma2(acc[4 * c + 0], acc[4 * c + 1], m4.x, m4.y, zg);
                            sffma2(acc[4 * c + 2], acc[4 * c + 3], m4.z, m4.w, zg);
                        }
                    }
#pragma unroll
                    for (int f = 0; f < 32; ++f) {
                        xs[(32 * h + f) * NT] = acc[f];
                        amax = fmaxf(amax, fabsf(acc[f]));
                    }
                }
                phase(4);
                // in the dual case the vector is x on every coordinate (x_A = 0 up to rounding): its maximum scales the tolerance
                const float thr = dual ? amax * (SEPS * 8.f) : tol;
                {
                    unsigned long long neg = 0;
#pragma unroll
                    for (int f = 0; f < SKP; ++f) neg |= (xs[f * NT] < -thr) ? (1ull << f) : 0ull;
                    infeas |= neg & (dual ? F : ((~F) & kmask));
                }
                if (!active) continue;
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
                    ++pivots_local;
                    continue;
                }
                if (infeas != 0ull) atomicAdd(&cnt->nnls_unconverged, 1ull);
                // ---- solved: x row (64 floats), passive set (clears the flag) ----
                {
                    float4* xrow = reinterpret_cast<float4*>(xout + (size_t)col * SKP);
#pragma unroll
                    for (int c = 0; c < SKP / 4; ++c) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int f = 4 * c + e;
                            if (dual) {
                                const float x = xs[f * NT];
                                v[e] = (((F >> f) & 1ull) && x > 0.f) ? x : 0.f;
                            } else {
                                v[e] = fmaxf(ze[f * NT], 0.f);
                            }
                        }
                        xrow[c] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                    masks[col] = F;
                    active = false;
                    if (dbg) {
                        atomicAdd(&g_scol_dbg[min(state >> 16, 39)], 1ull);
                        atomicAdd(&g_scol_dbg[42], 1ull);
                    }
                }
            }
        }
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
    if (dbg && tid == 0) {
        const unsigned long long cyc = (unsigned long long)(clock64() - t_start);
        atomicMax(&g_scol_dbg[40], cyc);
        atomicAdd(&g_scol_dbg[41], cyc);
    }
}

}  // namespace

static int scol_threads(int k) {
    return scol_layout(k, 128).total <= (size_t)227 * 1024 - 2048 ? 128 : 64;
}

// continues the columns k_nnls_reg<64> flagged (bit 63 of their warm-start mask); all other columns are left alone
void launch_nnls_scol(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, cudaStream_t st) {
    if (total_cols <= 0) return;
    const int nt = scol_threads(k);
    const size_t smem = scol_layout(k, nt).total;
    long long want = (total_cols + nt - 1) / nt;
    const int grid = (int)(want < num_sms ? want : num_sms);
    static const int dbg = getenv("LGR_SCOL_DBG") ? 1 : 0;
    if (dbg) {
        unsigned long long z[56] = {0};
        LGR_CUDA(cudaMemcpyToSymbolAsync(g_scol_dbg, z, sizeof(z), 0, cudaMemcpyHostToDevice, st));
    }
    if (nt == 128)
        k_nnls_scol<128><<<grid, 128, smem, st>>>(groups, ngroups, total_cols, k, cnt, dbg);
    else
        k_nnls_scol<64><<<grid, 64, smem, st>>>(groups, ngroups, total_cols, k, cnt, dbg);
    LGR_CUDA(cudaGetLastError());
    if (dbg) {
        unsigned long long h[56];
        LGR_CUDA(cudaStreamSynchronize(st));
        LGR_CUDA(cudaMemcpyFromSymbol(h, g_scol_dbg, sizeof(h)));
        fprintf(stderr, "[scol dbg] grid %d x %d: columns %llu, warp rounds %llu (mean system %.1f), CTA cycles max %llu mean %llu; rounds/column:",
                grid, nt, h[42], h[43], h[43] ? (double)h[44] / (double)h[43] : 0.0, h[40], h[41] / (unsigned long long)grid);
        for (int i = 0; i < 40; ++i)
            if (h[i]) fprintf(stderr, " %d:%llu", i, h[i]);
        fprintf(stderr, "\n[scol dbg] CTA 0 warp 0 cycles: rhs %llu gather %llu cholesky %llu solves %llu product %llu\n", h[48], h[49], h[50], h[51], h[52]);
    }
}

void nnls_scol_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_scol<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    LGR_CUDA(cudaFuncSetAttribute(k_nnls_scol<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
}

}  // namespace lgr
