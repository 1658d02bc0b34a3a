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
template <int KP>
__global__ void __launch_bounds__(128) k_gram_inv(const GramSpec* __restrict__ specs, int k) {
    const GramSpec sp = specs[blockIdx.x];
    __shared__ double M[KP][KP + 1];
    __shared__ double colbuf[KP];
    __shared__ double piv_s;
    __shared__ int ok_s;
    const int tid = threadIdx.x;
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
    __syncthreads();
    for (int c = 0; c < k; ++c) {  // in-place Gauss-Jordan, SPD => no pivoting
        if (tid == 0) {
            const double p = M[c][c];
            if (!(p > 1e-11 * maxdiag)) ok_s = 0;
            piv_s = p;
        }
        for (int r = tid; r < k; r += blockDim.x) colbuf[r] = M[r][c];
        __syncthreads();
        if (!ok_s) break;
        const double ip = 1.0 / piv_s;
        for (int col = tid; col < k; col += blockDim.x) M[c][col] = (col == c) ? ip : M[c][col] * ip;
        __syncthreads();
        for (int e = tid; e < k * k; e += blockDim.x) {
            const int r = e / k, col = e % k;
            if (r == c) continue;
            const double base = (col == c) ? 0.0 : M[r][col];
            M[r][col] = base - colbuf[r] * M[c][col];
        }
        __syncthreads();
    }
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
        k_gram_inv<32><<<nspecs, 128, 0, st>>>(specs, k);
    else
        k_gram_inv<64><<<nspecs, 128, 0, st>>>(specs, k);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Solve M_II z = r_I by Cholesky (left-looking, packed lower-triangular factor in the pool).
//   idx[t*IS]   t-th member of the index set I (ascending)
//   M           k x k symmetric matrix in shared memory (row stride ldm)
//   r[f*RS]     right-hand side indexed by factor id
//   fac[e*FS]   packed factor, e = i(i+1)/2 + j;  the diagonal stores 1/L_jj
//   z[t*FS]     solution in compact order (same stride as the factor: it lives in the pool)
// A non-positive pivot pins that coordinate to zero (numerically singular principal block).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool chol_solve(int n, const unsigned char* idx, int IS, const T* __restrict__ M, int ldm,
                                           const T* r, int RS, T* fac, T* z, int FS) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        const int gj = idx[j * IS];
        const T* Mj = M + gj * ldm;
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
        // rows below, two at a time (shares the loads of row j, doubles the ILP of the FMA chain)
        int i = j + 1;
        for (; i + 1 < n; i += 2) {
            T* r0 = fac + (i * (i + 1) / 2) * FS;
            T* r1 = fac + ((i + 1) * (i + 2) / 2) * FS;
            T s0 = Mj[idx[i * IS]], s1 = Mj[idx[(i + 1) * IS]];
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
            T s0 = Mj[idx[i * IS]];
#pragma unroll 4
            for (int t = 0; t < j; ++t) s0 -= r0[t * FS] * rowj[t * FS];
            r0[j * FS] = s0 * invd;
        }
    }
    for (int j = 0; j < n; ++j) {  // forward substitution
        const T* rowj = fac + (j * (j + 1) / 2) * FS;
        T s = r[idx[j * IS] * RS];
#pragma unroll 4
        for (int t = 0; t < j; ++t) s -= rowj[t * FS] * z[t * FS];
        z[j * FS] = s * rowj[j * FS];
    }
    for (int j = n - 1; j >= 0; --j) {  // back substitution
        T s = z[j * FS];
        for (int i = j + 1; i < n; ++i) s -= fac[((i * (i + 1) / 2) + j) * FS] * z[i * FS];
        z[j * FS] = s * fac[((j * (j + 1) / 2) + j) * FS];
    }
    return ok;
}

struct NnlsSmem {
    // byte offsets of the carve-up (host and device must agree)
    size_t G, P, b, u, x, pool, mask, tol, state, listA, listB, sorted, bins, wmeta, size, idx, total;
};

__host__ __device__ inline NnlsSmem nnls_layout(int kp, int nt, int pool_elems, size_t ts) {
    NnlsSmem L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += (bytes + 15) & ~(size_t)15;
        return at;
    };
    L.G = take((size_t)kp * kp * ts);
    L.P = take((size_t)kp * kp * ts);
    L.b = take((size_t)kp * (nt + 1) * ts);
    L.u = take((size_t)kp * (nt + 1) * ts);
    L.x = take((size_t)kp * (nt + 1) * ts);
    L.pool = take((size_t)pool_elems * ts);
    L.mask = take((size_t)nt * 8);
    L.tol = take((size_t)nt * ts);
    L.state = take((size_t)nt * 4);
    L.listA = take((size_t)nt * 4);
    L.listB = take((size_t)nt * 4);
    L.sorted = take((size_t)nt * 4);
    L.bins = take((size_t)(kp + 2) * 4 * 2);
    L.wmeta = take((size_t)(nt / 32 + 1) * 4 * 4);
    L.size = take((size_t)nt);
    L.idx = take((size_t)kp * nt);
    L.total = o;
    return L;
}

template <typename T, int KP>
__global__ void k_nnls(const NnlsGroup* __restrict__ groups, int ngroups, int total_blocks, int k, int pool_elems,
                       Counters* cnt, int flags) {
    const int NT = blockDim.x;
    const int NW = NT >> 5;
    const int XS = NT + 1;  // padded stride of the transposed [factor][slot] tiles
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NnlsSmem L = nnls_layout(KP, NT, pool_elems, sizeof(T));
    T* Gs = reinterpret_cast<T*>(smem_raw + L.G);
    T* Ps = reinterpret_cast<T*>(smem_raw + L.P);
    T* bS = reinterpret_cast<T*>(smem_raw + L.b);
    T* uS = reinterpret_cast<T*>(smem_raw + L.u);
    T* xS = reinterpret_cast<T*>(smem_raw + L.x);
    T* pool = reinterpret_cast<T*>(smem_raw + L.pool);
    unsigned long long* maskS = reinterpret_cast<unsigned long long*>(smem_raw + L.mask);
    T* tolS = reinterpret_cast<T*>(smem_raw + L.tol);
    int* stateS = reinterpret_cast<int*>(smem_raw + L.state);  // alpha | beta << 8 | iter << 16
    int* listA = reinterpret_cast<int*>(smem_raw + L.listA);
    int* listB = reinterpret_cast<int*>(smem_raw + L.listB);
    int* sorted = reinterpret_cast<int*>(smem_raw + L.sorted);
    int* bins = reinterpret_cast<int*>(smem_raw + L.bins);     // [0..KP+1] counts, [KP+2..] starts
    int* wmeta = reinterpret_cast<int*>(smem_raw + L.wmeta);   // per warp: offset, lanes, tri ; [4*NW] = nact
    unsigned char* sizeS = smem_raw + L.size;
    unsigned char* idxS = smem_raw + L.idx;
    __shared__ int wsum[32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool cold_start = (flags & 1) != 0;
    const bool want_hist = (flags & 2) != 0;
    const bool hist_p = (flags & 4) != 0;
    const bool no_dual = (flags & 8) != 0;
    const unsigned long long kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    const int MAXIT = 4 * k + 16;

    int cur_group = -1;
    int dual_ok = 0;
    unsigned long long pivots_local = 0;
    for (int blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
        int g = 0;
        while (g + 1 < ngroups && blk >= groups[g + 1].blk_off) ++g;
        const NnlsGroup grp = groups[g];
        __syncthreads();  // previous tile fully consumed
        if (g != cur_group) {
            cur_group = g;
            dual_ok = no_dual ? 0 : *grp.ok;
            for (int e = tid; e < KP * KP; e += NT) {
                Gs[e] = (T)grp.gram[e];
                Ps[e] = (T)grp.ginv[e];
            }
        }
        const int col0 = (blk - grp.blk_off) * NT;
        const int ncol = min(NT, grp.ncols - col0);
        const int ld = grp.ld;
        const T* __restrict__ rhs = reinterpret_cast<const T*>(grp.rhs) + (size_t)col0 * ld;
        T* __restrict__ xout = reinterpret_cast<T*>(grp.x) + (size_t)col0 * ld;
        for (int e = tid; e < ncol * KP; e += NT) {  // coalesced load, transposed into bS[f][slot]
            const int j = e / KP, f = e % KP;
            bS[f * XS + j] = (f < k) ? rhs[(size_t)j * ld + f] : (T)0;
        }
        if (tid < ncol) {
            maskS[tid] = cold_start ? 0ull : (grp.mask[col0 + tid] & kmask);
            stateS[tid] = 3 | ((k + 1) << 8);
            listA[tid] = tid;
        }
        __syncthreads();
        if (tid < ncol) {
            const T* b = bS + tid;
            T bm = (T)0;
            for (int f = 0; f < k; ++f) bm = fmax(bm, fabs(b[f * XS]));
            tolS[tid] = bm * Eps<T>::v() * (T)8;
            if (dual_ok) {  // u = P b : register accumulators, broadcast shared-memory reads of P
                for (int a0 = 0; a0 < k; a0 += 16) {
                    T acc[16];
#pragma unroll
                    for (int a = 0; a < 16; ++a) acc[a] = (T)0;
                    for (int t = 0; t < k; ++t) {
                        const T bt = b[t * XS];
                        const T* Pt = Ps + t * KP + a0;
#pragma unroll
                        for (int a = 0; a < 16; ++a) acc[a] = fma(Pt[a], bt, acc[a]);
                    }
#pragma unroll
                    for (int a = 0; a < 16; ++a)
                        if (a0 + a < k) uS[(a0 + a) * XS + tid] = acc[a];
                }
            }
        }

        int nact = ncol;
        int* cur = listA;
        int* nxt = listB;
        while (nact > 0) {
            // ---- 1. size of every active slot's solve, counting sort by size ----
            for (int e = tid; e < 2 * (KP + 2); e += NT) bins[e] = 0;
            __syncthreads();
            int myslot = -1, mysize = 0, myrank = 0;
            if (tid < nact) {
                myslot = cur[tid];
                const int p = __popcll(maskS[myslot]);
                const int q = k - p;
                mysize = (dual_ok && q < p) ? q : p;
                sizeS[myslot] = (unsigned char)mysize;
                myrank = atomicAdd(&bins[mysize], 1);
            }
            __syncthreads();
            if (warp == 0) {  // exclusive prefix over the KP + 1 size bins
                int carry = 0;
                for (int base = 0; base < KP + 1; base += 32) {
                    const int i = base + lane;
                    const int v = (i < KP + 1) ? bins[i] : 0;
                    int incl = v;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int n = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += n;
                    }
                    if (i < KP + 1) bins[KP + 2 + i] = carry + incl - v;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
            if (tid < nact) sorted[bins[KP + 2 + mysize] + myrank] = myslot;
            __syncthreads();
            // ---- 2. pool allocation per warp (sorted ascending: the warp's largest size is its last slot) ----
            if (tid == 0) {
                int used = 0;
                for (int w = 0; w < NW; ++w) {
                    const int first = w * 32;
                    int lanes = 0, need = 1, off = used;
                    if (first < nact) {
                        const int last = min(first + 32, nact) - 1;
                        const int smax = sizeS[sorted[last]];
                        need = smax * (smax + 1) / 2 + smax;  // factor + compact solution
                        if (need < 1) need = 1;
                        lanes = min(last - first + 1, (pool_elems - used) / need);
                        used += lanes * need;
                    }
                    wmeta[w * 4 + 0] = off;
                    wmeta[w * 4 + 1] = lanes;
                    wmeta[w * 4 + 2] = need;
                }
            }
            __syncthreads();
            // ---- 3. one pivot round for every slot that got pool space ----
            bool keep = false;  // stays in the active list (not finished, or deferred for lack of pool space)
            if (tid < nact) {
                const int slot = sorted[tid];
                const int wl = wmeta[warp * 4 + 1];
                if (lane >= wl) {
                    keep = true;  // deferred to the next pass
                } else {
                    const int s = sizeS[slot];
                    T* fac = pool + wmeta[warp * 4 + 0] + lane;
                    const int FS = wl;
                    unsigned long long F = maskS[slot];
                    const int p = __popcll(F);
                    const bool dual = dual_ok && (k - p) < p;
                    unsigned char* idx = idxS + slot;
                    {
                        unsigned long long r = dual ? ((~F) & kmask) : F;
                        int t = 0;
                        while (r) {
                            idx[t * NT] = (unsigned char)(__ffsll((long long)r) - 1);
                            r &= r - 1;
                            ++t;
                        }
                    }
                    T* z = fac + (size_t)(s * (s + 1) / 2) * FS;  // compact solution right after the factor
                    const T* b = bS + slot;
                    const T tol = tolS[slot];
                    unsigned long long infeas = 0ull;
                    bool ok = true;
                    if (!dual) {
                        if (s > 0) ok = chol_solve<T>(s, idx, NT, Gs, KP, b, XS, fac, z, FS);
                        T xmax = (T)0;
                        for (int t = 0; t < s; ++t) xmax = fmax(xmax, fabs(z[t * FS]));
                        const T tolx = xmax * Eps<T>::v() * (T)8;
                        for (int t = 0; t < s; ++t)
                            if (z[t * FS] < -tolx) infeas |= 1ull << idx[t * NT];
                        unsigned long long A = (~F) & kmask;
                        while (A) {
                            const int j = __ffsll((long long)A) - 1;
                            A &= A - 1;
                            T y = -b[j * XS];
                            const T* Gj = Gs + j * KP;
                            for (int t = 0; t < s; ++t) y = fma(Gj[idx[t * NT]], z[t * FS], y);
                            if (y < -tol) infeas |= 1ull << j;
                        }
                    } else {
                        const T* u = uS + slot;
                        T* x = xS + slot;
                        if (s > 0) ok = chol_solve<T>(s, idx, NT, Ps, KP, u, XS, fac, z, FS);
                        // y_A = -z ; infeasible where y_A < 0  <=>  z > 0
                        for (int t = 0; t < s; ++t)
                            if (z[t * FS] > tol) infeas |= 1ull << idx[t * NT];
                        // x_F = u_F - P_FA z
                        T xmax = (T)0;
                        unsigned long long Fb = F;
                        while (Fb) {
                            const int f = __ffsll((long long)Fb) - 1;
                            Fb &= Fb - 1;
                            T xf = u[f * XS];
                            const T* Pf = Ps + f * KP;
                            for (int t = 0; t < s; ++t) xf = fma(-Pf[idx[t * NT]], z[t * FS], xf);
                            x[f * XS] = xf;
                            xmax = fmax(xmax, fabs(xf));
                        }
                        const T tolx = xmax * Eps<T>::v() * (T)8;
                        Fb = F;
                        while (Fb) {
                            const int f = __ffsll((long long)Fb) - 1;
                            Fb &= Fb - 1;
                            if (x[f * XS] < -tolx) infeas |= 1ull << f;
                        }
                    }
                    if (!ok) atomicAdd(&cnt->nnls_bad_pivot, 1ull);
                    const int st = stateS[slot];
                    int alpha = st & 0xff, beta = (st >> 8) & 0xff, iter = st >> 16;
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
                        maskS[slot] = F ^ flip;
                        stateS[slot] = alpha | (beta << 8) | ((iter + 1) << 16);
                        keep = true;
                        ++pivots_local;
                    } else {
                        if (infeas != 0ull) atomicAdd(&cnt->nnls_unconverged, 1ull);
                        if (want_hist) atomicAdd(&cnt->hist[hist_p ? (p < 39 ? p : 39) : (iter < 39 ? iter : 39)], 1ull);
                        // final x, expanded by factor id, over b's slot (b is no longer needed)
                        T* out = bS + slot;
                        if (!dual) {
                            for (int f = 0; f < KP; ++f) out[f * XS] = (T)0;
                            for (int t = 0; t < s; ++t) {
                                const T v = z[t * FS];
                                out[idx[t * NT] * XS] = v > (T)0 ? v : (T)0;
                            }
                        } else {
                            const T* x = xS + slot;
                            for (int f = 0; f < KP; ++f) {
                                T v = (T)0;
                                if ((F >> f) & 1ull) {
                                    v = x[f * XS];
                                    if (!(v > (T)0)) v = (T)0;
                                }
                                out[f * XS] = v;
                            }
                        }
                    }
                }
            }
            // ---- 4. order-preserving compaction of the slots that stay active ----
            {
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) wsum[warp] = __popc(bal);
                __syncthreads();
                int base = 0, total = 0;
                for (int w = 0; w < NW; ++w) {
                    const int cw = wsum[w];
                    if (w < warp) base += cw;
                    total += cw;
                }
                if (keep) nxt[base + __popc(bal & ((1u << lane) - 1u))] = sorted[tid];
                nact = total;
                __syncthreads();
                int* tmp = cur;
                cur = nxt;
                nxt = tmp;
            }
        }
        for (int e = tid; e < ncol * KP; e += NT) {  // coalesced store of the solution tile
            const int j = e / KP, f = e % KP;
            if (f < k) xout[(size_t)j * ld + f] = bS[f * XS + j];
        }
        if (tid < ncol) grp.mask[col0 + tid] = maskS[tid];
    }
    if (pivots_local) atomicAdd(&cnt->nnls_pivots, pivots_local);
}

NnlsConfig nnls_config(int k, int kp, bool is_double, int num_sms) {
    NnlsConfig c{};
    const size_t ts = is_double ? 8 : 4;
    if (kp == 32) {
        c.threads = is_double ? 64 : 128;
    } else {
        c.threads = is_double ? 32 : 64;
    }
    // pool: must hold at least one column at the worst-case size (primal-only fallback: s = k)
    const int worst = k * (k + 1) / 2 + k;
    const int half = (k / 2 + 1);
    const int typical = half * (half + 1) / 2 + half;            // dual/primal split: s <= k/2
    int pool = typical * c.threads / 2;                           // half the columns at the largest usual size
    if (pool < 2 * worst) pool = 2 * worst;
    const size_t fixed = nnls_layout(kp, c.threads, 0, ts).total;
    const size_t limit = 100 * 1024;                              // two CTAs per SM
    if (fixed + (size_t)pool * ts > limit && fixed + (size_t)2 * worst * ts <= limit)
        pool = (int)((limit - fixed) / ts);
    c.pool_elems = pool;
    c.smem = nnls_layout(kp, c.threads, pool, ts).total;
    const int per_sm = (int)((226 * 1024) / (c.smem + 1024));
    c.grid_cap = num_sms * (per_sm > 0 ? per_sm : 1);
    return c;
}

void launch_nnls(bool is_double, const NnlsGroup* groups, int ngroups, int total_blocks, int k, int kp,
                 const NnlsConfig& cfg, Counters* cnt, int flags, cudaStream_t st) {
    if (total_blocks <= 0) return;
    const int grid = total_blocks < cfg.grid_cap ? total_blocks : cfg.grid_cap;
    if (is_double) {
        if (kp == 32)
            k_nnls<double, 32><<<grid, cfg.threads, cfg.smem, st>>>(groups, ngroups, total_blocks, k, cfg.pool_elems, cnt, flags);
        else
            k_nnls<double, 64><<<grid, cfg.threads, cfg.smem, st>>>(groups, ngroups, total_blocks, k, cfg.pool_elems, cnt, flags);
    } else {
        if (kp == 32)
            k_nnls<float, 32><<<grid, cfg.threads, cfg.smem, st>>>(groups, ngroups, total_blocks, k, cfg.pool_elems, cnt, flags);
        else
            k_nnls<float, 64><<<grid, cfg.threads, cfg.smem, st>>>(groups, ngroups, total_blocks, k, cfg.pool_elems, cnt, flags);
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
