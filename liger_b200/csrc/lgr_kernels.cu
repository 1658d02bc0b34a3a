// liger_b200 -- hot-path kernels of the iNMF / UINMF alternating-NNLS iteration (sm_100a).
//
//   k_prep        L~ = [W+V; U] (fp32 rows of KP), LtL = L~^T L~, VtV = V~^T V~          (R/cINMF.R:478-479)
//   k_spmm_ltx    B = L~^T X     CSC gather of L rows, quarter-warp per non-zero            (R/cINMF.R:480)
//   k_nnls<T>     batched block-principal-pivoting NNLS on a shared Gram                    (RcppPlanc::bppnnls_prod)
//   k_spmm_xh     R = X H  (tiled CSR gather of H rows staged by TMA bulk copy) + G = H^T H (R/cINMF.R:486-488,497-499)
//   k_vrhs/k_wrhs right-hand sides of the V_i/U_i and W solves                              (R/cINMF.R:488,499)
//   k_objective   sum_i objErr_i from the sufficient statistics                             (src/objErr.cpp:19-29)
#include <cooperative_groups.h>

#include "lgr_common.cuh"

namespace lgr {

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ float4 ldg_nc4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// streaming loads for the sparse arrays: read once, do not pollute L1
__device__ __forceinline__ int ldg_stream(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned short ldg_stream(const unsigned short* p) {
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <typename DS>
__device__ __forceinline__ int find_ds(const DS* ds, int nds, int blk, int DS::*off) {
    int i = 0;
    while (i + 1 < nds && blk >= ds[i + 1].*off) ++i;
    return i;
}

// ------------------------------------------------------------------------------------------------
// k_prep: L~ rows + partial Grams.  One CTA = PREP_ROWS_PER_BLOCK rows of one dataset.
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_prep(const DsDev* __restrict__ ds, int nds, const double* __restrict__ W, int k) {
    constexpr int RB = 2048 / KP;
    __shared__ double Ls[RB][KP + 1];
    __shared__ double Vs[RB][KP + 1];
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::prep_blk_off);
    const DsDev D = ds[i];
    const int r0 = ((int)blockIdx.x - D.prep_blk_off) * RB;
    const int nr = min(RB, D.rows - r0);
    for (int e = threadIdx.x; e < RB * KP; e += blockDim.x) {
        const int r = e / KP, f = e % KP;
        double l = 0.0, v = 0.0;
        if (r < nr) {
            const int gr = r0 + r;
            if (f < k) {
                v = D.V[(size_t)gr * KP + f];
                l = v + (gr < D.m ? W[(size_t)gr * KP + f] : 0.0);
            }
            D.L[(size_t)gr * KP + f] = (float)l;
        }
        Ls[r][f] = l;
        Vs[r][f] = v;
    }
    __syncthreads();
    const int npairs = k * (k + 1) / 2;
    for (int pr = threadIdx.x; pr < npairs; pr += blockDim.x) {
        int a = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (a * (a + 1) / 2 > pr) --a;
        while ((a + 1) * (a + 2) / 2 <= pr) ++a;
        const int b = pr - a * (a + 1) / 2;  // b <= a
        double sl = 0.0, sv = 0.0;
        for (int r = 0; r < nr; ++r) {
            sl = fma(Ls[r][a], Ls[r][b], sl);
            sv = fma(Vs[r][a], Vs[r][b], sv);
        }
        atomicAdd(&D.LtL[a * KP + b], sl);
        atomicAdd(&D.VtV[a * KP + b], sv);
        if (a != b) {
            atomicAdd(&D.LtL[b * KP + a], sl);
            atomicAdd(&D.VtV[b * KP + a], sv);
        }
    }
}

void launch_prep(const DsDev* ds, int nds, int total_blocks, const double* W, int k, int kp, cudaStream_t st) {
    if (kp == 32)
        k_prep<32><<<total_blocks, 256, 0, st>>>(ds, nds, W, k);
    else
        k_prep<64><<<total_blocks, 256, 0, st>>>(ds, nds, W, k);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// k_spmm_ltx: B[c][:] = sum_{g in nz(c)} x_gc * L[g][:]      (CSC gather; warp per cell)
//   KP/4 lanes cover one row of L with one 128-bit load each, so one warp instruction serves
//   32/(KP/4) non-zeros; (row, value) pairs are loaded coalesced, one per lane, and broadcast by shuffle.
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_spmm_ltx(const DsDev* __restrict__ ds, int nds) {
    constexpr int LPN = KP / 4;    // lanes per non-zero
    constexpr int NPI = 32 / LPN;  // non-zeros per warp instruction
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::ltx_blk_off);
    const DsDev& D = ds[i];
    const int n = D.n;
    const int* __restrict__ colptr = D.colptr;
    const int* __restrict__ rowidx = D.rowidx;
    const float* __restrict__ val = D.val;
    const float4* __restrict__ L4 = reinterpret_cast<const float4*>(D.L);
    float4* __restrict__ B4 = reinterpret_cast<float4*>(D.B);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int q = lane / LPN, s = lane % LPN;
    const int cell0 = ((int)blockIdx.x - D.ltx_blk_off) * LTX_CELLS_PER_BLOCK;
    const int cell1 = min(cell0 + LTX_CELLS_PER_BLOCK, n);
    for (int c = cell0 + warp; c < cell1; c += nwarps) {
        const int beg = __ldg(colptr + c), end = __ldg(colptr + c + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = beg; base < end; base += 32) {
            const int e = base + lane;
            int r = 0;
            float v = 0.f;
            if (e < end) {
                r = ldg_stream(rowidx + e);
                v = ldg_stream(val + e);
            }
            const int cnt = min(32, end - base);
#pragma unroll
            for (int stp = 0; stp < LPN; ++stp) {
                const int t = stp * NPI + q;
                const int rr = __shfl_sync(0xffffffffu, r, t);
                const float vv = __shfl_sync(0xffffffffu, v, t);
                if (stp * NPI < cnt) {  // warp-uniform
                    const float4 l = ldg_nc4(L4 + (size_t)rr * LPN + s);
                    acc.x = fmaf(vv, l.x, acc.x);
                    acc.y = fmaf(vv, l.y, acc.y);
                    acc.z = fmaf(vv, l.z, acc.z);
                    acc.w = fmaf(vv, l.w, acc.w);
                }
            }
        }
#pragma unroll
        for (int off = LPN; off < 32; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
        }
        if (q == 0) B4[(size_t)c * LPN + s] = acc;
    }
}

void launch_spmm_ltx(const DsDev* ds, int nds, int total_blocks, int kp, cudaStream_t st) {
    if (total_blocks <= 0) return;
    if (kp == 32)
        k_spmm_ltx<32><<<total_blocks, 256, 0, st>>>(ds, nds);
    else
        k_spmm_ltx<64><<<total_blocks, 256, 0, st>>>(ds, nds);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// k_spmm_xh: one CTA = one cell tile (TC cells) of one dataset.
//   phase 0: the tile's H rows (TC x KP fp32, contiguous in HBM) are staged in shared memory by one
//            TMA bulk copy (cp.async.bulk + mbarrier).
//   phase 1: warp per gene: R[g][:] += sum_{c in tile, x_gc != 0} x_gc * Hs[c][:]   (gather from smem,
//            quarter-warp per non-zero), one vector RED (red.global.add.v4.f32) per (gene, tile).
//   phase 2: G += Hs^T Hs  (4x4 register patches over the upper triangle, fp64 atomics per CTA).
// ------------------------------------------------------------------------------------------------
constexpr int XH_THREADS = 512;

size_t spmm_xh_smem(int kp, int tc) {
    const int np = kp / 4;
    const int npatch = np * (np + 1) / 2;
    const int ncg = XH_THREADS / npatch > 0 ? XH_THREADS / npatch : 1;
    return (size_t)tc * kp * sizeof(float) + (size_t)ncg * npatch * 16 * sizeof(float) + 64;
}

template <int KP>
__global__ void __launch_bounds__(XH_THREADS) k_spmm_xh(const DsDev* __restrict__ ds, int nds, int k, int TC) {
    constexpr int LPN = KP / 4;
    constexpr int NPI = 32 / LPN;
    constexpr int NP = KP / 4;                 // 4-wide patches per side
    constexpr int NPATCH = NP * (NP + 1) / 2;  // upper-triangular patches
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Hs = reinterpret_cast<float*>(smem_raw);
    float* red = Hs + (size_t)TC * KP;
    const int ncg = (XH_THREADS / NPATCH) > 0 ? (XH_THREADS / NPATCH) : 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + (size_t)ncg * NPATCH * 16);

    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::xh_blk_off);
    const DsDev& D = ds[i];
    const int tile = (int)blockIdx.x - D.xh_blk_off;
    const int c0 = tile * TC;
    const int ncell = min(TC, D.n - c0);
    const int rows = D.rows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)ncell * KP * sizeof(float);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(D.H + (size_t)c0 * KP);
        char* dst = reinterpret_cast<char*>(Hs);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t nb = min(32768u, bytes - off);
            bulk_g2s(dst + off, src + off, nb, bar);
        }
    }
    mbar_wait(bar, 0);

    // ---- phase 1: R += X_tile * Hs ----
    {
        const int q = lane / LPN, s = lane % LPN;
        const int* __restrict__ tptr = D.tptr + (size_t)tile * rows;
        const unsigned short* __restrict__ tcell = D.tcell;
        const float* __restrict__ tval = D.tval;
        const float4* Hs4 = reinterpret_cast<const float4*>(Hs);
        for (int g0 = warp * 32; g0 < rows; g0 += nwarps * 32) {
            // lane holds the segment bounds of gene g0 + lane (coalesced), broadcast per gene below
            const int gl = g0 + lane;
            int pb = 0, pe = 0;
            if (gl < rows) {
                pb = __ldg(tptr + gl);
                pe = __ldg(tptr + gl + 1);
            }
            const int ng = min(32, rows - g0);
            for (int gg = 0; gg < ng; ++gg) {
                const int beg = __shfl_sync(0xffffffffu, pb, gg);
                const int end = __shfl_sync(0xffffffffu, pe, gg);
                if (beg == end) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int base = beg; base < end; base += 32) {
                    const int e = base + lane;
                    int cc = 0;
                    float v = 0.f;
                    if (e < end) {
                        cc = (int)ldg_stream(tcell + e);
                        v = ldg_stream(tval + e);
                    }
                    const int cnt = min(32, end - base);
#pragma unroll
                    for (int stp = 0; stp < LPN; ++stp) {
                        const int t = stp * NPI + q;
                        const int c = __shfl_sync(0xffffffffu, cc, t);
                        const float vv = __shfl_sync(0xffffffffu, v, t);
                        if (stp * NPI < cnt) {
                            const float4 h = Hs4[c * LPN + s];
                            acc.x = fmaf(vv, h.x, acc.x);
                            acc.y = fmaf(vv, h.y, acc.y);
                            acc.z = fmaf(vv, h.z, acc.z);
                            acc.w = fmaf(vv, h.w, acc.w);
                        }
                    }
                }
#pragma unroll
                for (int off = LPN; off < 32; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
                    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                }
                if (q == 0) red_add_v4(D.R + (size_t)(g0 + gg) * KP + s * 4, acc);
            }
        }
    }

    // ---- phase 2: G += Hs^T Hs over this tile ----
    {
        const int t = threadIdx.x;
        const int cg = t / NPATCH, pt = t % NPATCH;
        int pi = 0, pj = 0;
        {  // decode upper-triangular patch index: pt -> (pi <= pj)
            int rem = pt;
            while (rem >= NP - pi) {
                rem -= NP - pi;
                ++pi;
            }
            pj = pi + rem;
        }
        float a[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) a[x][y] = 0.f;
        const float4* Hs4 = reinterpret_cast<const float4*>(Hs);
        if (cg < ncg) {
            for (int c = cg; c < ncell; c += ncg) {
                const float4 u = Hs4[c * NP + pi];
                const float4 v = Hs4[c * NP + pj];
                const float uu[4] = {u.x, u.y, u.z, u.w};
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) a[x][y] = fmaf(uu[x], vv[y], a[x][y]);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) red[((size_t)cg * NPATCH + pt) * 16 + x * 4 + y] = a[x][y];
        }
        __syncthreads();
        // reduce over cell groups, one thread per (patch, element); fp64 atomics into G (symmetric fill)
        for (int e = t; e < NPATCH * 16; e += blockDim.x) {
            const int p = e / 16, xy = e % 16, x = xy / 4, y = xy % 4;
            double sum = 0.0;
            for (int g = 0; g < ncg; ++g) sum += (double)red[((size_t)g * NPATCH + p) * 16 + xy];
            int ppi = 0, rem = p;
            while (rem >= NP - ppi) {
                rem -= NP - ppi;
                ++ppi;
            }
            const int ppj = ppi + rem;
            const int fa = ppi * 4 + x, fb = ppj * 4 + y;
            if (fa < k && fb < k) {
                if (ppi == ppj) {
                    atomicAdd(&D.G[fa * KP + fb], sum);  // diagonal patch holds the full 4x4 block
                } else {
                    atomicAdd(&D.G[fa * KP + fb], sum);
                    atomicAdd(&D.G[fb * KP + fa], sum);
                }
            }
        }
    }
}

void launch_spmm_xh(const DsDev* ds, int nds, int total_tiles, int k, int kp, int tc, cudaStream_t st) {
    if (total_tiles <= 0) return;
    const size_t smem = spmm_xh_smem(kp, tc);
    if (kp == 32)
        k_spmm_xh<32><<<total_tiles, XH_THREADS, smem, st>>>(ds, nds, k, tc);
    else
        k_spmm_xh<64><<<total_tiles, XH_THREADS, smem, st>>>(ds, nds, k, tc);
    LGR_CUDA(cudaGetLastError());
}
// ------------------------------------------------------------------------------------------------
// right-hand sides of the V/U and W solves (fp64; tiny)
//   CtBv_i[g][:] = R_i[g][:] - G_i W[g][:]  (g < m),  R_i[g][:]  (unshared rows)      R/cINMF.R:488
//   CtBw[g][:]   = sum_i ( R_i[g][:] - G_i V_i[g][:] ),  Gsum = sum_i G_i               R/cINMF.R:497-499
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void k_vrhs(const DsDev* __restrict__ ds, const double* __restrict__ W, int k) {
    const DsDev& D = ds[blockIdx.y];
    __shared__ double Gs[KP * KP];
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) Gs[e] = D.G[e];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = idx / KP, f = idx % KP;
    if (g >= D.rows) return;
    double v = 0.0;
    if (f < k) {
        v = (double)D.R[(size_t)g * KP + f];
        if (g < D.m) {
            const double* w = W + (size_t)g * KP;
            double acc = 0.0;
            for (int f2 = 0; f2 < k; ++f2) acc = fma(Gs[f * KP + f2], w[f2], acc);
            v -= acc;
        }
    }
    D.CtBv[(size_t)g * KP + f] = v;
}

void launch_vrhs(const DsDev* ds, int nds, const double* W, int k, int kp, int max_rows, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)max_rows * kp + 255) / 256), (unsigned)nds);
    if (kp == 32)
        k_vrhs<32><<<grid, 256, 0, st>>>(ds, W, k);
    else
        k_vrhs<64><<<grid, 256, 0, st>>>(ds, W, k);
    LGR_CUDA(cudaGetLastError());
}

template <int KP>
__global__ void k_wrhs(const DsDev* __restrict__ ds, int nds, int m, double* __restrict__ CtBw, double* __restrict__ Gsum,
                       int k) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
            double s = 0.0;
            for (int i = 0; i < nds; ++i) s += ds[i].G[e];
            Gsum[e] = s;
        }
    }
    const int g = idx / KP, f = idx % KP;
    if (g >= m) return;
    double v = 0.0;
    if (f < k) {
        for (int i = 0; i < nds; ++i) {
            const DsDev& D = ds[i];
            const double* vi = D.V + (size_t)g * KP;
            const double* Gi = D.G + (size_t)f * KP;
            double acc = 0.0;
            for (int f2 = 0; f2 < k; ++f2) acc = fma(Gi[f2], vi[f2], acc);
            v += (double)D.R[(size_t)g * KP + f] - acc;
        }
    }
    CtBw[(size_t)g * KP + f] = v;
}

void launch_wrhs(const DsDev* ds, int nds, int m, double* CtBw, double* Gsum, int k, int kp, cudaStream_t st) {
    const unsigned grid = (unsigned)(((size_t)m * kp + 255) / 256);
    if (kp == 32)
        k_wrhs<32><<<grid, 256, 0, st>>>(ds, nds, m, CtBw, Gsum, k);
    else
        k_wrhs<64><<<grid, 256, 0, st>>>(ds, nds, m, CtBw, Gsum, k);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// k_objective: obj += sum_i [ ||X_i||^2 - 2 <L~_i, R_i> + tr(LtL_i G_i) + lambda_i tr(VtV_i G_i) ]
//   (src/objErr.cpp:19-29 with Tr(H E^T L) rewritten as <L, E H^T> = <L, R>; the unshared block of UINMF
//    is covered by the stacked rows.)  LtL/VtV must be those of the CURRENT W, V (k_prep ran after the solves).
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_objective(const DsDev* __restrict__ ds, int k, double* obj) {
    const DsDev& D = ds[blockIdx.x];
    double acc = 0.0;
    const size_t tot = (size_t)D.rows * KP;
    for (size_t e = threadIdx.x; e < tot; e += blockDim.x) acc += (double)D.L[e] * (double)D.R[e];
    acc *= -2.0;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
        const int a = e / KP, b = e % KP;
        if (a < k && b < k) acc += (D.LtL[a * KP + b] + D.lambda * D.VtV[a * KP + b]) * D.G[b * KP + a];
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(obj, sh[0] + D.sqnorm);
}

void launch_objective(const DsDev* ds, int nds, int k, int kp, double* obj_slot, cudaStream_t st) {
    if (kp == 32)
        k_objective<32><<<nds, 256, 0, st>>>(ds, k, obj_slot);
    else
        k_objective<64><<<nds, 256, 0, st>>>(ds, k, obj_slot);
    LGR_CUDA(cudaGetLastError());
}

void kernels_init() {
    nnls_init();
    const int maxs = 227 * 1024 - 1024;  // static smem of the kernels counts against the 227 KB limit
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_xh<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_xh<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
}

}  // namespace lgr
