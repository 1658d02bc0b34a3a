// liger_b200 -- hot-path kernels of the iNMF / UINMF alternating-NNLS iteration (sm_100a).
//
//   k_prep        L~ = [W+V; U] (fp32 rows of KP), LtL = L~^T L~, VtV = V~^T V~          (R/cINMF.R:478-479)
//   k_nnls<T>     batched block-principal-pivoting NNLS on a shared Gram                    (RcppPlanc::bppnnls_prod)
//   k_spmm_xh     R = X H  (tiled CSR gather of H rows staged by TMA bulk copy) + G = H^T H (R/cINMF.R:486-488,497-499)
//   k_vrhs/k_wrhs right-hand sides of the V_i/U_i and W solves                              (R/cINMF.R:488,499)
//   k_objective   sum_i objErr_i from the sufficient statistics                             (src/objErr.cpp:19-29)
#include <cooperative_groups.h>

#include <mutex>

#include "lgr_common.cuh"
#include "lgr_ptx.cuh"

namespace lgr {

// ------------------------------------------------------------------------------------------------
// k_prep: L~ rows + partial Grams.  One CTA = PREP_ROWS_PER_BLOCK rows of one dataset.
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(512) k_prep(const DsDev* __restrict__ ds, int nds, const double* __restrict__ W, int k) {
    constexpr int RB = 2048 / KP;
    constexpr int PER = RB * KP / 512;   // entries per thread (4)
    __shared__ double Ls[RB][KP + 1];
    __shared__ double Vs[RB][KP + 1];
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::prep_blk_off);
    const DsDev D = ds[i];
    const int r0 = ((int)blockIdx.x - D.prep_blk_off) * RB;
    const int nr = min(RB, D.rows - r0);
    // all global loads of the thread first (they are independent: one memory latency instead of PER), then the arithmetic
    double lv[PER], vv[PER];
    float rsv[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = threadIdx.x + q * 512;
        const int r = e / KP, f = e % KP, gr = r0 + r;
        double l = 0.0, v = 0.0;
        float rs = 0.f;
        if (r < nr) {
            if (f < k) {
                v = D.V[(size_t)gr * KP + f];
                l = gr < D.m ? W[(size_t)gr * KP + f] : 0.0;
            }
            if (D.Asplit) rs = D.rs[gr];
        }
        lv[q] = l;
        vv[q] = v;
        rsv[q] = rs;
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = threadIdx.x + q * 512;
        const int r = e / KP, f = e % KP;
        const double v = vv[q], l = v + lv[q];
        if (r < nr) {
            const int gr = r0 + r;
            D.L[(size_t)gr * KP + f] = (float)l;
            if (D.Asplit) {
                // operand-slab image of k_dense_ltx: [64-gene stage][k-step of 16][chunk of 8 genes][row 4 f + t][gene % 8],
                // t = hi / mid / lo bf16 term of rs[g] * L~[g, f] (the three terms sum to the fp32 value exactly).  KP = 64: one
                // image per half of the factors (f / 32), asplit_half words apart -- the sweeps run once per half
                const float x = (float)l * rsv[q];
                const unsigned xb = __float_as_uint(x), hb = xb & 0xFFFF0000u;
                const float r1 = x - __uint_as_float(hb);
                const unsigned mb = __float_as_uint(r1) & 0xFFFF0000u;
                const float r2 = r1 - __uint_as_float(mb);
                unsigned short* img = D.Asplit + (size_t)(f >> 5) * D.asplit_half +
                                      ((size_t)(gr >> 6) * 16384 + ((gr >> 4) & 3) * 4096 + ((gr >> 3) & 1) * 2048 + (gr & 7) * 2) / 2;
                const int fl = f & 31;
                img[(4 * fl + 0) * 8] = (unsigned short)(hb >> 16);
                img[(4 * fl + 1) * 8] = (unsigned short)(mb >> 16);
                img[(4 * fl + 2) * 8] = (unsigned short)(__float_as_uint(r2) >> 16);
            }
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
        k_prep<32><<<total_blocks, 512, 0, st>>>(ds, nds, W, k);
    else
        k_prep<64><<<total_blocks, 512, 0, st>>>(ds, nds, W, k);
    LGR_CUDA(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------------
// right-hand sides of the V/U and W solves (fp64; tiny)
//   CtBv_i[g][:] = R_i[g][:] - G_i W[g][:]  (g < m),  R_i[g][:]  (unshared rows)      R/cINMF.R:488
//   CtBw[g][:]   = sum_i ( R_i[g][:] - G_i V_i[g][:] ),  Gsum = sum_i G_i               R/cINMF.R:497-499
// ------------------------------------------------------------------------------------------------
// One CTA = RHS_ROWS feature rows x KP factors; G_i and the W / V_i rows are staged in shared memory.
constexpr int RHS_ROWS = 8;

template <int KP>
__global__ void __launch_bounds__(RHS_ROWS* KP) k_vrhs(const DsDev* __restrict__ ds, const double* __restrict__ W, int k) {
    const DsDev& D = ds[blockIdx.y];
    __shared__ double Gs[KP][KP + 1];
    __shared__ double Ws[RHS_ROWS][KP];
    const int r0 = blockIdx.x * RHS_ROWS;
    if (r0 >= D.rows) return;
    const int f = threadIdx.x % KP, rl = threadIdx.x / KP, g = r0 + rl;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) Gs[e / KP][e % KP] = D.G[e];
    Ws[rl][f] = (g < D.m) ? W[(size_t)g * KP + f] : 0.0;
    __syncthreads();
    if (g >= D.rows) return;
    double v = 0.0;
    if (f < k) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int f2 = 0;
        for (; f2 + 3 < k; f2 += 4) {
            a0 = fma(Gs[f][f2], Ws[rl][f2], a0);
            a1 = fma(Gs[f][f2 + 1], Ws[rl][f2 + 1], a1);
            a2 = fma(Gs[f][f2 + 2], Ws[rl][f2 + 2], a2);
            a3 = fma(Gs[f][f2 + 3], Ws[rl][f2 + 3], a3);
        }
        for (; f2 < k; ++f2) a0 = fma(Gs[f][f2], Ws[rl][f2], a0);
        v = (double)D.R[(size_t)g * KP + f] - ((a0 + a1) + (a2 + a3));
    }
    D.CtBv[(size_t)g * KP + f] = v;
}

void launch_vrhs(const DsDev* ds, int nds, const double* W, int k, int kp, int max_rows, cudaStream_t st) {
    dim3 grid((unsigned)((max_rows + RHS_ROWS - 1) / RHS_ROWS), (unsigned)nds);
    if (kp == 32)
        k_vrhs<32><<<grid, RHS_ROWS * 32, 0, st>>>(ds, W, k);
    else
        k_vrhs<64><<<grid, RHS_ROWS * 64, 0, st>>>(ds, W, k);
    LGR_CUDA(cudaGetLastError());
}

template <int KP>
__global__ void __launch_bounds__(RHS_ROWS* KP) k_wrhs(const DsDev* __restrict__ ds, int nds, int m, double* __restrict__ CtBw,
                                                       double* __restrict__ Gsum, int k) {
    __shared__ double Gs[KP][KP + 1];
    __shared__ double Vs[RHS_ROWS][KP];
    const int r0 = blockIdx.x * RHS_ROWS;
    const int f = threadIdx.x % KP, rl = threadIdx.x / KP, g = r0 + rl;
    double v = 0.0;
    for (int i = 0; i < nds; ++i) {
        const DsDev& D = ds[i];
        __syncthreads();
        for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) Gs[e / KP][e % KP] = D.G[e];
        Vs[rl][f] = (g < m) ? D.V[(size_t)g * KP + f] : 0.0;
        __syncthreads();
        if (g < m && f < k) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int f2 = 0;
            for (; f2 + 3 < k; f2 += 4) {
                a0 = fma(Gs[f][f2], Vs[rl][f2], a0);
                a1 = fma(Gs[f][f2 + 1], Vs[rl][f2 + 1], a1);
                a2 = fma(Gs[f][f2 + 2], Vs[rl][f2 + 2], a2);
                a3 = fma(Gs[f][f2 + 3], Vs[rl][f2 + 3], a3);
            }
            for (; f2 < k; ++f2) a0 = fma(Gs[f][f2], Vs[rl][f2], a0);
            v += (double)D.R[(size_t)g * KP + f] - ((a0 + a1) + (a2 + a3));
        }
        if (blockIdx.x == 0) {  // Gsum = sum_i G_i (accumulated by block 0 while it walks the datasets)
            for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
                const double prev = (i == 0) ? 0.0 : Gsum[e];
                Gsum[e] = prev + Gs[e / KP][e % KP];
            }
        }
    }
    if (g < m) CtBw[(size_t)g * KP + f] = (f < k) ? v : 0.0;
}

void launch_wrhs(const DsDev* ds, int nds, int m, double* CtBw, double* Gsum, int k, int kp, cudaStream_t st) {
    const unsigned grid = (unsigned)((m + RHS_ROWS - 1) / RHS_ROWS);
    if (kp == 32)
        k_wrhs<32><<<grid, RHS_ROWS * 32, 0, st>>>(ds, nds, m, CtBw, Gsum, k);
    else
        k_wrhs<64><<<grid, RHS_ROWS * 64, 0, st>>>(ds, nds, m, CtBw, Gsum, k);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// k_objective: obj += sum_i [ ||X_i||^2 - 2 <L~_i, R_i> + tr(LtL_i G_i) + lambda_i tr(VtV_i G_i) ]
//   (src/objErr.cpp:19-29 with Tr(H E^T L) rewritten as <L, E H^T> = <L, R>; the unshared block of UINMF
//    is covered by the stacked rows.)  LtL/VtV must be those of the CURRENT W, V (k_prep ran after the solves).
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_objective(const DsDev* __restrict__ ds, int k, double* obj) {
    const DsDev& D = ds[blockIdx.y];
    double acc = 0.0;
    const size_t tot = (size_t)D.rows * KP;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x)
        acc += (double)D.L[e] * (double)D.R[e];
    acc *= -2.0;
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
            const int a = e / KP, b = e % KP;
            if (a < k && b < k) acc += (D.LtL[a * KP + b] + D.lambda * D.VtV[a * KP + b]) * D.G[b * KP + a];
        }
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(obj, sh[0] + (blockIdx.x == 0 ? D.sqnorm : 0.0));
}

void launch_objective(const DsDev* ds, int nds, int k, int kp, double* obj_slot, cudaStream_t st) {
    dim3 grid(16, (unsigned)nds);
    if (kp == 32)
        k_objective<32><<<grid, 256, 0, st>>>(ds, k, obj_slot);
    else
        k_objective<64><<<grid, 256, 0, st>>>(ds, k, obj_slot);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// One all-reduce per iteration instead of two: the fp64 Grams G_i ride in the tail of the fp32 statistics arena as (hi, lo)
// float pairs (G = hi + lo to ~2^-48; the cross-rank sums of the two halves are fp32 sums of at most `world` terms).
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_gram(const double* __restrict__ G, float* __restrict__ tail, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double g = G[e];
    const float hi = (float)g;
    tail[e] = hi;
    tail[n + e] = (float)(g - (double)hi);
}
__global__ void k_unpack_gram(double* __restrict__ G, const float* __restrict__ tail, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) G[e] = (double)tail[e] + (double)tail[n + e];
}
void launch_pack_gram(const double* G, float* tail, int n, cudaStream_t st) {
    k_pack_gram<<<(n + 255) / 256, 256, 0, st>>>(G, tail, n);
    LGR_CUDA(cudaGetLastError());
}
void launch_unpack_gram(double* G, const float* tail, int n, cudaStream_t st) {
    k_unpack_gram<<<(n + 255) / 256, 256, 0, st>>>(G, tail, n);
    LGR_CUDA(cudaGetLastError());
}

void kernels_init() {
    // cudaFuncSetAttribute is per device: remember which devices have been initialised
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    LGR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    nnls_init();
    nnls_reg_init();
    nnls_wbig_init();
    nnls_scol_init();
    nnls_pair_init();
    spmm_init();
    tc_init();
    dense_init();
    if (dev >= 0 && dev < 64) done[dev] = true;
}

}  // namespace lgr
