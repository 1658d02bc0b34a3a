// liger_b200 -- kernels of online iNMF (front end R/integration.R:722-933, which delegates
// the arithmetic to RcppPlanc::onlineINMF -- not under the reference tree; restated from the published algorithm, Gao et al.
// 2021, Algorithm 1, with the bookkeeping of rliger 1.0's online_iNMF: see DESIGN.md section 8.  PARITY UNPINNED).
//
// One minibatch iteration =
//   k_online_gram      L~_i = W + V_i (fp32 rows of 32) and the two Grams L~^T L~, V^T V (fp64)      -> k_gram_inv (lgr_nnls.cu)
//   k_online_ltx       B[c, :] = sum_g X[g, cell_c] L~[g, :] for the minibatch's cells (warp per cell, lane = factor)
//   k_nnls_reg         the H solve of the minibatch (lgr_nnls_reg.cu, unchanged)
//   k_online_stats_a   A_i <- s A_i + H_b^T H_b / b_i            (k x k, fp64)
//   k_online_scale_all / k_online_stats_b_all   B_i <- s B_i + X_b H_b / b_i      (m x k, fp64 atomics)
//   k_online_hals      one HALS sweep over the columns of W, then of every V_i (warp per gene row, lane = factor: the
//                      column updates of a row depend only on that row, so rows are independent and the sweep over
//                      columns is sequential inside a warp)
// The minibatches are short (5000 cells by default), so these are plain gather kernels over the fp32 CSC copy of X: the
// tiled layouts of the batch path would have to be rebuilt per minibatch.
#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int OKP = 32;

// grid (d), 1024 threads: thread (a, b) of the 32 x 32 Grams; also writes L~_i
__global__ void __launch_bounds__(1024) k_online_gram(const double* __restrict__ W, const double* const* __restrict__ V, int m, int k,
                                                      float* const* __restrict__ Lt, double* const* __restrict__ LtL,
                                                      double* const* __restrict__ VtV) {
    const int i = blockIdx.x, a = threadIdx.x >> 5, b = threadIdx.x & 31;
    const double* Vi = V[i];
    float* L = Lt[i];
    for (int e = threadIdx.x; e < m * OKP; e += blockDim.x) {
        const int g = e >> 5, f = e & 31;
        L[e] = f < k ? (float)(W[(size_t)g * OKP + f] + Vi[(size_t)g * OKP + f]) : 0.f;
    }
    double sl = 0.0, sv = 0.0;
    if (a < k && b < k) {
        for (int g = 0; g < m; ++g) {
            const double va = Vi[(size_t)g * OKP + a], vb = Vi[(size_t)g * OKP + b];
            const double la = W[(size_t)g * OKP + a] + va, lb = W[(size_t)g * OKP + b] + vb;
            sl = fma(la, lb, sl);
            sv = fma(va, vb, sv);
        }
    }
    LtL[i][a * OKP + b] = sl;
    VtV[i][a * OKP + b] = sv;
}

// warp per minibatch cell; cells[c] = column of X (null: the identity, for the final solve over all cells)
__global__ void __launch_bounds__(256) k_online_ltx(const int* __restrict__ colptr, const int* __restrict__ rowidx, const float* __restrict__ val,
                                                    const int* __restrict__ cells, int nb, const float* __restrict__ Lt, float* __restrict__ B) {
    const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= nb) return;
    const int c = cells ? cells[w] : w;
    const int p0 = colptr[c], p1 = colptr[c + 1];
    float acc = 0.f;
    for (int p = p0; p < p1; p += 32) {
        const int q = p + lane;
        const int r = q < p1 ? rowidx[q] : 0;
        const float v = q < p1 ? val[q] : 0.f;
        const int cnt = min(32, p1 - p);
        for (int t = 0; t < cnt; ++t) {
            const int rr = __shfl_sync(0xffffffffu, r, t);
            const float vv = __shfl_sync(0xffffffffu, v, t);
            acc = fmaf(vv, Lt[(size_t)rr * OKP + lane], acc);
        }
    }
    B[(size_t)w * OKP + lane] = acc;
}

// grid (d), 1024 threads: A_i <- s A_i + H^T H / b  (H: nb x 32 fp32), zero diagonal entries lifted to 1e-15
__global__ void __launch_bounds__(1024) k_online_stats_a(const OnlineDs* __restrict__ ds, int k, double s) {
    const OnlineDs D = ds[blockIdx.x];
    const int a = threadIdx.x >> 5, b = threadIdx.x & 31;
    double acc = 0.0;
    if (a < k && b < k)
        for (int c = 0; c < D.nb; ++c) acc = fma((double)D.H[(size_t)c * OKP + a], (double)D.H[(size_t)c * OKP + b], acc);
    double v = s * D.A[a * OKP + b] + acc / (double)D.nb;
    if (a == b && a < k && v == 0.0) v = 1e-15;
    D.A[a * OKP + b] = (a < k && b < k) ? v : 0.0;
}
// HALS: warp per gene row, lane l = factor l.  A_i are staged in shared memory with stride 33.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
constexpr int HALS_MAX_D = 16;   // datasets whose V row lives in registers (more: several passes would be needed)
__global__ void __launch_bounds__(256) k_online_hals(double* __restrict__ W, const HalsDs* __restrict__ ds, int d, int m, int k,
                                                     double lambda, int sweeps, int first_v) {
    extern __shared__ double As[];   // d x 32 x 33
    for (int e = threadIdx.x; e < d * OKP * OKP; e += blockDim.x) {
        const int i = e / (OKP * OKP), r = (e / OKP) % OKP, c = e % OKP;
        As[(size_t)i * OKP * 33 + r * 33 + c] = ds[i].A[r * OKP + c];
    }
    __syncthreads();
    const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g >= m) return;
    const bool live = lane < k;
    double w = live ? W[(size_t)g * OKP + lane] : 0.0;
    double v[HALS_MAX_D], b[HALS_MAX_D];
#pragma unroll
    for (int i = 0; i < HALS_MAX_D; ++i) {
        v[i] = (i < d && live) ? ds[i].V[(size_t)g * OKP + lane] : 0.0;
        b[i] = (i < d && live) ? ds[i].B[(size_t)g * OKP + lane] : 0.0;
    }
    for (int sw = 0; sw < sweeps; ++sw) {
        for (int j = 0; j < k; ++j) {          // W[, j]
            double p = 0.0, den = 0.0, bs = 0.0;
#pragma unroll
            for (int i = 0; i < HALS_MAX_D; ++i) {
                if (i < d) {
                    const double* Ai = As + (size_t)i * OKP * 33;
                    p = fma(w + v[i], Ai[lane * 33 + j], p);     // ((W + V_i) A_i[, j])_g, this lane's term
                    den += Ai[j * 33 + j];
                    bs += b[i];
                }
            }
            p = warp_sum(live ? p : 0.0);
            if (lane == j) w = fmax(w + (bs - p) / den, 1e-16);
        }
        for (int j = 0; j < k; ++j) {          // V_i[, j]
#pragma unroll
            for (int i = 0; i < HALS_MAX_D; ++i) {
                if (i >= first_v && i < d) {   // (scenario 2: the V_i of the already factorised datasets stay as they are)
                    const double* Ai = As + (size_t)i * OKP * 33;
                    double p = live ? (w + (1.0 + lambda) * v[i]) * Ai[lane * 33 + j] : 0.0;
                    p = warp_sum(p);
                    if (lane == j) v[i] = fmax(v[i] + (b[i] - p) / ((1.0 + lambda) * Ai[j * 33 + j]), 1e-16);
                }
            }
        }
    }
    if (live) {
        W[(size_t)g * OKP + lane] = w;
#pragma unroll
        for (int i = 0; i < HALS_MAX_D; ++i)
            if (i >= first_v && i < d) ds[i].V[(size_t)g * OKP + lane] = v[i];
    }
}

// ---- the minibatch iteration as ONE replayable launch sequence: the iteration number lives on the device (*t_dev, 1-based;
//      k_online_next increments it), every kernel derives the minibatch's cells (sched + (t - 1) * mb_total) and the running-
//      average weight s_t from it, and one launch covers all datasets -- so the ten launches of an iteration are captured
//      into a CUDA graph once and replayed ----
__device__ __forceinline__ double online_weight(int t) { return t == 1 ? 0.0 : (t == 2 ? 1.0 : (double)(t - 2) / (double)(t - 1)); }
__device__ __forceinline__ int online_find_ds(const OnlineX* __restrict__ xs, int d, int w) {
    int i = 0;
    while (i + 1 < d && w >= xs[i + 1].off) ++i;
    return i;
}
__global__ void __launch_bounds__(256) k_online_ltx_all(const OnlineX* __restrict__ xs, int d, const int* __restrict__ sched, int mb_total,
                                                        const int* __restrict__ t_dev, float* __restrict__ B) {
    const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= mb_total) return;
    const OnlineX X = xs[online_find_ds(xs, d, w)];
    const int c = sched[(size_t)(*t_dev - 1) * mb_total + w];
    const int p0 = X.colptr[c], p1 = X.colptr[c + 1];
    float acc = 0.f;
    for (int p = p0; p < p1; p += 32) {
        const int q = p + lane;
        const int r = q < p1 ? X.rowidx[q] : 0;
        const float v = q < p1 ? X.val[q] : 0.f;
        const int cnt = min(32, p1 - p);
        for (int t = 0; t < cnt; ++t) {
            const int rr = __shfl_sync(0xffffffffu, r, t);
            const float vv = __shfl_sync(0xffffffffu, v, t);
            acc = fmaf(vv, X.Lt[(size_t)rr * OKP + lane], acc);
        }
    }
    B[(size_t)w * OKP + lane] = acc;
}
__global__ void __launch_bounds__(1024) k_online_stats_a_t(const OnlineDs* __restrict__ ds, int k, const int* __restrict__ t_dev) {
    const OnlineDs D = ds[blockIdx.x];
    const double s = online_weight(*t_dev);
    const int a = threadIdx.x >> 5, b = threadIdx.x & 31;
    double acc = 0.0;
    if (a < k && b < k)
        for (int c = 0; c < D.nb; ++c) acc = fma((double)D.H[(size_t)c * OKP + a], (double)D.H[(size_t)c * OKP + b], acc);
    double v = s * D.A[a * OKP + b] + acc / (double)D.nb;
    if (a == b && a < k && v == 0.0) v = 1e-15;
    D.A[a * OKP + b] = (a < k && b < k) ? v : 0.0;
}
__global__ void k_online_scale_all(const OnlineX* __restrict__ xs, int m, const int* __restrict__ t_dev) {
    const double s = online_weight(*t_dev);
    double* x = xs[blockIdx.y].Bstat;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)m * OKP) x[i] *= s;
}
__global__ void __launch_bounds__(256) k_online_stats_b_all(const OnlineX* __restrict__ xs, int d, const int* __restrict__ sched, int mb_total,
                                                            const int* __restrict__ t_dev, const float* __restrict__ H) {
    const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= mb_total) return;
    const OnlineX X = xs[online_find_ds(xs, d, w)];
    const int c = sched[(size_t)(*t_dev - 1) * mb_total + w];
    const double h = (double)H[(size_t)w * OKP + lane] * X.inv_b;
    if (h == 0.0) return;
    const int p0 = X.colptr[c], p1 = X.colptr[c + 1];
    for (int p = p0; p < p1; ++p) atomicAdd(X.Bstat + (size_t)X.rowidx[p] * OKP + lane, (double)X.val[p] * h);
}
__global__ void k_online_next(int* t_dev) { *t_dev += 1; }

// objective pieces over all cells of one dataset: cross = sum_c H[c, :] . B[c, :]   (B = L~^T X)
__global__ void k_online_cross(const float* __restrict__ H, const float* __restrict__ B, size_t n, double* __restrict__ out) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += (double)H[i] * (double)B[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
__global__ void k_online_sumsq(const float* __restrict__ x, size_t n, double* __restrict__ out) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += (double)x[i] * (double)x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace

void launch_online_gram(const double* W, const double* const* V, int d, int m, int k, float* const* Lt, double* const* LtL,
                        double* const* VtV, cudaStream_t st) {
    k_online_gram<<<d, 1024, 0, st>>>(W, V, m, k, Lt, LtL, VtV);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_ltx(const int* colptr, const int* rowidx, const float* val, const int* cells, int nb, const float* Lt, float* B,
                       cudaStream_t st) {
    if (nb <= 0) return;
    k_online_ltx<<<(unsigned)(((size_t)nb * 32 + 255) / 256), 256, 0, st>>>(colptr, rowidx, val, cells, nb, Lt, B);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_stats_a(const OnlineDs* ds, int d, int k, double s, cudaStream_t st) {
    k_online_stats_a<<<d, 1024, 0, st>>>(ds, k, s);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_hals(double* W, const HalsDs* ds, int d, int m, int k, double lambda, int sweeps, int first_v, cudaStream_t st) {
    LGR_REQUIRE(d <= HALS_MAX_D, "online iNMF: at most 16 datasets");
    const size_t smem = (size_t)d * OKP * 33 * sizeof(double);
    static bool attr = false;
    if (!attr) {
        LGR_CUDA(cudaFuncSetAttribute(k_online_hals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)HALS_MAX_D * OKP * 33 * sizeof(double))));
        attr = true;
    }
    k_online_hals<<<(unsigned)(((size_t)m * 32 + 255) / 256), 256, smem, st>>>(W, ds, d, m, k, lambda, sweeps, first_v);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_ltx_all(const OnlineX* xs, int d, const int* sched, int mb_total, const int* t_dev, float* B, cudaStream_t st) {
    k_online_ltx_all<<<(unsigned)(((size_t)mb_total * 32 + 255) / 256), 256, 0, st>>>(xs, d, sched, mb_total, t_dev, B);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_stats_all(const OnlineDs* ds, const OnlineX* xs, int d, int m, int k, const int* sched, int mb_total, int* t_dev,
                             const float* H, cudaStream_t st) {
    k_online_stats_a_t<<<d, 1024, 0, st>>>(ds, k, t_dev);
    dim3 g((unsigned)(((size_t)m * OKP + 255) / 256), (unsigned)d);
    k_online_scale_all<<<g, 256, 0, st>>>(xs, m, t_dev);
    k_online_stats_b_all<<<(unsigned)(((size_t)mb_total * 32 + 255) / 256), 256, 0, st>>>(xs, d, sched, mb_total, t_dev, H);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_next(int* t_dev, cudaStream_t st) {
    k_online_next<<<1, 1, 0, st>>>(t_dev);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_cross(const float* H, const float* B, size_t n, double* out, cudaStream_t st) {
    if (!n) return;
    k_online_cross<<<296, 256, 0, st>>>(H, B, n, out);
    LGR_CUDA(cudaGetLastError());
}
void launch_online_sumsq(const float* x, size_t n, double* out, cudaStream_t st) {
    if (!n) return;
    k_online_sumsq<<<296, 256, 0, st>>>(x, n, out);
    LGR_CUDA(cudaGetLastError());
}

}  // namespace lgr
