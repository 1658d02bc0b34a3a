// liger_b200 -- centroidAlign on the device (second alignment method).
// Mirrors .centroidAlign.Hraw (reference R/integration.R:2015-2084): Z = safe_scale(H) (src/data_processing.cpp:61-85),
// soft cluster memberships R = normalize_byCol(safe_scale(H)) (src/data_processing.cpp:10-20), and the mixture-of-experts
// ridge correction moe_correct_ridge_cpp (src/centoidAlign.cpp:16-85).  With Phi the one-hot dataset design, the correction
// of cluster c only needs, per dataset b, the weighted sums  s[c][b] = sum_n R[c,n]  and  S[c][b][:] = sum_n R[c,n] Z[:,n]:
//     cov_c = [[sum_b s, s^T], [s, diag(s + lambda)]],   W_c = cov_c^-1 [sum_b S; S_1; ..; S_B]  (row 0 zeroed),
//     Z_corr[:, n] = Z[:, n] - sum_c R[c, n] W_c[b(n) + 1, :]
// so the device work is two passes over H (k x k accumulations per dataset, then a k x k product per cell), FP64 throughout;
// the (B + 1) x (B + 1) inverses are host work.
#include <algorithm>
#include <cmath>
#include <vector>

#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int CA_MAX_K = 64;
constexpr int CA_TILE = 32;   // cells staged per step of the accumulation kernel

struct CaCols {     // per used column: how Z and R are derived from the raw value
    double z_sub, z_mul;   // z = (x - z_sub) * z_mul           (z_mul = 1 / sd, or 0 for a constant column)
    double r_sub, r_mul;   // r = (x - r_sub) * r_mul - shift
};

// sums[j] += sum_i (H[i, col_j] - sub[j])^p  over one dataset (p = 1 or 2)
__global__ void k_ca_colsum(const double* __restrict__ H, int n, const int* __restrict__ cols, const double* __restrict__ sub,
                            int square, double* __restrict__ sums) {
    const int j = blockIdx.y;
    const double* col = H + (size_t)cols[j] * n;
    const double s0 = sub ? sub[j] : 0.0;
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = col[i] - s0;
        s += square ? v * v : v;
    }
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(sums + j, s);
    }
}

// minimum of the pre-normalisation R over one dataset (ordered-int trick on the sign-flipped bits is avoided: one slot per block)
__global__ void k_ca_rmin(const double* __restrict__ H, int n, const int* __restrict__ cols, int kd, const CaCols* __restrict__ cc,
                          double* __restrict__ block_min) {
    double m = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int j = 0; j < kd; ++j) m = fmin(m, (H[(size_t)cols[j] * n + i] - cc[j].r_sub) * cc[j].r_mul);
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, sh[w]);
        block_min[blockIdx.x] = m;
    }
}

// S[c][j] += sum_n R[c, n] Z[j, n],  s[c] += sum_n R[c, n]   over one dataset.  Thread (c, j) of a kd x kd block.
__global__ void k_ca_accum(const double* __restrict__ H, int n, const int* __restrict__ cols, int kd, const CaCols* __restrict__ cc,
                           double shift, double* __restrict__ S, double* __restrict__ s) {
    extern __shared__ double sm[];
    double* zt = sm;                       // [CA_TILE][kd]
    double* rt = sm + CA_TILE * kd;        // [CA_TILE][kd]
    const int tid = threadIdx.x;
    const int c = tid / kd, j = tid % kd;
    const bool worker = tid < kd * kd;
    double acc = 0.0, accs = 0.0;
    for (int base = blockIdx.x * CA_TILE; base < n; base += gridDim.x * CA_TILE) {
        __syncthreads();
        for (int e = tid; e < CA_TILE * kd; e += blockDim.x) {
            const int t = e % CA_TILE, jj = e / CA_TILE;   // consecutive threads -> consecutive cells of one column
            const int i = base + t;
            double z = 0.0, r = 0.0;
            if (i < n) {
                const double x = H[(size_t)cols[jj] * n + i];
                z = (x - cc[jj].z_sub) * cc[jj].z_mul;
                r = (x - cc[jj].r_sub) * cc[jj].r_mul - shift;
            }
            zt[t * kd + jj] = z;
            rt[t * kd + jj] = r;
        }
        __syncthreads();
        if (tid < CA_TILE) {   // normalize_byCol: memberships of a cell sum to 1 (all-zero cells stay 0)
            double sum = 0.0;
            for (int jj = 0; jj < kd; ++jj) sum += rt[tid * kd + jj];
            for (int jj = 0; jj < kd; ++jj) rt[tid * kd + jj] = sum == 0.0 ? 0.0 : rt[tid * kd + jj] / sum;
        }
        __syncthreads();
        if (worker) {
#pragma unroll 8
            for (int t = 0; t < CA_TILE; ++t) {
                const double r = rt[t * kd + c];
                acc += r * zt[t * kd + j];
                if (j == 0) accs += r;
            }
        }
    }
    if (worker) {
        atomicAdd(S + c * kd + j, acc);
        if (j == 0) atomicAdd(s + c, accs);
    }
}

// out[i, j] = Z[j, i] - sum_c R[c, i] Wb[c][j]   (Wb = this dataset's rows of all W_c, kd x kd), plus the which.max diagnostics
__global__ void k_ca_correct(const double* __restrict__ H, int n, const int* __restrict__ cols, int kd, const CaCols* __restrict__ cc,
                             double shift, const double* __restrict__ Wb, double* __restrict__ out, int* __restrict__ raw_max,
                             int* __restrict__ z_max, int* __restrict__ r_max) {
    extern __shared__ double sm[];
    double* W = sm;   // kd x kd
    for (int e = threadIdx.x; e < kd * kd; e += blockDim.x) W[e] = Wb[e];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[CA_MAX_K], z[CA_MAX_K];
    double sum = 0.0;
    int rawm = 0;
    double rawv = 0.0;
    for (int j = 0; j < kd; ++j) {
        const double x = H[(size_t)cols[j] * n + i];
        z[j] = (x - cc[j].z_sub) * cc[j].z_mul;
        r[j] = (x - cc[j].r_sub) * cc[j].r_mul - shift;
        sum += r[j];
        if (j == 0 || x > rawv) {
            rawv = x;
            rawm = j;
        }
    }
    int rm = 0;
    for (int j = 0; j < kd; ++j) {
        r[j] = sum == 0.0 ? 0.0 : r[j] / sum;
        if (r[j] > r[rm]) rm = j;
    }
    int zm = 0;
    double zv = 0.0;
    for (int j = 0; j < kd; ++j) {
        double v = z[j];
        for (int c = 0; c < kd; ++c) v -= r[c] * W[c * kd + j];
        out[(size_t)j * n + i] = v;
        if (j == 0 || v > zv) {
            zv = v;
            zm = j;
        }
    }
    if (raw_max) raw_max[i] = rawm + 1;
    if (z_max) z_max[i] = zm + 1;
    if (r_max) r_max[i] = rm + 1;
}

// inverse of a small dense matrix by Gauss-Jordan with partial pivoting (host, FP64); false when singular
bool invert_small(std::vector<double>& a, int n) {
    std::vector<double> inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
        if (a[(size_t)piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                std::swap(a[(size_t)piv * n + j], a[(size_t)c * n + j]);
                std::swap(inv[(size_t)piv * n + j], inv[(size_t)c * n + j]);
            }
        const double d = 1.0 / a[(size_t)c * n + c];
        for (int j = 0; j < n; ++j) {
            a[(size_t)c * n + j] *= d;
            inv[(size_t)c * n + j] *= d;
        }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                a[(size_t)r * n + j] -= f * a[(size_t)c * n + j];
                inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
            }
        }
    }
    a.swap(inv);
    return true;
}

}  // namespace

// Hdev[i]: n_i x k column-major FP64 (read only); out_dev[i]: n_i x kd column-major, kd = number of used dims
void centroid_align_core(int d, int k, const int* n, const double* const* Hdev, const lgr_ca_params& p, double* const* out_dev,
                         int* const* raw_max, int* const* z_max, int* const* r_max, cudaStream_t st) {
    LGR_REQUIRE(d >= 1 && k >= 1 && k <= CA_MAX_K, "centroidAlign: d >= 1 and 1 <= k <= 64 required");
    LGR_REQUIRE(p.lambda, "centroidAlign: lambda is required (one value per dataset)");
    const int kd = p.n_dims_use > 0 ? p.n_dims_use : k;
    LGR_REQUIRE(kd <= 32, "centroidAlign: at most 32 dims are supported");
    std::vector<int> cols(kd);
    for (int j = 0; j < kd; ++j) {
        cols[j] = p.n_dims_use > 0 ? p.dims_use[j] - 1 : j;
        LGR_REQUIRE(cols[j] >= 0 && cols[j] < k, "centroidAlign: useDims out of range");
    }
    long long N = 0;
    for (int i = 0; i < d; ++i) N += n[i];
    LGR_REQUIRE(N >= 2, "centroidAlign: at least two cells required");
    DBuf<int> d_cols(kd);
    LGR_CUDA(cudaMemcpyAsync(d_cols.p, cols.data(), sizeof(int) * kd, cudaMemcpyHostToDevice, st));

    // ---- safe_scale statistics over ALL cells: means, then the sums of squares of the (optionally centred) columns ----
    DBuf<double> d_sum(kd), d_sub(kd), d_ss(kd);
    auto colsum = [&](const double* sub, int square, double* dst, std::vector<double>& host) {
        LGR_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * kd, st));
        for (int i = 0; i < d; ++i) {
            if (n[i] == 0) continue;
            dim3 g((unsigned)std::min(64, (n[i] + 255) / 256), (unsigned)kd);
            k_ca_colsum<<<g, 256, 0, st>>>(Hdev[i], n[i], d_cols.p, sub, square, dst);
        }
        LGR_CUDA(cudaGetLastError());
        host.resize(kd);
        LGR_CUDA(cudaMemcpyAsync(host.data(), dst, sizeof(double) * kd, cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaStreamSynchronize(st));
    };
    std::vector<double> sums, mean(kd), ss_centered, ss_raw;
    colsum(nullptr, 0, d_sum.p, sums);
    for (int j = 0; j < kd; ++j) mean[j] = sums[j] / (double)N;
    colsum(nullptr, 1, d_ss.p, ss_raw);
    LGR_CUDA(cudaMemcpyAsync(d_sub.p, mean.data(), sizeof(double) * kd, cudaMemcpyHostToDevice, st));
    colsum(d_sub.p, 1, d_ss.p, ss_centered);
    std::vector<CaCols> cc(kd);
    auto derive = [&](bool center, bool scale, int j, double& sub, double& mul) {
        sub = center ? mean[j] : 0.0;
        mul = 1.0;
        if (scale) {
            const double sd = std::sqrt((center ? ss_centered[j] : ss_raw[j]) / (double)(N - 1));
            mul = sd == 0.0 ? 0.0 : 1.0 / sd;   // a constant column is filled with 0
        }
    };
    for (int j = 0; j < kd; ++j) {
        derive(p.center_emb != 0, p.scale_emb != 0, j, cc[j].z_sub, cc[j].z_mul);
        derive(p.center_cluster != 0, p.scale_cluster != 0, j, cc[j].r_sub, cc[j].r_mul);
    }
    DBuf<CaCols> d_cc(kd);
    LGR_CUDA(cudaMemcpyAsync(d_cc.p, cc.data(), sizeof(CaCols) * kd, cudaMemcpyHostToDevice, st));

    // ---- shift / sign check of the memberships before normalisation ----
    double rmin = INFINITY;
    {
        DBuf<double> bm(64);
        std::vector<double> hb(64);
        for (int i = 0; i < d; ++i) {
            if (n[i] == 0) continue;
            const int nb = std::min(64, (n[i] + 255) / 256);
            k_ca_rmin<<<nb, 256, 0, st>>>(Hdev[i], n[i], d_cols.p, kd, d_cc.p, bm.p);
            LGR_CUDA(cudaGetLastError());
            LGR_CUDA(cudaMemcpyAsync(hb.data(), bm.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
            LGR_CUDA(cudaStreamSynchronize(st));
            for (int b = 0; b < nb; ++b) rmin = std::min(rmin, hb[b]);
        }
    }
    const double shift = p.shift ? rmin : 0.0;
    if (rmin - shift < 0.0)
        throw Error(LGR_ERR_INVALID,
                    "Negative values found prior to normalizing the cluster assignment probablity distribution. Can only do either "
                    "centerCluster = FALSE or centerCluster = TRUE, shift = TRUE.");

    // ---- pass 1: weighted sums per dataset ----
    const int threads = std::max(64, ((kd * kd + 31) / 32) * 32);
    LGR_REQUIRE(threads <= 1024, "centroidAlign: kd * kd must be <= 1024 (k <= 32)");
    const size_t smem1 = (size_t)2 * CA_TILE * kd * sizeof(double);
    DBuf<double> dS((size_t)d * kd * kd), ds((size_t)d * kd);
    dS.zero(st);
    ds.zero(st);
    for (int i = 0; i < d; ++i) {
        if (n[i] == 0) continue;
        const int nb = std::min(592, (n[i] + CA_TILE - 1) / CA_TILE);
        k_ca_accum<<<nb, threads, smem1, st>>>(Hdev[i], n[i], d_cols.p, kd, d_cc.p, shift, dS.p + (size_t)i * kd * kd, ds.p + (size_t)i * kd);
    }
    LGR_CUDA(cudaGetLastError());
    std::vector<double> S((size_t)d * kd * kd), s((size_t)d * kd);
    LGR_CUDA(cudaMemcpyAsync(S.data(), dS.p, sizeof(double) * S.size(), cudaMemcpyDeviceToHost, st));
    LGR_CUDA(cudaMemcpyAsync(s.data(), ds.p, sizeof(double) * s.size(), cudaMemcpyDeviceToHost, st));
    LGR_CUDA(cudaStreamSynchronize(st));

    // ---- host: (B + 1) x (B + 1) ridge systems, one per cluster ----
    const int B1 = d + 1;
    std::vector<double> Wb((size_t)d * kd * kd, 0.0);   // [dataset][cluster][dim]
    std::vector<double> cov((size_t)B1 * B1), rhs((size_t)B1 * kd);
    for (int c = 0; c < kd; ++c) {
        std::fill(cov.begin(), cov.end(), 0.0);
        double tot = 0.0;
        for (int b = 0; b < d; ++b) {
            const double sb = s[(size_t)b * kd + c];
            tot += sb;
            cov[(size_t)0 * B1 + b + 1] = sb;
            cov[(size_t)(b + 1) * B1 + 0] = sb;
            cov[(size_t)(b + 1) * B1 + b + 1] = sb + p.lambda[b];
        }
        cov[0] = tot;
        if (!invert_small(cov, B1)) throw Error(LGR_ERR_INVALID, "centroidAlign: singular ridge system (a cluster with no weight and lambda = 0)");
        for (int j = 0; j < kd; ++j) {
            double t = 0.0;
            for (int b = 0; b < d; ++b) {
                rhs[(size_t)(b + 1) * kd + j] = S[((size_t)b * kd + c) * kd + j];
                t += rhs[(size_t)(b + 1) * kd + j];
            }
            rhs[j] = t;
        }
        for (int b = 0; b < d; ++b)       // row b + 1 of W_c = inv_cov[b + 1, :] rhs   (row 0, the intercept, is zeroed)
            for (int j = 0; j < kd; ++j) {
                double w = 0.0;
                for (int q = 0; q < B1; ++q) w += cov[(size_t)(b + 1) * B1 + q] * rhs[(size_t)q * kd + j];
                Wb[((size_t)b * kd + c) * kd + j] = w;
            }
    }
    DBuf<double> dW(Wb.size());
    LGR_CUDA(cudaMemcpyAsync(dW.p, Wb.data(), sizeof(double) * Wb.size(), cudaMemcpyHostToDevice, st));

    // ---- pass 2: the correction ----
    for (int i = 0; i < d; ++i) {
        if (n[i] == 0) continue;
        k_ca_correct<<<(n[i] + 127) / 128, 128, (size_t)kd * kd * sizeof(double), st>>>(
            Hdev[i], n[i], d_cols.p, kd, d_cc.p, shift, dW.p + (size_t)i * kd * kd, out_dev[i], raw_max ? raw_max[i] : nullptr,
            z_max ? z_max[i] : nullptr, r_max ? r_max[i] : nullptr);
    }
    LGR_CUDA(cudaGetLastError());
    LGR_CUDA(cudaStreamSynchronize(st));
}

}  // namespace lgr
