// liger_b200 -- quantileNorm on the device: the step that consumes the H_i of the hot path where they already live
// (SURVEY section 8 (f) rank 3).  Mirrors .quantileNorm.HList (reference R/integration.R:1621-1711) with its two compiled
// helpers max_factor_rcpp / cluster_vote_rcpp (src/quantile_norm.cpp:8-67) and the library calls it makes:
//
//   1. cluster = arg-max factor of the column-scaled H                          k_col_sumsq, k_max_factor
//   2. refineKNN: RANN::nn2(H_i, k = nNeighbors, eps) and a majority vote       ann_build (host), k_ann_search, k_vote_sweep
//   3. per (dataset, cluster, factor): type-7 quantiles of both sides (of a     k_quantile_warp
//      base::sample() of maxSample values when the cluster is larger),
//      approxfun(q2, q1, rule = 2) applied to the dataset's values
//
// Faithfulness notes (each one is what makes data/pbmcPlot.rda's H.norm come out to round-off, tests/test_gpu_parity.py):
//   * nn2(eps > 0) is ANN's (1 + eps)-approximate search: WHICH neighbours come back depends on the kd-tree (bucket size 1,
//     sliding-midpoint splits) and on the order and pruning rule of its depth-first search.  The tree is built on the host
//     exactly as ANN builds it and searched on the device with one thread per query following ANN's traversal (explicit
//     stack, the far subtree is tested when it is popped, i.e. after the near subtree has updated the k-th best distance),
//     in FP64 with no fused multiply-adds.
//   * cluster_vote_rcpp overwrites the label vector it reads, so cell i sees the NEW labels of neighbours j < i.  That
//     triangular dependency is solved by Jacobi sweeps to its (unique) fixed point instead of a serial loop.
//   * labels are R factor codes of the labels AS STRINGS: vote ties go to the lexicographically smallest label.
//   * the datasets are warped in order and the reference dataset is warped onto itself when its turn comes (later datasets
//     see the warped reference).
//   * base::sample() consumes R's RNG stream for every (dataset, cluster, factor) even when it only permutes; the host
//     replays all calls (R >= 3.6 rejection sampling) when some cluster exceeds maxSample, and ships the drawn positions.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "lgr_common.cuh"

namespace lgr {
namespace {

constexpr int QN_MAX_SAMPLE = 4096;     // shared-memory sort capacity per side
constexpr int QN_MAX_QUANT = 512;       // quantile knots
constexpr int ANN_STACK = 192;          // pending far subtrees per query (tree depth is ~2 log2 n)
constexpr int ANN_MAX_K = 64;
constexpr int ANN_MAX_DIM = 64;

struct AnnNode {     // split: cd >= 0, children lo / hi;  leaf: cd = -1, lo = point index (or -1: empty)
    int cd, lo, hi, pad;
    double cv, lv, hv;
};

// ---------------------------------------------------------------------------------------------------------------------
// ANN 1.1.2 kd-tree (the library behind RANN::nn2; not part of the reference tree -- restated from its published
// algorithm: bucket size 1, sliding-midpoint rule with ERR = 0.001, two-pass plane split)
// ---------------------------------------------------------------------------------------------------------------------
struct AnnBuilder {
    const double* pts;   // row-major n x dim
    int dim;
    std::vector<int> pidx;
    std::vector<AnnNode> nodes;
    std::vector<double> blo, bhi;
    int depth_max = 0;

    double P(int i, int d) const { return pts[(size_t)pidx[i] * dim + d]; }

    void plane_split(int a, int n, int d, double cv, int& br1, int& br2) {
        int l = 0, r = n - 1;
        for (;;) {
            while (l < n && P(a + l, d) < cv) l++;
            while (r >= 0 && P(a + r, d) >= cv) r--;
            if (l > r) break;
            std::swap(pidx[a + l], pidx[a + r]);
            l++;
            r--;
        }
        br1 = l;
        r = n - 1;
        for (;;) {
            while (l < n && P(a + l, d) <= cv) l++;
            while (r >= br1 && P(a + r, d) > cv) r--;
            if (l > r) break;
            std::swap(pidx[a + l], pidx[a + r]);
            l++;
            r--;
        }
        br2 = l;
    }

    int build(int a, int n, int depth) {
        if (depth > depth_max) depth_max = depth;
        if (n <= 1) {
            nodes.push_back(AnnNode{-1, n == 1 ? pidx[a] : -1, -1, 0, 0.0, 0.0, 0.0});
            return (int)nodes.size() - 1;
        }
        double max_length = 0.0;
        for (int d = 0; d < dim; ++d) max_length = std::max(max_length, bhi[d] - blo[d]);
        double max_spread = -1.0;
        int cd = 0;
        for (int d = 0; d < dim; ++d) {
            if (bhi[d] - blo[d] >= (1.0 - 0.001) * max_length) {
                double mn = P(a, d), mx = mn;
                for (int i = 1; i < n; ++i) {
                    const double v = P(a + i, d);
                    mn = std::min(mn, v);
                    mx = std::max(mx, v);
                }
                if (mx - mn > max_spread) {
                    max_spread = mx - mn;
                    cd = d;
                }
            }
        }
        const double ideal = (blo[cd] + bhi[cd]) / 2.0;
        double mn = P(a, cd), mx = mn;
        for (int i = 1; i < n; ++i) {
            const double v = P(a + i, cd);
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        const double cv = ideal < mn ? mn : (ideal > mx ? mx : ideal);
        int br1, br2;
        plane_split(a, n, cd, cv, br1, br2);
        int n_lo;
        if (ideal < mn)
            n_lo = 1;
        else if (ideal > mx)
            n_lo = n - 1;
        else if (br1 > n / 2)
            n_lo = br1;
        else if (br2 < n / 2)
            n_lo = br2;
        else
            n_lo = n / 2;
        const double lv = blo[cd], hv = bhi[cd];
        bhi[cd] = cv;
        const int lo_child = build(a, n_lo, depth + 1);
        bhi[cd] = hv;
        blo[cd] = cv;
        const int hi_child = build(a + n_lo, n - n_lo, depth + 1);
        blo[cd] = lv;
        nodes.push_back(AnnNode{cd, lo_child, hi_child, 0, cv, lv, hv});
        return (int)nodes.size() - 1;
    }
};

// one thread per query; qs: this block's query points as [dim][thread].  The queries are taken in the tree's leaf order
// (qorder = the builder's final point permutation): the threads of a warp then hold neighbouring points, walk nearly the
// same path and touch the same nodes and points (coherent branches and loads); every query's result is unaffected.
__global__ void __launch_bounds__(128) k_ann_search(const AnnNode* __restrict__ nodes, int root, const double* __restrict__ pts, int n,
                                                    int dim, const double* __restrict__ blo, const double* __restrict__ bhi, int kk,
                                                    double max_err, const int* __restrict__ qorder, int* __restrict__ nn_idx,
                                                    int* __restrict__ overflow) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = slot < n;
    if (!live) return;
    const int q = qorder[slot];
    double qv[ANN_MAX_DIM];
    for (int d = 0; d < dim; ++d) qv[d] = pts[(size_t)q * dim + d];
    double keys[ANN_MAX_K + 1];
    int ids[ANN_MAX_K + 1];
    int cnt = 0;
    int snode[ANN_STACK];
    double sdist[ANN_STACK];
    int sp = 0;
    double bdist = 0.0;
    for (int d = 0; d < dim; ++d) {
        const double v = qv[d];
        if (v < blo[d]) {
            const double t = __dsub_rn(blo[d], v);
            bdist = __dadd_rn(bdist, __dmul_rn(t, t));
        } else if (v > bhi[d]) {
            const double t = __dsub_rn(v, bhi[d]);
            bdist = __dadd_rn(bdist, __dmul_rn(t, t));
        }
    }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    snode[0] = root;
    sdist[0] = bdist;
    sp = 1;
    bool first = true;
    while (sp > 0) {
        --sp;
        int node = snode[sp];
        const double bd = sdist[sp];
        const double mk = cnt == kk ? keys[kk - 1] : INF;
        if (!first && !(__dmul_rn(bd, max_err) < mk)) continue;   // the far-subtree test, made after the near subtree ran
        first = false;
        for (;;) {
            const AnnNode nd = nodes[node];
            if (nd.cd < 0) {
                const int pid = nd.lo;
                if (pid >= 0) {
                    const double md = cnt == kk ? keys[kk - 1] : INF;
                    const double* pp = pts + (size_t)pid * dim;
                    double dist = 0.0;
                    int d = 0;
                    for (; d < dim; ++d) {
                        const double t = __dsub_rn(qv[d], pp[d]);
                        dist = __dadd_rn(dist, __dmul_rn(t, t));
                        if (dist > md) break;
                    }
                    if (d >= dim) {
                        int i = cnt;
                        while (i > 0 && keys[i - 1] > dist) {
                            keys[i] = keys[i - 1];
                            ids[i] = ids[i - 1];
                            --i;
                        }
                        keys[i] = dist;
                        ids[i] = pid;
                        if (cnt < kk) ++cnt;
                    }
                }
                break;
            }
            const double qc = qv[nd.cd];
            const double cut_diff = __dsub_rn(qc, nd.cv);
            int near_child, far_child;
            double box_diff;
            if (cut_diff < 0) {
                near_child = nd.lo;
                far_child = nd.hi;
                box_diff = __dsub_rn(nd.lv, qc);
            } else {
                near_child = nd.hi;
                far_child = nd.lo;
                box_diff = __dsub_rn(qc, nd.hv);
            }
            if (box_diff < 0) box_diff = 0.0;
            const double bd2 = __dadd_rn(bd, __dsub_rn(__dmul_rn(cut_diff, cut_diff), __dmul_rn(box_diff, box_diff)));
            if (sp < ANN_STACK) {
                snode[sp] = far_child;
                sdist[sp] = bd2;
                ++sp;
            } else {
                atomicExch(overflow, 1);
            }
            node = near_child;
        }
        // NOTE: `bd` of the nodes below the popped one accumulates along the path in ANN (box_dist is passed down and
        // replaced by bd2 only for the far child), which is what the loop above does: the near child inherits `bd`.
    }
    for (int i = 0; i < kk; ++i) nn_idx[(size_t)q * kk + i] = i < cnt ? ids[i] : -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// max_factor_rcpp (src/quantile_norm.cpp:38-67)
// ---------------------------------------------------------------------------------------------------------------------
// stats[j] = (sum, sum of squares) of column j (H column-major n x k)
__global__ void k_col_stats(const double* __restrict__ H, int n, int k, double* __restrict__ stats) {
    const int j = blockIdx.y;
    double s = 0.0, s2 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = H[(size_t)j * n + i];
        s += v;
        s2 += v * v;
    }
    __shared__ double sh[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s;
        sh[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.0;
        s2 = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (threadIdx.x == 0) {
            atomicAdd(stats + 2 * j, s);
            atomicAdd(stats + 2 * j + 1, s2);
        }
    }
}
// label[i] = rank of (1 + arg-max_j (H[i, j] - mean_j) / sd_j over the used dims), "-1" when no scaled entry is > 0
__global__ void k_max_factor(const double* __restrict__ H, int n, int k, const double* __restrict__ stats, int center,
                             const int* __restrict__ dims_use, int ndims, const int* __restrict__ rank_of_label,
                             int* __restrict__ label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int best = -1;
    double best_val = 0.0;
    for (int t = 0; t < ndims; ++t) {
        const int j = dims_use ? dims_use[t] - 1 : t;
        const double mean = stats[2 * j] / n;
        double v = H[(size_t)j * n + i];
        double ss = stats[2 * j + 1];
        if (center) {
            v -= mean;
            ss -= n * mean * mean;   // sum (x - mean)^2
        }
        const double sd = sqrt(ss / (n - 1));
        v = v / sd;
        if (v > best_val) {
            best = j + 1;
            best_val = v;
        }
    }
    label[i] = rank_of_label[best < 0 ? 0 : best];
}

// ---------------------------------------------------------------------------------------------------------------------
// cluster_vote_rcpp (src/quantile_norm.cpp:8-35), in-place semantics by fixed-point sweeps:
//     next[i] = vote over the neighbours j of  (j < i ? cur[j] : old[j])
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_vote_sweep(const int* __restrict__ nn, int n, int kk, const int* __restrict__ old_rank, const int* __restrict__ cur,
                             int* __restrict__ next, int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lab[ANN_MAX_K];
    int m = 0;
    for (int a = 0; a < kk; ++a) {
        const int j = nn[(size_t)i * kk + a];
        if (j < 0) continue;
        lab[m++] = j < i ? cur[j] : old_rank[j];
    }
    int best = -1, best_cnt = 0;
    for (int a = 0; a < m; ++a) {
        int c = 0;
        for (int b = 0; b < m; ++b) c += lab[b] == lab[a];
        if (c > best_cnt || (c == best_cnt && lab[a] < best)) {
            best = lab[a];
            best_cnt = c;
        }
    }
    next[i] = best;
    if (best != cur[i]) atomicExch(changed, 1);
}

__global__ void k_rank_to_label(const int* __restrict__ rank, int n, const int* __restrict__ label_of_rank, int* __restrict__ out_label,
                                int* __restrict__ sort_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = label_of_rank[rank[i]];
    out_label[i] = l;
    sort_key[i] = l < 0 ? 0 : l;   // 0 = "no cluster"
}
__global__ void k_iota(int* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_hist_labels(const int* __restrict__ key, int n, int* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + key[i], 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// The quantile warp of one dataset: CTA (c, f) = cluster c + 1, factor f.
//   cells2 / cells1: row indices of the cluster's cells in the dataset / in the reference, in row order
//   samp2 / samp1:   positions (into those lists) that base::sample() drew, or null = all cells
// stats::quantile type 7, approxfun(rule = 2, ties = mean) restated from their documented algorithms; FP64, no FMA in the
// interpolation formulas so that the knots and the warped values round like R's.
// ---------------------------------------------------------------------------------------------------------------------
struct WarpJob {
    int n2, n1;          // cluster sizes (dataset, reference)
    int s2, s1;          // sample sizes
    int off2, off1;      // offsets of the cluster's cell lists
    long long samp2, samp1;   // offsets into the sampled-position arena for (this cluster, factor 0), -1 = all cells; stride per factor = s2 + s1?  (see host)
};

__device__ __forceinline__ void bitonic_sort(double* a, int np2) {
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ double quantile7_at(const double* x, int n, double prob) {
    const double index = __dadd_rn(1.0, __dmul_rn((double)(n > 1 ? n - 1 : 0), prob));
    const double lo = floor(index), hi = ceil(index);
    double qs = x[(int)lo - 1];
    const double xh = x[(int)hi - 1];
    if (index > lo && xh != qs) {
        const double h = __dsub_rn(index, lo);
        qs = __dadd_rn(__dmul_rn(__dsub_rn(1.0, h), qs), __dmul_rn(h, xh));
    }
    return qs;
}

__global__ void __launch_bounds__(256) k_quantile_warp(double* Hd, int nd, const double* Href, int nref,   /* Hd == Href for the reference dataset */
                                                       const WarpJob* __restrict__ jobs, const int* __restrict__ cells_d,
                                                       const int* __restrict__ cells_ref, const int* __restrict__ samp, int k,
                                                       int nq, const double* __restrict__ probs, int cap) {
    extern __shared__ double sm[];
    double* a2 = sm;            // cap
    double* a1 = sm + cap;      // cap
    double* q2 = a1 + cap;      // nq
    double* q1 = q2 + nq;       // nq
    double* ux = q1 + nq;       // nq   regularised knots
    double* uy = ux + nq;       // nq
    __shared__ int nu, degenerate;
    const WarpJob J = jobs[blockIdx.x];
    if (J.n2 <= 0) return;      // cluster skipped (fewer than minCells on one side)
    const int f = blockIdx.y;
    double* col_d = Hd + (size_t)f * nd;
    const double* col_r = Href + (size_t)f * nref;
    const int* c2 = cells_d + J.off2;
    const int* c1 = cells_ref + J.off1;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (J.n2 == 1) {            // R/integration.R:1671-1675: a single cell takes the reference cluster's mean
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int t = 0; t < J.n1; ++t) s += col_r[c1[t]];
            col_d[c2[0]] = s / J.n1;
        }
        return;
    }
    int np2 = 1;
    while (np2 < J.s2) np2 <<= 1;
    int np1 = 1;
    while (np1 < J.s1) np1 <<= 1;
    const int* sp2 = J.samp2 >= 0 ? samp + J.samp2 + (long long)f * (J.s2 + J.s1) : nullptr;
    const int* sp1 = J.samp1 >= 0 ? samp + J.samp1 + (long long)f * (J.s2 + J.s1) : nullptr;
    for (int t = threadIdx.x; t < np2; t += blockDim.x) a2[t] = t < J.s2 ? col_d[c2[sp2 ? sp2[t] : t]] : INF;
    for (int t = threadIdx.x; t < np1; t += blockDim.x) a1[t] = t < J.s1 ? col_r[c1[sp1 ? sp1[t] : t]] : INF;
    __syncthreads();
    bitonic_sort(a2, np2);
    bitonic_sort(a1, np1);
    for (int t = threadIdx.x; t < nq; t += blockDim.x) {
        q2[t] = quantile7_at(a2, J.s2, probs[t]);
        q1[t] = quantile7_at(a1, J.s1, probs[t]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // sum(q) == 0 or fewer than two distinct knots on either side: the cluster's values become 0
        double s1 = 0.0, s2 = 0.0;
        int d1 = 1, d2 = 1;
        for (int t = 0; t < nq; ++t) {
            s1 += q1[t];
            s2 += q2[t];
            if (t > 0 && q1[t] != q1[t - 1]) ++d1;   // quantiles are non-decreasing: distinct = changes + 1
            if (t > 0 && q2[t] != q2[t - 1]) ++d2;
        }
        degenerate = (s1 == 0.0 || s2 == 0.0 || d1 < 2 || d2 < 2) ? 1 : 0;
        // approxfun's regularize.values(ties = mean): q2 is already sorted; tied x collapse to one knot with the mean y
        int m = 0;
        for (int t = 0; t < nq;) {
            int e = t + 1;
            double s = q1[t];
            while (e < nq && q2[e] == q2[t]) {
                s += q1[e];
                ++e;
            }
            ux[m] = q2[t];
            uy[m] = e - t == 1 ? q1[t] : s / (double)(e - t);
            ++m;
            t = e;
        }
        nu = m;
    }
    __syncthreads();
    const int m = nu;
    for (int t = threadIdx.x; t < J.n2; t += blockDim.x) {
        const int row = c2[t];
        double out = 0.0;
        if (!degenerate) {
            const double v = col_d[row];
            if (v < ux[0]) {
                out = uy[0];
            } else if (v > ux[m - 1]) {
                out = uy[m - 1];
            } else {
                int i = 0, j = m - 1;
                while (i < j - 1) {   // R_approxfun's bisection
                    const int ij = (i + j) / 2;
                    if (v < ux[ij])
                        j = ij;
                    else
                        i = ij;
                }
                if (v == ux[j])
                    out = uy[j];
                else if (v == ux[i])
                    out = uy[i];
                else
                    out = __dadd_rn(uy[i], __dmul_rn(__dsub_rn(uy[j], uy[i]), __ddiv_rn(__dsub_rn(v, ux[i]), __dsub_rn(ux[j], ux[i]))));
            }
        }
        // every thread reads its value before anyone writes: the reads above and this write touch the same entry only
        col_d[row] = out;
    }
}

}  // namespace

// Hdev[i]: n_i x k column-major FP64 on the device, replaced by H.norm; labels_dev[i]: n_i cluster labels (1-based factor id,
// -1 = none).  Everything runs on `st`; returns after a final synchronisation.
void quantile_norm_core(int d, int k, const int* n, double* const* Hdev, const lgr_qn_params& p, int* const* labels_dev,
                        lgr_qn_info* info, cudaStream_t st) {
    LGR_REQUIRE(d >= 1 && k >= 1 && k <= ANN_MAX_DIM, "quantileNorm: d >= 1 and 1 <= k <= 64 required");
    LGR_REQUIRE(p.reference >= 0 && p.reference < d, "Should specify one existing dataset as reference.");
    LGR_REQUIRE(p.quantiles >= 1 && p.quantiles + 1 <= QN_MAX_QUANT, "quantileNorm: quantiles must be in 1..511");
    LGR_REQUIRE(p.max_sample >= 2 && p.max_sample <= QN_MAX_SAMPLE, "quantileNorm: maxSample must be in 2..4096");
    LGR_REQUIRE(p.n_neighbors >= 1 && p.n_neighbors <= ANN_MAX_K, "quantileNorm: nNeighbors must be in 1..64");
    LGR_REQUIRE(p.eps >= 0.0, "quantileNorm: eps must be >= 0");
    for (int t = 0; t < p.n_dims_use; ++t)
        LGR_REQUIRE(p.dims_use && p.dims_use[t] >= 1 && p.dims_use[t] <= k, "quantileNorm: useDims out of range");
    for (int i = 0; i < d; ++i) LGR_REQUIRE(n[i] >= 2, "quantileNorm: every dataset needs at least two cells");
    if (info) *info = lgr_qn_info{};
    cudaEvent_t e0, e1, e2, e3;
    LGR_CUDA(cudaEventCreate(&e0));
    LGR_CUDA(cudaEventCreate(&e1));
    LGR_CUDA(cudaEventCreate(&e2));
    LGR_CUDA(cudaEventCreate(&e3));
    LGR_CUDA(cudaEventRecord(e0, st));

    // ---- factor levels: labels as strings ("-1", "1", .., "k"), sorted; ranks decide vote ties ----
    std::vector<std::string> lv;
    lv.push_back("-1");
    for (int j = 1; j <= k; ++j) lv.push_back(std::to_string(j));
    std::vector<int> order(lv.size());
    for (size_t t = 0; t < lv.size(); ++t) order[t] = (int)t;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return lv[a] < lv[b]; });
    std::vector<int> rank_of_label(k + 1), label_of_rank(k + 1);   // index 0 = label -1, j = label j
    for (int r = 0; r <= k; ++r) {
        rank_of_label[order[r]] = r;
        label_of_rank[r] = order[r] == 0 ? -1 : order[r];
    }
    DBuf<int> d_rank_of(k + 1), d_label_of(k + 1), d_dims;
    LGR_CUDA(cudaMemcpyAsync(d_rank_of.p, rank_of_label.data(), sizeof(int) * (k + 1), cudaMemcpyHostToDevice, st));
    LGR_CUDA(cudaMemcpyAsync(d_label_of.p, label_of_rank.data(), sizeof(int) * (k + 1), cudaMemcpyHostToDevice, st));
    if (p.n_dims_use > 0) {
        d_dims.alloc(p.n_dims_use);
        LGR_CUDA(cudaMemcpyAsync(d_dims.p, p.dims_use, sizeof(int) * p.n_dims_use, cudaMemcpyHostToDevice, st));
    }

    // ---- 1. arg-max factor ----
    std::vector<DBuf<int>> rank(d);
    for (int i = 0; i < d; ++i) {
        rank[i].alloc(n[i]);
        DBuf<double> stats(2 * (size_t)k);
        stats.zero(st);
        dim3 g((unsigned)std::min(64, (n[i] + 255) / 256), (unsigned)k);
        k_col_stats<<<g, 256, 0, st>>>(Hdev[i], n[i], k, stats.p);
        k_max_factor<<<(n[i] + 255) / 256, 256, 0, st>>>(Hdev[i], n[i], k, stats.p, p.center, d_dims.p,
                                                         p.n_dims_use > 0 ? p.n_dims_use : k, d_rank_of.p, rank[i].p);
        LGR_CUDA(cudaGetLastError());
        LGR_CUDA(cudaStreamSynchronize(st));
    }
    LGR_CUDA(cudaEventRecord(e1, st));

    // ---- 2. kNN refinement ----
    if (p.refine_knn) {
        // host copies of the points (row-major) for the tree builds, all datasets in parallel
        std::vector<std::vector<double>> pts(d);
        std::vector<AnnBuilder> B(d);
        for (int i = 0; i < d; ++i) {
            std::vector<double> colmajor((size_t)n[i] * k);
            LGR_CUDA(cudaMemcpyAsync(colmajor.data(), Hdev[i], colmajor.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
            LGR_CUDA(cudaStreamSynchronize(st));
            pts[i].resize(colmajor.size());
            for (int c = 0; c < k; ++c)
                for (int r = 0; r < n[i]; ++r) pts[i][(size_t)r * k + c] = colmajor[(size_t)c * n[i] + r];
        }
        const auto t_build0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int i = 0; i < d; ++i) {
            th.emplace_back([&, i] {
                AnnBuilder& b = B[i];
                b.pts = pts[i].data();
                b.dim = k;
                b.pidx.resize(n[i]);
                for (int r = 0; r < n[i]; ++r) b.pidx[r] = r;
                b.blo.assign(k, 0.0);
                b.bhi.assign(k, 0.0);
                for (int c = 0; c < k; ++c) {
                    double mn = pts[i][c], mx = mn;
                    for (int r = 1; r < n[i]; ++r) {
                        const double v = pts[i][(size_t)r * k + c];
                        mn = std::min(mn, v);
                        mx = std::max(mx, v);
                    }
                    b.blo[c] = mn;
                    b.bhi[c] = mx;
                }
                b.nodes.reserve(2 * (size_t)n[i]);
                std::vector<double> lo0 = b.blo, hi0 = b.bhi;
                const int root = b.build(0, n[i], 0);
                (void)root;
                b.blo = lo0;   // the root box (build() restores it, kept explicit)
                b.bhi = hi0;
            });
        }
        for (auto& t : th) t.join();
        if (info) info->tree_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
        cudaEvent_t s0, s1;
        LGR_CUDA(cudaEventCreate(&s0));
        LGR_CUDA(cudaEventCreate(&s1));
        const int kk_all = p.n_neighbors;
        for (int i = 0; i < d; ++i) {
            AnnBuilder& b = B[i];
            LGR_REQUIRE(b.depth_max < ANN_STACK, "quantileNorm: kd-tree deeper than the search stack (degenerate H)");
            const int kk = std::min(kk_all, n[i]);
            DBuf<AnnNode> nodes(b.nodes.size());
            DBuf<double> dpts(pts[i].size()), dlo(k), dhi(k);
            DBuf<int> nn((size_t)n[i] * kk), overflow(1), changed(1), curA(n[i]), curB(n[i]), qorder(n[i]);
            LGR_CUDA(cudaMemcpyAsync(qorder.p, b.pidx.data(), sizeof(int) * n[i], cudaMemcpyHostToDevice, st));
            overflow.zero(st);
            LGR_CUDA(cudaMemcpyAsync(nodes.p, b.nodes.data(), sizeof(AnnNode) * b.nodes.size(), cudaMemcpyHostToDevice, st));
            LGR_CUDA(cudaMemcpyAsync(dpts.p, pts[i].data(), sizeof(double) * pts[i].size(), cudaMemcpyHostToDevice, st));
            LGR_CUDA(cudaMemcpyAsync(dlo.p, b.blo.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st));
            LGR_CUDA(cudaMemcpyAsync(dhi.p, b.bhi.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st));
            const double max_err = (1.0 + p.eps) * (1.0 + p.eps);
            const size_t qsmem = 0;
            LGR_CUDA(cudaEventRecord(s0, st));
            k_ann_search<<<(n[i] + 127) / 128, 128, qsmem, st>>>(
                nodes.p, (int)b.nodes.size() - 1, dpts.p, n[i], k, dlo.p, dhi.p, kk, max_err, qorder.p, nn.p, overflow.p);
            LGR_CUDA(cudaGetLastError());
            LGR_CUDA(cudaEventRecord(s1, st));
            int ov = 0;
            LGR_CUDA(cudaMemcpyAsync(&ov, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            LGR_CUDA(cudaStreamSynchronize(st));
            LGR_REQUIRE(ov == 0, "quantileNorm: kd-tree search stack overflow");
            if (info) {
                float sms = 0.f;
                cudaEventElapsedTime(&sms, s0, s1);
                info->search_ms += sms;
            }
            // in-place vote = fixed point of the sweeps
            LGR_CUDA(cudaMemcpyAsync(curA.p, rank[i].p, sizeof(int) * n[i], cudaMemcpyDeviceToDevice, st));
            int* cur = curA.p;
            int* nxt = curB.p;
            int sweeps = 0;
            for (;;) {
                changed.zero(st);
                k_vote_sweep<<<(n[i] + 255) / 256, 256, 0, st>>>(nn.p, n[i], kk, rank[i].p, cur, nxt, changed.p);
                LGR_CUDA(cudaGetLastError());
                int ch = 0;
                LGR_CUDA(cudaMemcpyAsync(&ch, changed.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                LGR_CUDA(cudaStreamSynchronize(st));
                std::swap(cur, nxt);
                ++sweeps;
                if (!ch) break;
                LGR_REQUIRE(sweeps <= n[i] + 1, "quantileNorm: vote sweeps did not converge");
            }
            if (info) info->vote_sweeps = std::max(info->vote_sweeps, sweeps);
            LGR_CUDA(cudaMemcpyAsync(rank[i].p, cur, sizeof(int) * n[i], cudaMemcpyDeviceToDevice, st));
            LGR_CUDA(cudaStreamSynchronize(st));
        }
        cudaEventDestroy(s0);
        cudaEventDestroy(s1);
    }
    LGR_CUDA(cudaEventRecord(e2, st));

    // ---- 3. cluster cell lists (row order inside a cluster: stable sort by label) ----
    std::vector<DBuf<int>> cells(d);
    std::vector<std::vector<int>> hist(d), off(d);
    for (int i = 0; i < d; ++i) {
        DBuf<int> key(n[i]), key2(n[i]), iota(n[i]), dh(k + 1);
        cells[i].alloc(n[i]);
        dh.zero(st);
        k_rank_to_label<<<(n[i] + 255) / 256, 256, 0, st>>>(rank[i].p, n[i], d_label_of.p, labels_dev[i], key.p);
        k_iota<<<(n[i] + 255) / 256, 256, 0, st>>>(iota.p, n[i]);
        k_hist_labels<<<(n[i] + 255) / 256, 256, 0, st>>>(key.p, n[i], dh.p);
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, iota.p, cells[i].p, n[i], 0, 8, st);
        DBuf<unsigned char> tmp(tmp_bytes);
        LGR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, iota.p, cells[i].p, n[i], 0, 8, st));
        hist[i].resize(k + 1);
        LGR_CUDA(cudaMemcpyAsync(hist[i].data(), dh.p, sizeof(int) * (k + 1), cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaStreamSynchronize(st));
        off[i].assign(k + 2, 0);
        for (int c = 0; c <= k; ++c) off[i][c + 1] = off[i][c] + hist[i][c];
    }

    // ---- 4. base::sample(): replay R's stream when some cluster is larger than maxSample ----
    const int ref = p.reference;
    bool need_rng = false;
    for (int i = 0; i < d; ++i)
        for (int c = 1; c <= k; ++c) need_rng = need_rng || hist[i][c] > p.max_sample;
    std::vector<std::vector<WarpJob>> jobs(d, std::vector<WarpJob>(k));
    std::vector<int> samp;   // arena of sampled positions: per (dataset, cluster): k factors x [s2 draws | s1 draws]
    {
        std::unique_ptr<lgr_rng, void (*)(lgr_rng*)> rng(need_rng ? lgr_r_set_seed(p.seed) : nullptr, lgr_r_free);
        RSampler rs{rng.get()};
        std::vector<int> scratch;
        for (int i = 0; i < d; ++i) {
            for (int c = 1; c <= k; ++c) {
                WarpJob& J = jobs[i][c - 1];
                const int n2 = hist[i][c], n1 = hist[ref][c];
                J = WarpJob{0, 0, 0, 0, off[i][c], off[ref][c], -1, -1};
                if (n1 < p.min_cells || n2 < p.min_cells) continue;
                J.n2 = n2;
                J.n1 = n1;
                J.s2 = std::min(n2, p.max_sample);
                J.s1 = std::min(n1, p.max_sample);
                if (n2 == 1 || !need_rng) continue;
                const bool keep2 = n2 > p.max_sample, keep1 = n1 > p.max_sample;
                const size_t base = samp.size();
                if (keep2 || keep1) samp.resize(base + (size_t)k * (J.s2 + J.s1));
                if (keep2) J.samp2 = (long long)base;
                if (keep1) J.samp1 = (long long)base + J.s2;
                for (int f = 0; f < k; ++f) {   // q2's sample() is called first, then q1's (R/integration.R:1676-1681)
                    int* o = (keep2 || keep1) ? samp.data() + base + (size_t)f * (J.s2 + J.s1) : nullptr;
                    rs.sample(n2, J.s2, scratch, keep2 ? o : nullptr);
                    rs.sample(n1, J.s1, scratch, keep1 ? o + J.s2 : nullptr);
                }
            }
        }
    }
    DBuf<int> d_samp(std::max<size_t>(samp.size(), 1));
    if (!samp.empty()) LGR_CUDA(cudaMemcpyAsync(d_samp.p, samp.data(), sizeof(int) * samp.size(), cudaMemcpyHostToDevice, st));
    const int nq = p.quantiles + 1;
    std::vector<double> probs(nq);
    {
        const double by = 1.0 / p.quantiles;
        for (int t = 0; t < nq; ++t) probs[t] = std::min(0.0 + (double)t * by, 1.0);   // seq(0, 1, by = 1 / quantiles)
    }
    DBuf<double> d_probs(nq);
    LGR_CUDA(cudaMemcpyAsync(d_probs.p, probs.data(), sizeof(double) * nq, cudaMemcpyHostToDevice, st));

    // ---- 5. the warp, dataset by dataset (the reference is warped onto itself at its turn) ----
    int cap = 2;
    for (int i = 0; i < d; ++i)
        for (int c = 0; c < k; ++c) cap = std::max(cap, std::max(jobs[i][c].s2, jobs[i][c].s1));
    int cap2 = 1;
    while (cap2 < cap) cap2 <<= 1;
    const size_t smem = ((size_t)2 * cap2 + 4 * (size_t)nq) * sizeof(double);
    LGR_CUDA(cudaFuncSetAttribute(k_quantile_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int warped = 0;
    for (int i = 0; i < d; ++i) {
        DBuf<WarpJob> dj(k);
        LGR_CUDA(cudaMemcpyAsync(dj.p, jobs[i].data(), sizeof(WarpJob) * k, cudaMemcpyHostToDevice, st));
        for (int c = 0; c < k; ++c) warped += jobs[i][c].n2 > 0;
        dim3 g((unsigned)k, (unsigned)k);
        k_quantile_warp<<<g, 256, smem, st>>>(Hdev[i], n[i], Hdev[ref], n[ref], dj.p, cells[i].p, cells[ref].p, d_samp.p, k, nq,
                                              d_probs.p, cap2);
        LGR_CUDA(cudaGetLastError());
        LGR_CUDA(cudaStreamSynchronize(st));
    }
    LGR_CUDA(cudaEventRecord(e3, st));
    LGR_CUDA(cudaStreamSynchronize(st));
    if (info) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        info->cluster_ms = ms;
        cudaEventElapsedTime(&ms, e1, e2);
        info->knn_ms = ms;
        cudaEventElapsedTime(&ms, e2, e3);
        info->warp_ms = ms;
        info->warped_clusters = warped;
        info->sampled = need_rng ? 1 : 0;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaEventDestroy(e3);
}

}  // namespace lgr
