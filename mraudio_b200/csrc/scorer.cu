// Moment-retrieval scorer: per-query average precision at 10 IoU thresholds + top-1 IoU, one thread per query.
//
// Restates, in fp64 with numpy's operation order / nan semantics, the reference's
//   compute_temporal_iou_batch_cross     eval/mr_utils.py:40-67   (true-union IoU, 0/0 -> nan)
//   compute_temporal_iou_batch_paired    eval/mr_utils.py:16-37   (hull "union", 0 where union == 0)
//   compute_average_precision_detection  eval/mr_utils.py:89-171  (greedy matching in list order, VOC-2011 AP)
//   interpolated_precision_recall        eval/mr_utils.py:70-86
//   compute_mr_r1 (per-query part)       eval/mr_eval.py:97-131
// The work per query is a few hundred scalar fp64 operations on <= a few dozen windows: latency-bound integer/fp64
// code, not tensor work.  Queries are independent, so the grid is one thread per query with coalesced reads of the
// padded [Q, Pmax, 2] / [Q, Gmax, 2] window arrays.
#include "common.h"

namespace mra {
namespace {

constexpr int NT = MRA_NUM_IOU_THDS;
constexpr int GMAX = 64;
constexpr int PMAX = 256;

// numpy's sort order for floats: nan sorts last.  a "less than" b ?
__device__ __forceinline__ bool np_lt(double a, double b) { return a < b || (b != b && a == a); }

// numpy pairwise_sum for n <= 128 (DOUBLE_pairwise_sum): < 8 -> plain loop, else 8 interleaved partial sums.
__device__ double np_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

__device__ __forceinline__ double cross_iou(double p0, double p1, double g0, double g1) {
    const double area1 = p1 - p0, area2 = g1 - g0;
    const double left = fmax(p0, g0), right = fmin(p1, g1);
    double inter = right - left;
    if (inter < 0.0) inter = 0.0;  // np.clip(x, 0, None)
    const double uni = area1 + area2 - inter;
    return inter / uni;            // 0/0 -> nan, x/0 -> inf exactly like numpy
}

__global__ void __launch_bounds__(128)
mr_score_kernel(const double* __restrict__ pred, const int32_t* __restrict__ n_pred, const double* __restrict__ gt,
                const int32_t* __restrict__ n_gt, const double* __restrict__ thds, int Q, int Pmax, int Gmax,
                double* __restrict__ out_ap, double* __restrict__ out_iou, uint8_t* __restrict__ out_invalid) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int np_ = n_pred[q], ng = n_gt[q];
    const double* P = pred + static_cast<int64_t>(q) * Pmax * 2;
    const double* G = gt + static_cast<int64_t>(q) * Gmax * 2;
    double thd[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) thd[t] = thds[t];

    // ---------------- AP: greedy matching (eval/mr_utils.py:126-159)
    unsigned long long lock[NT];
    unsigned long long tp[NT][PMAX / 64];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        lock[t] = 0ull;
        for (int w = 0; w < PMAX / 64; ++w) tp[t][w] = 0ull;
    }
    double iou[GMAX];
    int order[GMAX];
    bool tie_ambiguous = false;
    for (int idx = 0; idx < np_; ++idx) {
        const double p0 = P[2 * idx], p1 = P[2 * idx + 1];
        for (int j = 0; j < ng; ++j) iou[j] = cross_iou(p0, p1, G[2 * j], G[2 * j + 1]);
        // stable ascending insertion sort (nan last), visited in reverse == tiou_arr.argsort()[::-1]
        for (int j = 0; j < ng; ++j) {
            const double vj = iou[j];
            int k = j;
            while (k > 0 && np_lt(vj, iou[order[k - 1]])) {
                order[k] = order[k - 1];
                --k;
            }
            order[k] = j;
        }
        // Two DIFFERENT ground-truth windows with exactly the same IoU (>= the lowest threshold, or nan) for one prediction:
        // which of them is matched first is decided by the tie order of numpy's argsort, which is an unstable SIMD sort on
        // AVX-512 / AVX2 builds for >= 4 elements, i.e. the reference's own result is platform-dependent for this query.
        // This kernel uses the stable order; the query is flagged (bit 1 of out_invalid) so that callers can tell.
        for (int k = 0; k + 1 < ng; ++k) {
            const int j1 = order[k], j2 = order[k + 1];
            const double v1 = iou[j1], v2 = iou[j2];
            const bool eq = v1 == v2 || (v1 != v1 && v2 != v2);
            if (eq && !(v1 < thd[0]) && (G[2 * j1] != G[2 * j2] || G[2 * j1 + 1] != G[2 * j2 + 1])) tie_ambiguous = true;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            for (int k = ng - 1; k >= 0; --k) {
                const int j = order[k];
                if (iou[j] < thd[t]) break;                 // false positive (nan < thd is false: nan matches)
                if ((lock[t] >> j) & 1ull) continue;        // this GT already taken at this threshold
                tp[t][idx >> 6] |= 1ull << (idx & 63);
                lock[t] |= 1ull << j;
                break;
            }
        }
    }
    // ---------------- AP: cumulative precision / recall and the VOC-2011 envelope (:161-170, :70-86)
    double terms[GMAX + 2];
    double mprec[PMAX + 2];
    for (int t = 0; t < NT; ++t) {
        double ap = 0.0;
        if (np_ > 0) {
            // mprecision = [0, precision..., 0]; running max from the right
            double tpc = 0.0;
            mprec[0] = 0.0;
            for (int i = 0; i < np_; ++i) {
                const bool is_tp = (tp[t][i >> 6] >> (i & 63)) & 1ull;
                if (is_tp) tpc += 1.0;
                mprec[i + 1] = tpc / (tpc + (static_cast<double>(i + 1) - tpc));
            }
            mprec[np_ + 1] = 0.0;
            for (int i = np_; i >= 0; --i) mprec[i] = mprec[i] > mprec[i + 1] ? mprec[i] : mprec[i + 1];
            // mrecall = [0, recall..., 1]; sum over positions where it changes
            int nterms = 0;
            double prev = 0.0;
            tpc = 0.0;
            const double npos = static_cast<double>(ng);
            for (int i = 1; i <= np_ + 1; ++i) {
                double cur;
                if (i <= np_) {
                    if ((tp[t][(i - 1) >> 6] >> ((i - 1) & 63)) & 1ull) tpc += 1.0;
                    cur = tpc / npos;
                } else {
                    cur = 1.0;
                }
                if (cur != prev) terms[nterms++] = (cur - prev) * mprec[i];
                prev = cur;
            }
            ap = np_sum(terms, nterms);
        }
        out_ap[static_cast<int64_t>(q) * NT + t] = ap;
    }
    // ---------------- R1: top-1 window vs the GT with the highest cross IoU, paired (hull) IoU (eval/mr_eval.py:101-118)
    {
        const double p0 = P[0], p1 = P[1];
        int best = 0;
        double bv = cross_iou(p0, p1, G[0], G[1]);
        if (!(bv != bv)) {  // np.argmax: first maximum; the first nan wins outright
            for (int j = 1; j < ng; ++j) {
                const double v = cross_iou(p0, p1, G[2 * j], G[2 * j + 1]);
                if (v != v) { best = j; break; }
                if (v > bv) { bv = v; best = j; }
            }
        }
        const double g0 = G[2 * best], g1 = G[2 * best + 1];
        double inter = fmin(p1, g1) - fmax(p0, g0);
        if (inter < 0.0) inter = 0.0;
        const double uni = fmax(p1, g1) - fmin(p0, g0);
        out_iou[q] = (uni != 0.0) ? inter / uni : 0.0;
        out_invalid[q] = ((p0 == -1.0 || p1 == -1.0) ? 1 : 0) | (tie_ambiguous ? 2 : 0);
    }
}

}  // namespace

int launch_mr_score(const double* pred, const int32_t* n_pred, const double* gt, const int32_t* n_gt, const double* thds,
                    int Q, int Pmax, int Gmax, double* out_ap, double* out_iou, uint8_t* out_invalid, cudaStream_t s) {
    MRA_REQUIRE(Q > 0, "mr_score: no queries");
    MRA_REQUIRE(Pmax >= 1 && Pmax <= PMAX, "mr_score: Pmax %d out of range [1, %d]", Pmax, PMAX);
    MRA_REQUIRE(Gmax >= 1 && Gmax <= GMAX, "mr_score: Gmax %d out of range [1, %d]", Gmax, GMAX);
    mr_score_kernel<<<(Q + 127) / 128, 128, 0, s>>>(pred, n_pred, gt, n_gt, thds, Q, Pmax, Gmax, out_ap, out_iou, out_invalid);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mra
