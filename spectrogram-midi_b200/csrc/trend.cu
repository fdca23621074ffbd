// K5: batched float64 trend filters over f0 series (NaN = unvoiced frame).
//
// Replaces FinancialNoiseFilters.savitzky_golay / kalman_filter / holt_winters and
// multi_filter_consensus (aegis_engine_core_v2/financial_filters.py:25-141, 256-298) and
// FinancialPitchAnalyzer.simple_moving_average / exponential_moving_average / bollinger_bands /
// macd (aegis_engine_core_v2/financial_analysis.py:45-146, 203-226).
//
// One CTA (4 warps) per series.  The scalar recurrences (Kalman, Holt, EMA, MACD) are inherently
// sequential: each runs on one lane of its own warp, concurrently with warp 0's ballot compaction
// of the valid samples; the windowed filters (Savitzky-Golay on the compacted series, SMA,
// rolling sigma, consensus median / sigma) are data parallel over the CTA.  This file is compiled
// with -fmad=false so every recurrence performs the same IEEE operations, in the same order, as
// the reference's Python loops.
#include <cmath>
#include "common.cuh"

namespace aegis {

constexpr int TR_THREADS = 128;

__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ void kalman_series(const double* __restrict__ x, double* __restrict__ out, int n, double q, double r) {
    bool started = false;
    double x_est = 0.0, p_est = 1.0;
    for (int i = 0; i < n; ++i) {
        const double v = x[i];
        if (isnan(v)) { out[i] = nan64(); continue; }
        if (!started) { x_est = v; started = true; }
        const double p_pred = p_est + q;
        const double k = p_pred / (p_pred + r);
        x_est = x_est + k * (v - x_est);
        p_est = (1.0 - k) * p_pred;
        out[i] = x_est;
    }
}

__device__ void holt_series(const double* __restrict__ x, double* __restrict__ out, int n, double alpha, double beta) {
    int i0 = -1, i1 = -1;
    for (int i = 0; i < n && i1 < 0; ++i) {
        if (!isnan(x[i])) { if (i0 < 0) i0 = i; else i1 = i; }
    }
    if (i1 < 0) {  // fewer than two valid points: the reference returns its input unchanged
        for (int i = 0; i < n; ++i) out[i] = x[i];
        return;
    }
    double level = x[i0], trend = x[i1] - x[i0];
    const double one_m_alpha = 1.0 - alpha, one_m_beta = 1.0 - beta;
    for (int i = 0; i < n; ++i) {
        const double v = x[i];
        if (isnan(v)) { out[i] = nan64(); continue; }
        const double forecast = level + trend;
        const double level_new = alpha * v + one_m_alpha * forecast;
        const double trend_new = beta * (level_new - level) + one_m_beta * trend;
        out[i] = level_new;
        level = level_new;
        trend = trend_new;
    }
}

constexpr int BOLL_MAX_WINDOW = 128;

// numpy's float64 add.reduce over a contiguous 1-D array of n <= 128 elements (pairwise_sum's leaf: eight running
// sums combined as a tree, the remainder added one by one; fewer than eight elements are added in order)
__device__ __forceinline__ double numpy_sum_f64(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = __dadd_rn(acc[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                           __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

struct EmaState {
    double alpha, one_m_alpha, prev;
    __device__ EmaState(int span) : alpha(2.0 / static_cast<double>(span + 1)), prev(nan64()) { one_m_alpha = 1.0 - alpha; }
    __device__ double step(double v) {
        if (isnan(v)) { prev = nan64(); return prev; }
        prev = isnan(prev) ? v : alpha * v + one_m_alpha * prev;
        return prev;
    }
};

__global__ void __launch_bounds__(TR_THREADS)
trend_kernel(const aegis_trend_params p) {
    __shared__ int s_nvalid;
    const int series = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const long long off = static_cast<long long>(series) * n;
    const double* __restrict__ x = p.x + off;
    double* compact = p.compact ? p.compact + off : nullptr;
    double* scratch = p.scratch ? p.scratch + off : nullptr;
    const bool want_savgol = p.savgol != nullptr;

    // ---- phase 1: compaction (warp 0) || recurrences (lane 0 of warps 1..3)
    if (warp == 0) {
        int nv = 0;
        if (want_savgol) {
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const double v = (i < n) ? x[i] : nan64();
                const bool ok = (i < n) && !isnan(v);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) compact[nv + __popc(m & ((1u << lane) - 1u))] = v;
                nv += __popc(m);
            }
        }
        if (lane == 0) s_nvalid = nv;
    } else if (lane == 0) {
        if (warp == 1 && p.kalman) kalman_series(x, p.kalman + off, n, p.kalman_q, p.kalman_r);
        if (warp == 2 && p.holt) holt_series(x, p.holt + off, n, p.holt_alpha, p.holt_beta);
        if (warp == 3) {
            if (p.ema) {
                EmaState e(p.ema_span);
                double* out = p.ema + off;
                for (int i = 0; i < n; ++i) out[i] = e.step(x[i]);
            }
            if (p.macd_line) {
                EmaState ef(p.macd_fast), es(p.macd_slow), eg(p.macd_signal);
                for (int i = 0; i < n; ++i) {
                    const double v = x[i];
                    const double line = ef.step(v) - es.step(v);
                    p.macd_line[off + i] = line;
                    const double sig = eg.step(line);
                    if (p.macd_sig) p.macd_sig[off + i] = sig;
                    if (p.macd_hist) p.macd_hist[off + i] = line - sig;
                }
            }
        }
    }
    __syncthreads();
    const int nv = s_nvalid;

    // ---- phase 2: windowed filters, data parallel
    if (want_savgol && nv > p.savgol_window) {  // correlate1d(compact, coeffs, mode='nearest')
        const int w = p.savgol_window, half = w / 2;
        for (int r = tid; r < nv; r += TR_THREADS) {
            double acc = 0.0;
            for (int j = 0; j < w; ++j) {
                const int src = min(max(r + j - half, 0), nv - 1);
                acc += __ldg(p.savgol_coeffs + j) * compact[src];
            }
            scratch[r] = acc;
        }
    }
    for (int pass = 0; pass < 2; ++pass) {  // np.convolve(nan->0, ones(w)/w, 'same'), NaNs restored
        double* out = pass == 0 ? p.sma : p.boll_ma;
        const int w = pass == 0 ? p.sma_window : p.boll_window;
        if (out == nullptr) continue;
        const double kv = 1.0 / static_cast<double>(w);
        for (int i = tid; i < n; i += TR_THREADS) {
            double r = nan64();
            if (!isnan(x[i])) {
                double acc = 0.0;
                const int lo = i - w / 2, hi = i + (w - 1) / 2;
                // numpy's correlate is a dot product per output: ascending index, product rounded before the add
                for (int k = max(lo, 0); k <= min(hi, n - 1); ++k) {
                    const double v = x[k];
                    acc = __dadd_rn(acc, __dmul_rn(isnan(v) ? 0.0 : v, kv));
                }
                r = acc;
            }
            out[off + i] = r;
        }
    }
    __syncthreads();

    // ---- phase 3: scatter Savitzky-Golay back to the original time axis (warp 0)
    if (want_savgol && warp == 0) {
        int cnt = 0;
        const bool active = nv > p.savgol_window;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool ok = (i < n) && !isnan(x[i]);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (i < n) p.savgol[off + i] = (ok && active) ? scratch[cnt + __popc(m & ((1u << lane) - 1u))] : nan64();
            cnt += __popc(m);
        }
    }
    // Bollinger bands: population sigma of the valid points in the trailing window (needs >= 2)
    if (p.boll_upper || p.boll_lower) {
        const int w = p.boll_window;
        for (int i = tid; i < n; i += TR_THREADS) {
            // np.std of the window's valid points, with numpy's summation order (bit-exact: on a steady pitch the
            // band is a few ulps wide and the consumers compare f0 against it)
            double v[BOLL_MAX_WINDOW];
            int cnt = 0;
            for (int k = max(0, i - w + 1); k <= i; ++k) { const double xv = x[k]; if (!isnan(xv)) v[cnt++] = xv; }
            double sd = nan64();
            if (cnt > 1) {
                const double mean = numpy_sum_f64(v, cnt) / static_cast<double>(cnt);
                for (int k = 0; k < cnt; ++k) { const double d = __dsub_rn(v[k], mean); v[k] = __dmul_rn(d, d); }
                sd = sqrt(numpy_sum_f64(v, cnt) / static_cast<double>(cnt));
            }
            const double ma = p.boll_ma[off + i];
            const double dev = p.boll_num_std * sd;
            if (p.boll_upper) p.boll_upper[off + i] = ma + dev;
            if (p.boll_lower) p.boll_lower[off + i] = ma - dev;
        }
    }
    __syncthreads();

    // ---- phase 4: consensus = nanmedian, confidence = 1 / (1 + nanstd) over the three filters
    if (p.consensus || p.consensus_conf) {
        // a filter that does not vote contributes NaN, which nanmedian / nanstd skip (financial_filters.py:270-296)
        const int votes = (p.consensus_mask & 7) ? (p.consensus_mask & 7) : 7;
        for (int i = tid; i < n; i += TR_THREADS) {
            const double r[3] = {(votes & 1) ? p.savgol[off + i] : nan64(), (votes & 2) ? p.kalman[off + i] : nan64(),
                                 (votes & 4) ? p.holt[off + i] : nan64()};
            double v[3];
            int cnt = 0;
            double sum = 0.0;
            for (int j = 0; j < 3; ++j) {
                const bool bad = isnan(r[j]);
                sum += bad ? 0.0 : r[j];
                if (!bad) v[cnt++] = r[j];
            }
            if (p.consensus) {
                double med = nan64();
                if (cnt == 1) med = v[0];
                else if (cnt == 2) med = (fmin(v[0], v[1]) + fmax(v[0], v[1])) / 2.0;
                else if (cnt == 3) med = fmax(fmin(v[0], v[1]), fmin(fmax(v[0], v[1]), v[2]));
                p.consensus[off + i] = med;
            }
            if (p.consensus_conf) {
                const double avg = sum / static_cast<double>(cnt);  // 0/0 -> NaN, as numpy
                double sq = 0.0;
                for (int j = 0; j < 3; ++j) {
                    const double d = isnan(r[j]) ? 0.0 : r[j] - avg;
                    sq += d * d;
                }
                const double sd = sqrt(sq / static_cast<double>(cnt));
                p.consensus_conf[off + i] = 1.0 / (1.0 + sd);
            }
        }
    }
}

}  // namespace aegis

extern "C" int aegis_trend_filters(const aegis_trend_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->x, "aegis_trend_filters: x must be set");
    AEGIS_REQUIRE(p->n_series >= 0 && p->n >= 0, "aegis_trend_filters: negative size");
    if (p->savgol) {
        AEGIS_REQUIRE(p->compact && p->scratch && p->savgol_coeffs, "aegis_trend_filters: savgol needs compact/scratch workspaces and coefficients");
        AEGIS_REQUIRE(p->savgol_window >= 1 && (p->savgol_window & 1), "aegis_trend_filters: savgol window must be odd");
    }
    if (p->consensus || p->consensus_conf) {
        const int votes = (p->consensus_mask & 7) ? (p->consensus_mask & 7) : 7;
        AEGIS_REQUIRE((!(votes & 1) || p->savgol) && (!(votes & 2) || p->kalman) && (!(votes & 4) || p->holt),
                      "aegis_trend_filters: consensus needs the outputs of the filters that vote (consensus_mask=%d)", p->consensus_mask);
    }
    if (p->boll_upper || p->boll_lower) AEGIS_REQUIRE(p->boll_ma && p->boll_window >= 1 && p->boll_window <= BOLL_MAX_WINDOW, "aegis_trend_filters: bands need boll_ma and 1 <= boll_window <= 128");
    if (p->sma) AEGIS_REQUIRE(p->sma_window >= 1, "aegis_trend_filters: bad sma_window");
    if (p->macd_sig || p->macd_hist) AEGIS_REQUIRE(p->macd_line != nullptr, "aegis_trend_filters: MACD signal/hist need macd_line");
    if (p->n_series == 0 || p->n == 0) return 0;
    trend_kernel<<<p->n_series, TR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p);
    return check_launch("aegis_trend_filters");
}
