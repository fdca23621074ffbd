// K8: the v2 ("financial") logic filter -- frames -> note events -> RSI ghost-note filter -> key / scale / chord pass.
//
// Replaces get_midi_events_financial with use_financial=True (aegis_engine_core_v2/midi_logic_financial.py:117-388,
// called from aegis_engine_financial.py:160-171) together with the pieces of FinancialPitchAnalyzer and
// HarmonicAnalyzer it calls (aegis_engine_core_v2/financial_analysis.py:148-196,228-271,277-364;
// aegis_engine_core_v2/harmonic_analysis.py:46-283).  The numeric series it reads -- consensus trend, Bollinger bands
// of f0, MACD of the semitone series -- come from K5 (trend.cu), launched by the caller between the two entry points.
//
//   prepare  (frame parallel)  f0_clean = f0 where voiced else NaN; semitones = hz_to_midi(f0_clean); rms -> dB
//   frames   (frame parallel)  one 16-byte record per frame: combined confidence, MIDI note of the trend (or -1),
//                              band side of f0 (above / inside / below / no pitch), slide label, velocity
//   events   (one warp per clip) adaptive threshold with numpy's summation order; the band-crossing counter and
//                              the note state machine in one walk over the records; duration filter and sustain merge
//                              in streaming form; then, on the clip's few events, the RSI density filter, the key
//                              histogram, the out-of-scale filter and the 2-second chord windows.
//
// This file is compiled with -fmad=false: every product and sum is rounded as in the reference's numpy expressions.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "notes_common.cuh"

namespace aegis {

constexpr int NF_THREADS = 256;

struct __align__(16) FinFrame {
    double combined;          // 0.5 voiced_prob + 0.5 / (1 + band width)
    short note;               // int(round(hz_to_midi(trend))) on frames that pass every gate, else -1
    signed char side;         // f0 against its bands: 1 above, 0 inside, -1 below, -2 no pitch (label None)
    unsigned char slide;      // 0 None, 1 normal, 2 slide_up, 3 slide_down
    unsigned char velocity;
    unsigned char _pad[3];
};
static_assert(sizeof(FinFrame) == 16, "frame record is one 16-byte load");

struct FinScratch {
    float* rms_max;
    float* rms_db;
    FinFrame* frames;
    double* compact;
};

__host__ __device__ inline long long align16(long long v) { return (v + 15) & ~15LL; }

inline FinScratch carve_scratch(void* base, int n_clips, int n_frames) {
    const long long frames = static_cast<long long>(n_clips) * n_frames;
    unsigned char* b = static_cast<unsigned char*>(base);
    FinScratch s;
    s.frames = reinterpret_cast<FinFrame*>(b);
    b += frames * static_cast<long long>(sizeof(FinFrame));
    s.compact = reinterpret_cast<double*>(b);
    b += frames * 8;
    s.rms_db = reinterpret_cast<float*>(b);
    b += align16(frames * 4);
    s.rms_max = reinterpret_cast<float*>(b);
    return s;
}

__global__ void __launch_bounds__(NF_THREADS)
fin_prepare_kernel(const aegis_fin_params p, const float* __restrict__ rms_max, float* __restrict__ rms_db) {
    const int blocks_per_clip = (p.n_frames + NF_THREADS - 1) / NF_THREADS;
    const int clip = blockIdx.x / blocks_per_clip;
    const int t = (blockIdx.x - clip * blocks_per_clip) * NF_THREADS + threadIdx.x;
    if (t >= p.n_frames) return;
    const long long i = static_cast<long long>(clip) * p.n_frames + t;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double f = p.voiced_flag[i] != 0 ? p.f0[i] : nan;             // np.where(voiced_flag, f0, np.nan) (:158)
    p.f0_clean[i] = f;
    // librosa.hz_to_midi on the valid frames (financial_analysis.py:244-248)
    p.semitones[i] = isnan(f) ? nan : 12.0 * (log2(f) - log2(440.0)) + 69.0;
    rms_db[i] = rms_db_f32(p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t], rms_max[clip]);
}

__global__ void __launch_bounds__(NF_THREADS)
fin_frames_kernel(const aegis_fin_params p, const float* __restrict__ rms_db, FinFrame* __restrict__ frames) {
    const int blocks_per_clip = (p.n_frames + NF_THREADS - 1) / NF_THREADS;
    const int clip = blockIdx.x / blocks_per_clip;
    const int t = (blockIdx.x - clip * blocks_per_clip) * NF_THREADS + threadIdx.x;
    if (t >= p.n_frames) return;
    const long long i = static_cast<long long>(clip) * p.n_frames + t;
    const double f = p.f0_clean[i];
    const double up = p.boll_upper[i], lo = p.boll_lower[i];
    const bool pitched = !isnan(f);
    FinFrame r;
    // band position (financial_analysis.py:172-177); comparisons against NaN bands are false -> inside
    r.side = static_cast<signed char>(!pitched ? -2 : (f > up ? 1 : (f < lo ? -1 : 0)));
    // Bollinger confidence (financial_analysis.py:402-416) and its blend with the pYIN probability (:169)
    const double bw = up - lo;
    double conf = 0.0;
    if (pitched && !isnan(bw)) conf = bw > 0 ? 1.0 / (1.0 + bw) : 1.0;
    r.combined = p.voiced_prob[i] * 0.5 + conf * 0.5;
    // slide label (financial_analysis.py:255-269)
    const double ml = p.macd_line[i], mh = p.macd_hist[i];
    r.slide = isnan(ml) ? 0 : ((ml > p.slide_threshold && mh > 0) ? 2 : ((ml < -p.slide_threshold && mh < 0) ? 3 : 1));
    // gates of the event loop (:204-216)
    const float energy = rms_db[i];
    const double tr = p.trend[i];
    bool voiced = !isnan(tr) && p.voiced_flag[i] != 0;
    if (energy < p.noise_gate_db) voiced = false;
    const bool active = voiced && tr > 0 && p.rake_mask[i] == 0;
    int note = -1;
    if (active) {
        const double m = rint(12.0 * (log2(tr) - log2(440.0)) + 69.0);   // round half to even, as Python's round()
        note = static_cast<int>(fmin(fmax(m, 0.0), 32767.0));
    }
    r.note = static_cast<short>(note);
    r.velocity = static_cast<unsigned char>(velocity_from_db(energy));
    r._pad[0] = r._pad[1] = r._pad[2] = 0;
    frames[i] = r;
}

// numpy's float64 pairwise_sum: blocks of at most 128 elements are summed with eight running sums (combined as a
// tree, the remainder added in order); longer arrays are split at n/2 rounded down to a multiple of 8, recursively.
__device__ __forceinline__ double numpy_pairwise_leaf_f64(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += a[i + j];
    }
    double res = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// the recursion unrolled onto an explicit stack (device recursion has no static stack bound)
__device__ double numpy_pairwise_f64(const double* a, int n) {
    constexpr int DEPTH = 28;                 // n / 2^DEPTH <= 128 for every int n
    int off[DEPTH], len[DEPTH], state[DEPTH];
    double left[DEPTH];
    int sp = 0;
    off[0] = 0; len[0] = n; state[0] = 0;
    double result = 0.0;
    while (sp >= 0) {
        bool done = false;
        if (len[sp] <= 128) {
            result = numpy_pairwise_leaf_f64(a + off[sp], len[sp]);
            done = true;
        } else {
            int n2 = len[sp] / 2;
            n2 -= n2 % 8;
            if (state[sp] == 0) {             // descend into the left half
                state[sp] = 1;
                off[sp + 1] = off[sp]; len[sp + 1] = n2; state[sp + 1] = 0;
                ++sp;
            } else if (state[sp] == 1) {      // left half is in `result`: descend into the right half
                left[sp] = result;
                state[sp] = 2;
                off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; state[sp + 1] = 0;
                ++sp;
            } else {
                result = left[sp] + result;
                done = true;
            }
        }
        if (done) --sp;
    }
    return result;
}

struct FinEv {
    int note, start, end, velocity;
    int track, art, slide;
    double confidence;
};

__device__ __forceinline__ void write_fin_event(aegis_fin_event* dst, const FinEv& e) {
    dst->note = e.note;
    dst->start = e.start;
    dst->end = e.end;
    dst->velocity = e.velocity;
    dst->track = static_cast<uint8_t>(e.track);
    dst->technique = static_cast<uint8_t>(e.art);
    dst->slide = static_cast<uint8_t>(e.slide);
    dst->harmonic_valid = -1;
    dst->_pad[0] = dst->_pad[1] = dst->_pad[2] = dst->_pad[3] = 0;
    dst->confidence = e.confidence;
}

__constant__ unsigned short c_scale_mask[3] = {
    // pitch-class sets relative to the root: major, natural minor, blues (harmonic_analysis.py:22-31)
    (1 << 0) | (1 << 2) | (1 << 4) | (1 << 5) | (1 << 7) | (1 << 9) | (1 << 11),
    (1 << 0) | (1 << 2) | (1 << 3) | (1 << 5) | (1 << 7) | (1 << 8) | (1 << 10),
    (1 << 0) | (1 << 3) | (1 << 5) | (1 << 6) | (1 << 7) | (1 << 10),
};

__device__ __forceinline__ bool in_scale(int pc, int root, int mode) {
    return (c_scale_mask[mode] >> ((pc - root + 12) % 12)) & 1;
}

// The RSI verdicts without walking the series.  The density series is 0 / 1 with at most two edges per event, and
// between edges Wilder's recurrence is a pure decay g <- fl(fl(13 g) / 14): n such steps are g (13/14)^n up to a relative
// error of 2 n ulp.  The reference's result depends on the RSI only through `rsi < threshold` at the event starts, so the
// closed form decides every verdict whose distance from the threshold exceeds a bound on its own error (4 L ulp relative
// on g and l, hence 50x that on the RSI, plus slack); anything closer -- or in the subnormal range, where the relative
// bound does not hold -- returns false and the caller walks the series exactly.  Same verdicts, ~n_events steps instead
// of 10 T sequential double divisions (12 920 for a 30 s clip: ~1 ms on one lane).
__device__ bool rsi_verdicts_closed_form(aegis_fin_event* out, int n_ev, long long L, double g, double l, double thr) {
    constexpr long long period = 14;
    constexpr double TINY = 1e-280;
    const double ratio = 13.0 / 14.0;
    const double rel = 4.0 * static_cast<double>(L) * 1.2e-16 + 1e-12;
    const double margin = 50.0 * rel + 1e-9;
    bool had_gain = g > 0, had_loss = l > 0;
    long long pos = period;
    auto advance = [&](long long i, double gain, double loss) {   // steps pos+1 .. i-1 decay, step i takes (gain, loss)
        const long long n = i - 1 - pos;
        if (n > 0) {
            const double f = pow(ratio, static_cast<double>(n));
            g = g * f;
            l = l * f;
        }
        g = (g * (period - 1) + gain) / period;
        l = (l * (period - 1) + loss) / period;
        pos = i;
    };
    for (int k = 0; k < n_ev; ++k) {
        const long long is = 10LL * out[k].start, ie = 10LL * out[k].end;
        const bool solid = ie > is;                    // an empty span leaves no edge in the series
        if (is > period && is < L) {
            advance(is, solid ? 1.0 : 0.0, 0.0);
            had_gain = had_gain || solid;
            const bool g_zero = !had_gain, l_zero = !had_loss;
            const bool g_tiny = had_gain && g < TINY, l_tiny = had_loss && l < TINY;
            int keep;
            if (l_zero) keep = 100.0 < thr;                                   // l == 0: the reference's RSI is 100
            else if (l_tiny) { if (g_tiny || thr > 99.0) return false; keep = 0; }      // RSI within 1e-200 of 100
            else if (g_zero) keep = 0.0 < thr;                                // 100 - 100 / (1 + 0)
            else if (g_tiny) { if (thr < 1.0) return false; keep = 1; }       // RSI within 1e-200 of 0
            else {
                const double rsi = 100.0 - (100.0 / (1.0 + g / l));
                if (!(fabs(rsi - thr) > margin)) return false;
                keep = rsi < thr;
            }
            out[k]._pad[0] = static_cast<unsigned char>(keep);
        }
        if (solid && ie > period && ie < L) {
            advance(ie, 0.0, 1.0);
            had_loss = true;
        }
    }
    return true;
}

// One warp per clip: the frame records are staged in shared memory with coalesced loads, the compaction and the squared
// deviations are lane parallel, the ordered sums and the note state machine run redundantly on every lane (warp-uniform
// control flow, shared-memory latency), lane 0 writes the events and then runs the per-event passes alone.  (One THREAD
// per clip, as in round 1, serialised 64 unrelated state machines per block on each other's branches: 6.4 ms.)
constexpr int FE_STAGE_MAX_FRAMES = 9000;   // 24 B per frame of dynamic shared memory

template <bool STAGED>
__global__ void __launch_bounds__(32)
fin_events_kernel(const aegis_fin_params p, const FinFrame* __restrict__ frames, double* __restrict__ compact, int exact_rsi) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    const int clip = blockIdx.x;
    const int lane = threadIdx.x;
    const int T = p.n_frames;
    const FinFrame* fr = frames + static_cast<long long>(clip) * T;
    double* c = compact + static_cast<long long>(clip) * T;
    if (STAGED) {
        FinFrame* fr_s = reinterpret_cast<FinFrame*>(fe_smem);
        for (int t = lane; t < T; t += 32) fr_s[t] = fr[t];
        fr = fr_s;
        c = reinterpret_cast<double*>(fr_s + T);
        __syncwarp();
    }
    aegis_fin_event* out = p.events + static_cast<long long>(clip) * p.max_events;

    // ---- threshold (midi_logic_financial.py:77-114,172-176)
    double thr = p.confidence_threshold;
    if (isnan(thr)) {
        int m = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {          // order-preserving compaction of the positive confidences
            const int t = t0 + lane;
            const double v = t < T ? fr[t].combined : 0.0;
            const bool pos = v > 0;
            const unsigned mask = __ballot_sync(0xffffffffu, pos);
            if (pos) c[m + __popc(mask & ((1u << lane) - 1u))] = v;
            m += __popc(mask);
        }
        __syncwarp();
        if (m == 0) {
            thr = 0.5;
        } else {
            const double mean = numpy_pairwise_f64(c, m) / static_cast<double>(m);
            __syncwarp();
            for (int i = lane; i < m; i += 32) { const double d = c[i] - mean; c[i] = d * d; }
            __syncwarp();
            const double sd = sqrt(numpy_pairwise_f64(c, m) / static_cast<double>(m));
            thr = fmin(fmax(mean - sd, 0.3), 0.8);
        }
    }
    if (lane == 0 && p.threshold_out != nullptr) p.threshold_out[clip] = thr;

    // ---- phases 2 and 3a: runs of one note, duration filter, sustain merge (:204-328)
    int n_out = 0;
    bool have_pending = false, open = false;
    FinEv pending{}, cur{};
    auto close = [&](const FinEv& e) {
        if (e.end - e.start < p.min_note_frames) return;                                  // :302
        if (have_pending) {
            // the label of a pitched frame is never None, so `not technique` only holds for label-less input
            if (e.note == pending.note && (e.start - pending.end) <= p.sustain_frames && pending.art == 0) {
                pending.end = e.end;                                                      // :314-318
                return;
            }
            if (lane == 0 && n_out < p.max_events) write_fin_event(out + n_out, pending);
            ++n_out;
        }
        pending = e;
        have_pending = true;
    };
    int prev_side = 0, crossings = 0;
    for (int t = 0; t < T; ++t) {
        const FinFrame r = fr[t];
        int label = 0;
        if (r.side != -2) {   // financial_analysis.py:179-194: the counter runs over the pitched frames only
            crossings = (r.side != prev_side && prev_side != 0) ? crossings + 1 : 0;
            label = crossings >= 2 ? 3 : (r.side > 0 ? 2 : (r.side < 0 ? 4 : 1));
            prev_side = r.side;
        }
        const int n = r.note;
        if (open && n != cur.note) {
            close(cur);
            open = false;
        }
        if (n >= 0) {
            if (!open) {
                cur.note = n;
                cur.start = t;
                cur.confidence = r.combined;
                cur.velocity = r.velocity;
                cur.track = r.combined >= thr ? 1 : 0;
                cur.art = label;
                cur.slide = r.slide;
                open = true;
            } else if (label > 1) {
                cur.art = label;                                                          // :240-242
            }
            cur.end = t;
        }
    }
    if (open) close(cur);
    if (have_pending) {
        if (lane == 0 && n_out < p.max_events) write_fin_event(out + n_out, pending);
        ++n_out;
    }
    if (lane != 0) return;   // the passes below touch the clip's few events only
    if (p.key_out != nullptr) p.key_out[clip] = -1;
    if (p.key_confidence_out != nullptr) p.key_confidence_out[clip] = 0.0;
    if (n_out > p.max_events) {   // overflow: report the count, the host raises
        p.n_events[clip] = n_out;
        return;
    }
    int n_ev = n_out;

    // ---- phase 3b: RSI of the note-density series (financial_analysis.py:277-364), only with more than ten events.
    // `start` / `end` are frame numbers and the bins are tenths of a frame; events are disjoint and ordered, so the
    // density is 1 on [10 start, 10 end) and 0 elsewhere and the RSI recurrence can be streamed.
    if (n_ev > 10) {
        int max_end = 0;
        for (int k = 0; k < n_ev; ++k) max_end = max(max_end, out[k].end);
        const long long L = 10LL * max_end;                 // len(np.linspace(0, max_time, int(max_time * 10)))
        constexpr int period = 14;
        // verdicts live in the records' first pad byte until the compaction below.  Default: RSI is 50 before the
        // first full period (and everywhere on a series shorter than it); an event starting past the series is kept.
        for (int k = 0; k < n_ev; ++k) out[k]._pad[0] = (10LL * out[k].start < L) ? (50.0 < p.rsi_threshold ? 1 : 0) : 1;
        if (L - 1 >= period) {
            int cov = 0;                                     // event whose span may cover the current bin
            auto density = [&](long long j) -> double {
                while (cov < n_ev && j >= 10LL * out[cov].end) ++cov;
                return (cov < n_ev && j >= 10LL * out[cov].start) ? 1.0 : 0.0;
            };
            // means of the first 14 gains / losses (whole numbers: exact in any order)
            double d_prev = density(0), g = 0.0, l = 0.0;
            for (long long j = 1; j <= period; ++j) {
                const double d = density(j), delta = d - d_prev;
                if (delta > 0) g += delta;
                if (delta < 0) l += -delta;
                d_prev = d;
            }
            g = g / period;
            l = l / period;
            if (exact_rsi != 0 || !rsi_verdicts_closed_form(out, n_ev, L, g, l, p.rsi_threshold)) {
                int nxt = 0;                                 // next event to judge
                for (long long i = period; i < L; ++i) {
                    if (i > period) {                        // Wilder smoothing with deltas[i - 1] = data[i] - data[i - 1]
                        const double d = density(i), delta = d - d_prev;
                        const double gain = delta > 0 ? delta : 0.0, loss = delta < 0 ? -delta : 0.0;
                        g = (g * (period - 1) + gain) / period;
                        l = (l * (period - 1) + loss) / period;
                        d_prev = d;
                    }
                    while (nxt < n_ev && 10LL * out[nxt].start < i) ++nxt;
                    if (nxt < n_ev && 10LL * out[nxt].start == i) {
                        const double rsi = (l == 0) ? 100.0 : 100.0 - (100.0 / (1.0 + g / l));
                        out[nxt]._pad[0] = rsi < p.rsi_threshold ? 1 : 0;
                        ++nxt;
                    }
                }
            }
        }
        int kept = 0;
        for (int k = 0; k < n_ev; ++k) {
            if (out[k]._pad[0]) {
                if (kept != k) out[kept] = out[k];
                out[kept]._pad[0] = 0;
                ++kept;
            }
        }
        n_ev = kept;
    }

    // ---- phase 4: key, out-of-scale notes, chord context (:334-384; harmonic_analysis.py:46-283)
    if (p.use_harmonic_filter != 0 && n_ev > 5) {
        double hist[12];
        for (int k = 0; k < 12; ++k) hist[k] = 0.0;
        for (int k = 0; k < n_ev; ++k) hist[((out[k].note % 12) + 12) % 12] += 1.0;
        const double total = static_cast<double>(n_ev) + 1e-6;     // np.sum of whole numbers is exact
        for (int k = 0; k < 12; ++k) hist[k] = hist[k] / total;
        int best_root = 0, best_mode = 0;
        double best = 0.0;
        for (int root = 0; root < 12; ++root) {
            for (int mode = 0; mode < 3; ++mode) {
                double s = 0.0;
                for (int iv = 0; iv < 12; ++iv)                    // ascending intervals, as the reference lists them
                    if ((c_scale_mask[mode] >> iv) & 1) s += hist[(root + iv) % 12];
                if (s > best) { best = s; best_root = root; best_mode = mode; }
            }
        }
        int removed = 0;
        for (int k = 0; k < n_ev; ++k) {
            const int pc = ((out[k].note % 12) + 12) % 12;
            int dmin = 12;
            for (int s = 0; s < 12; ++s) {
                if (!in_scale(s, best_root, best_mode)) continue;
                const int d = abs(pc - s);
                dmin = min(dmin, min(d, 12 - d));
            }
            const bool bad = dmin > p.harmonic_tolerance;
            out[k]._pad[0] = bad ? 1 : 0;
            removed += bad ? 1 : 0;
        }
        if (removed > 0) {
            int kept = 0;
            for (int k = 0; k < n_ev; ++k) {
                if (out[k]._pad[0] == 0) {
                    if (kept != k) out[kept] = out[k];
                    out[kept].harmonic_valid = 1;
                    ++kept;
                }
            }
            for (int k = 0; k < kept; ++k) out[k]._pad[0] = 0;
            n_ev = kept;
            if (p.key_out != nullptr) p.key_out[clip] = best_root | (best_mode << 8);
            if (p.key_confidence_out != nullptr) p.key_confidence_out[clip] = best;
            if (n_ev > 0) {
                const double frame_ms = static_cast<double>(p.hop) / p.sr;
                auto time_of = [&](int k) { return out[k].start * frame_ms * 1000; };   // start * (hop / sr) * 1000
                double max_time = time_of(0);
                for (int k = 1; k < n_ev; ++k) max_time = fmax(max_time, time_of(k));
                const int t_end = static_cast<int>(max_time);
                for (int t0 = 0; t0 < t_end; t0 += 2000) {
                    // pitch classes of the window, most frequent first seen = root
                    int cnt[12];
                    for (int k = 0; k < 12; ++k) cnt[k] = 0;
                    int first = -1, last = -1;
                    for (int k = 0; k < n_ev; ++k) {
                        const double tm = time_of(k);
                        if (tm >= t0 && tm < t0 + 2000) {
                            ++cnt[((out[k].note % 12) + 12) % 12];
                            if (first < 0) first = k;
                            last = k;
                        }
                    }
                    if (first < 0) continue;
                    int cmax = 0;
                    for (int k = 0; k < 12; ++k) cmax = max(cmax, cnt[k]);
                    int root = -1;
                    for (int k = first; k <= last && root < 0; ++k) {
                        const double tm = time_of(k);
                        const int pc = ((out[k].note % 12) + 12) % 12;
                        if (tm >= t0 && tm < t0 + 2000 && cnt[pc] == cmax) root = pc;
                    }
                    const int third = cnt[(root + 4) % 12] > 0 ? 4 : (cnt[(root + 3) % 12] > 0 ? 3 : 0);
                    if (third == 0) continue;                                   // quality unknown: no penalty
                    const int tone1 = (root + third) % 12, tone2 = (root + 7) % 12;
                    for (int k = first; k <= last; ++k) {
                        const double tm = time_of(k);
                        if (!(tm >= t0 && tm < t0 + 2000)) continue;
                        const int pc = ((out[k].note % 12) + 12) % 12;
                        if (pc != root && pc != tone1 && pc != tone2)
                            out[k].confidence = out[k].confidence * (in_scale(pc, best_root, best_mode) ? 0.8 : 0.5);
                    }
                }
                for (int k = 0; k < n_ev; ++k) out[k].track = out[k].confidence >= thr ? 1 : 0;   // :380-381
            }
        }
    }
    p.n_events[clip] = n_ev;
}

}  // namespace aegis

static unsigned frame_blocks(const aegis_fin_params* p) {
    return static_cast<unsigned>(static_cast<long long>((p->n_frames + aegis::NF_THREADS - 1) / aegis::NF_THREADS) * p->n_clips);
}

static int fin_check_common(const aegis_fin_params* p, const char* who) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "%s: null params", who);
    AEGIS_REQUIRE(p->n_clips >= 0 && p->n_frames >= 0, "%s: negative size", who);
    AEGIS_REQUIRE(p->rake_mask && p->f0 && p->voiced_flag && p->voiced_prob && p->rms, "%s: inputs missing", who);
    AEGIS_REQUIRE(p->rms_clip_stride >= p->n_frames && p->hop > 0 && p->sr > 0, "%s: bad rms stride / hop / sr", who);
    AEGIS_REQUIRE(p->f0_clean && p->semitones && p->scratch, "%s: f0_clean / semitones / scratch missing", who);
    AEGIS_REQUIRE(static_cast<long long>((p->n_frames + aegis::NF_THREADS - 1) / aegis::NF_THREADS) * p->n_clips < (1LL << 31), "%s: too many frames for one launch", who);
    return 0;
}

extern "C" int aegis_fin_prepare(const aegis_fin_params* p, void* stream) {
    using namespace aegis;
    if (int rc = fin_check_common(p, "aegis_fin_prepare")) return rc;
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const FinScratch s = carve_scratch(p->scratch, p->n_clips, p->n_frames);
    rms_max_kernel<NF_THREADS><<<p->n_clips, NF_THREADS, 0, st>>>(p->rms, p->rms_clip_stride, p->n_frames, s.rms_max);
    if (int rc = check_launch("aegis_fin_prepare(rms max)")) return rc;
    fin_prepare_kernel<<<frame_blocks(p), NF_THREADS, 0, st>>>(*p, s.rms_max, s.rms_db);
    return check_launch("aegis_fin_prepare(frames)");
}

extern "C" int aegis_fin_events(const aegis_fin_params* p, void* stream) {
    using namespace aegis;
    if (int rc = fin_check_common(p, "aegis_fin_events")) return rc;
    AEGIS_REQUIRE(p->trend && p->boll_upper && p->boll_lower && p->macd_line && p->macd_hist,
                  "aegis_fin_events: trend / bands / MACD series missing (run aegis_trend_filters after aegis_fin_prepare)");
    AEGIS_REQUIRE(p->events && p->n_events && p->max_events >= 0, "aegis_fin_events: outputs missing");
    AEGIS_REQUIRE(p->min_note_frames >= 0 && p->sustain_frames >= 0, "aegis_fin_events: negative frame counts");
    if (p->n_clips == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const FinScratch s = carve_scratch(p->scratch, p->n_clips, p->n_frames);
    if (p->n_frames > 0) {
        fin_frames_kernel<<<frame_blocks(p), NF_THREADS, 0, st>>>(*p, s.rms_db, s.frames);
        if (int rc = check_launch("aegis_fin_events(frames)")) return rc;
    }
    // AEGIS_FIN_EXACT_RSI=1 walks every density series step by step (the parity tests compare the two)
    const char* env = getenv("AEGIS_FIN_EXACT_RSI");
    const int exact_rsi = (env != nullptr && env[0] == '1') ? 1 : 0;
    if (p->n_frames <= FE_STAGE_MAX_FRAMES) {
        const size_t smem = static_cast<size_t>(p->n_frames) * 24 + 16;
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(fin_events_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            AEGIS_REQUIRE(e == cudaSuccess, "aegis_fin_events: %s", cudaGetErrorString(e));
        }
        fin_events_kernel<true><<<p->n_clips, 32, smem, st>>>(*p, s.frames, s.compact, exact_rsi);
    } else {
        fin_events_kernel<false><<<p->n_clips, 32, 0, st>>>(*p, s.frames, s.compact, exact_rsi);
    }
    return check_launch("aegis_fin_events(events)");
}

extern "C" long long aegis_fin_scratch_bytes(int n_clips, int n_frames) {
    const long long frames = static_cast<long long>(n_clips) * n_frames;
    return frames * static_cast<long long>(sizeof(aegis::FinFrame)) + frames * 8 + aegis::align16(frames * 4) +
           aegis::align16(static_cast<long long>(n_clips) * 4) + 16;
}
