// K2: batched probabilistic-YIN front half (librosa.pyin stages 1-9).
//
// Replaces the per-frame work of librosa.pyin reached from aegis_engine.py:63,67,190,216,
// aegis_engine_core/worker.py:9-15 and aegis_engine_financial.py:63-69:
//   difference function via FFT autocorrelation -> cumulative-mean normalisation -> troughs ->
//   beta(2,18)-weighted thresholds with a Boltzmann prior over trough rank -> parabolic period
//   refinement -> 10-cent pitch bins.  Output is the sparse column of the HMM observation matrix.
//
// hop == 512 (the reference's only hop, aegis_engine.py:17): yin_direct_kernel, no FFT at all.  The autocorrelation
//   acf_t[tau] = sum_{j=1..1024} y_t[j] * y_t[j + tau]
// of frame t is a sum of products over a 1024-sample window that slides by 512 samples per frame, so it is the sum of two
// BLOCK SUMS of 512 products each, and every block sum is shared by two consecutive frames: half the multiplies of a
// per-frame evaluation, only for the lags pYIN reads (tau <= max_period: 268 at 22.05 kHz, 536 at 44.1 kHz), in plain FP32
// FMAs with no twiddles, no exchanges and no three-pass code that overflows the instruction cache.  A CTA of 8 warps
// takes 15 consecutive frames of a clip = 16 blocks = 8 block pairs, one pair per warp: a lane owns 16 consecutive
// samples of each block (registers), slides a 16-sample window over them lag by lag (one shared-memory load per lag
// and block, conflict-free in a 17-word-per-16-samples skewed layout) and accumulates 16 lags at a time with packed
// f32x2 FMAs (lo = first block, hi = second block: the two never mix); a butterfly transpose-reduce sums the 32 lanes.
// Per block sum: a 16-term FMA chain per lane, then a binary tree over the lanes -- measured ~4x closer to the float64
// difference function than librosa's own float32 FFT evaluation (it is not bit-equal to it, and never was: pocketfft's
// rounding cannot be restated).  Block sums depend on nothing but their own samples, so a frame's result is the same
// whatever clip, batch or window it is analysed in.
// Energies / cumulative means are prefix scans in double, one warp per frame; the trough / threshold / prior stage runs
// one warp per frame with ballot compaction and shuffle reductions.
//
// Other hops (4..508): yin_fft_kernel, the round-1 path.  A CTA of 128 threads handles two consecutive frames
// of one clip (a "pair"), built on the same
// 2048-point FFT passes as the STFT kernel:
//   forward FFT of (frame + i*reversed-first-half) gives both spectra librosa multiplies;
//   the two frames' product spectra are packed as P1 + i*P2 and inverted with ONE transform
//   (acf of frame 1 in the real part, of frame 2 in the imaginary part).
// Energy / cumulative sums are prefix scans in double; the trough/threshold/prior stage runs one
// warp per frame with ballot compaction and shuffle reductions.
#include <cfloat>
#include <cstddef>
#include "common.cuh"
#include "fft2048.cuh"

namespace aegis {

constexpr int YIN_THREADS = 128;
constexpr int YIN_MAX_HOP = 512;
constexpr int YIN_MAX_LAGS = 1024;     // lags 0 .. max_period, max_period <= 1023
constexpr int YIN_MAX_TROUGHS = 512;

struct YinSmem {
    float en_part[2][4];  // per-warp partial frame energies (packing scale)
    float samples[FFT_N + YIN_MAX_HOP];
    cf buf[BUFA_SIZE];    // the transforms' single exchange buffer (in place); buf + P1 are reused as YinScratchE
    cf P1[FFT_N / 2 + 4]; // once the transforms are done (must follow buf directly)
    cf Z[FFT_N];          // reused as YinScratchC in the candidate phase
};
struct YinScratchE {       // lives in buf..P1 (25 632 B)
    double yin[2][YIN_MAX_LAGS];
    float d[2][YIN_MAX_LAGS];
    double tot[2][64];
    double e0[2][2];
};
struct YinScratchC {       // lives in Z (16 384 B)
    double tp[2][YIN_MAX_TROUGHS];
    unsigned short tk[2][YIN_MAX_TROUGHS];
    unsigned char tq[2][YIN_MAX_TROUGHS];
};
static_assert(sizeof(YinScratchE) <= sizeof(cf) * (BUFA_SIZE + FFT_N / 2 + 4), "scratch E too large");
static_assert(offsetof(YinSmem, P1) == offsetof(YinSmem, buf) + sizeof(cf) * BUFA_SIZE, "P1 must follow buf");
static_assert(sizeof(YinScratchC) <= sizeof(cf) * FFT_N, "scratch C too large");

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// passes 2 and 3 of a transform whose pass 1 has been written to buf, in place (one exchange buffer: 17 KB less
// shared memory than ping-ponging two, which is what lets a fourth CTA fit on the SM); natural-order result in out
__device__ __forceinline__ void fft_finish_inplace(int lt, cf* v, const FftTwiddles& tw, cf* buf, cf* out) {
    __syncthreads();
    fft2048_pass2_load(lt, buf, v);
    __syncthreads();
    fft2048_pass2_store(lt, v, tw, buf);
    __syncthreads();
    fft2048_pass3_load(lt, buf, v);
    fft2048_pass3_store(lt, v, out);   // out != buf: no barrier needed between the two halves
    __syncthreads();
}

// Candidate stage of one frame by one warp (librosa __pyin_helper): troughs of the CMND curve
// yv[0..L), their probabilities over the beta-weighted thresholds with the Boltzmann prior on the trough rank, parabolic
// refinement, 10-cent bins; writes the frame's sparse observation (bins ascending, unique) and its voiced probability.
// CMND value d[tau] / (mean of d[1..tau] + tiny) from the float32-rounded cumulative sum s: the mean's division is folded
// into the numerator, d tau / (s + tau tiny) -- one double division per lag instead of two (tau tiny only matters when
// s == 0, where it reproduces d / tiny; for any s != 0 the sum is s itself).  Within one ulp of the two-division form.
__device__ __forceinline__ double cmnd_value(float d, double s_rounded, int tau) {
    const double tt = static_cast<double>(tau);
    return (static_cast<double>(d) * tt) / fma(tt, DBL_MIN, s_rounded);
}

struct YinTables {   // the prior tables: global pointers (read through the read-only path) or copies in shared memory (beta_cumsum: always global)
    const double* thresholds;
    const double* beta_probs;
    const double* beta_cumsum;
    const double* boltz_fact;
    const double* boltz_exp;
};
template <bool SHARED>
__device__ __forceinline__ double tab(const double* q) { return SHARED ? *q : __ldg(q); }

template <bool SHARED>
__device__ __forceinline__ void yin_candidates_of_frame(const aegis_yin_params& p, const YinTables& tb, const double* yv, const int L, const int minp,
                                            unsigned short* tk, unsigned char* tq, double* tp, const int max_troughs,
                                            const long long fidx, const int lane) {
    // 1. troughs, compacted in lag order
    int nt = 0;
    for (int base = 0; base < L; base += 32) {
        const int k = base + lane;
        bool tr = false;
        if (k < L && L >= 2) {
            const double x = yv[k];
            if (k == 0) tr = x < yv[1];
            else if (k == L - 1) tr = x < yv[k - 1];
            else tr = (x < yv[k - 1]) && (x <= yv[k + 1]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, tr);
        if (tr) {
            const int pos = nt + __popc(m & ((1u << lane) - 1u));
            if (pos < max_troughs) tk[pos] = static_cast<unsigned short>(k);
        }
        nt += __popc(m);
    }
    nt = min(nt, max_troughs);
    __syncwarp();
    int count = 0;
    double vsum = 0.0;
    if (nt > 0) {
        // 2. first threshold index each trough is below; global minimum (first on ties)
        const int nth = p.n_thresholds;
        double best_h = DBL_MAX;
        int best_i = 0x7fffffff;
        for (int i = lane; i < nt; i += 32) {
            const double h = yv[tk[i]];
            int q = (h >= 1.0) ? nth : ((h <= 0.0) ? 0 : static_cast<int>(h * nth));  // guess, then fix
            q = max(0, min(q, nth));
            while (q > 0 && h < tab<SHARED>(tb.thresholds + q - 1)) --q;
            while (q < nth && !(h < tab<SHARED>(tb.thresholds + q))) ++q;
            tq[i] = static_cast<unsigned char>(q);
            tp[i] = 0.0;
            if (h < best_h) { best_h = h; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oh = __shfl_xor_sync(0xffffffffu, best_h, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oh < best_h || (oh == best_h && oi < best_i)) { best_h = oh; best_i = oi; }
        }
        __syncwarp();
        // 3. troughs below each threshold (this lane owns thresholds lane, lane+32, ...)
        int nj[4] = {0, 0, 0, 0}, pj[4] = {0, 0, 0, 0};
        for (int i = 0; i < nt; ++i) {
            const int q = tq[i];
#pragma unroll
            for (int r = 0; r < 4; ++r) nj[r] += (q <= lane + 32 * r) ? 1 : 0;
        }
        double fj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = lane + 32 * r;
            fj[r] = (j < nth && nj[r] > 0) ? tab<SHARED>(tb.boltz_fact + nj[r]) : 0.0;
        }
        // 4. probability of each trough: sum_j prior(rank among troughs below th_j) * beta_j
        for (int i = 0; i < nt; ++i) {
            const int q = tq[i];
            if (q >= nth) continue;  // warp-uniform
            double c = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = lane + 32 * r;
                if (j < nth && q <= j) {
                    c += (fj[r] * tab<SHARED>(tb.boltz_exp + pj[r])) * tab<SHARED>(tb.beta_probs + j);
                    ++pj[r];
                }
            }
            c = warp_sum_d(c);
            if (lane == 0) tp[i] = c;
        }
        __syncwarp();
        // 5. every lane turns its troughs' lags into pitch bins (parabolic refinement, log2: the expensive
        //    part, in parallel; tk[] is overwritten with the bin, 0xffff = not a candidate) ...
        if (lane == 0) tp[best_i] += p.no_trough_prob * __ldg(tb.beta_cumsum + tq[best_i]);   // one read per frame: stays global
        __syncwarp();
        {
            const double scale = 12.0 * p.bins_per_semitone;
            for (int i = lane; i < nt; i += 32) {
                unsigned short bin16 = 0xffffu;
                if (tp[i] != 0.0) {
                    const int k = tk[i];
                    double shift = 0.0;
                    if (k > 0 && k < L - 1) {
                        const double a = yv[k + 1] + yv[k - 1] - 2.0 * yv[k];
                        const double b = (yv[k + 1] - yv[k - 1]) / 2.0;
                        if (!(fabs(b) >= fabs(a))) shift = -b / a;
                    }
                    const double period = static_cast<double>(minp + k) + shift;
                    const double f0 = p.sr / period;
                    double bf = rint(scale * log2(f0 / p.fmin));
                    bf = fmin(fmax(bf, 0.0), static_cast<double>(p.n_pitch_bins));
                    const int bin = static_cast<int>(bf);
                    if (bin < p.n_pitch_bins) bin16 = static_cast<unsigned short>(bin);   // else: unvoiced rows, overwritten
                }
                tk[i] = bin16;
            }
        }
        __syncwarp();
        //    ... and they are emitted in ascending-bin order (= descending lag; bins never increase with the lag, so equal
        //    bins are neighbours among the candidates): the larger lag wins.  32 troughs per round, lane l takes trough
        //    hi - l; a candidate is kept unless the previous candidate in emission order has the same bin.
        {
            unsigned short* ob = p.cand_bin + fidx * p.max_cand;
            double* op = p.cand_prob + fidx * p.max_cand;
            int last_bin = -1;
            bool over = false;
            for (int hi = nt - 1; hi >= 0; hi -= 32) {
                const int i = hi - lane;
                const int bin = i >= 0 ? tk[i] : 0xffff;
                const bool valid = bin != 0xffff;
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                const unsigned below = vm & ((1u << lane) - 1u);
                const int prev = __shfl_sync(0xffffffffu, bin, below ? 31 - __clz(below) : 0);
                const bool keep = valid && bin != (below ? prev : last_bin);
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                const int pos = count + __popc(km & ((1u << lane) - 1u));
                if (keep) {
                    if (pos < p.max_cand) {
                        ob[pos] = static_cast<unsigned short>(bin);
                        op[pos] = tp[i];
                    } else {
                        over = true;
                    }
                }
                count += __popc(km);
                if (vm) last_bin = __shfl_sync(0xffffffffu, bin, 31 - __clz(vm));
            }
            if (__any_sync(0xffffffffu, over)) {
                if (lane == 0) atomicExch(p.overflow, 1);
                count = p.max_cand;
            }
            __syncwarp();
            if (lane == 0) {
                for (int c = 0; c < count; ++c) vsum += op[c];   // in emission order, as the oracle's row sum
            }
        }
    }
    if (lane == 0) {
        p.cand_count[fidx] = count;
        p.voiced_prob[fidx] = fmin(fmax(vsum, 0.0), 1.0);
    }
}

__global__ void __launch_bounds__(YIN_THREADS, 4)
yin_fft_kernel(const aegis_yin_params p, const int pairs_per_clip, const long long n_pairs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    YinSmem& s = *reinterpret_cast<YinSmem*>(smem_raw);
    YinScratchE& se = *reinterpret_cast<YinScratchE*>(s.buf);
    YinScratchC& sc = *reinterpret_cast<YinScratchC*>(s.Z);
    const int lt = threadIdx.x, lane = lt & 31, warp = lt >> 5;
    const int T = p.n_frames, hop = p.hop;
    const long long N = p.n_samples;
    const int maxp = p.max_period, minp = p.min_period;
    const int L = maxp - minp + 1;  // CMND length
    FftTwiddles tw;
    fft2048_load_twiddles(lt, reinterpret_cast<const cf*>(p.twiddle), tw);

    for (long long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int clip = static_cast<int>(pair / pairs_per_clip);
        const int t0 = static_cast<int>(pair - static_cast<long long>(clip) * pairs_per_clip) * 2;
        const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
        const long long g0 = static_cast<long long>(t0) * hop - p.pad;
        __syncthreads();
        #pragma unroll 4
        for (int i = lt; i < FFT_N + hop; i += YIN_THREADS) {
            const long long gi = g0 + i;
            s.samples[i] = (gi >= 0 && gi < N) ? __ldg(yc + gi) : 0.f;
        }
        __syncthreads();

        // The two frames' product spectra share the inverse transform, so (as in the STFT kernel) each
        // frame is scaled by an exact power of two ~ 1/||frame|| first: a quiet frame next to a loud one
        // keeps an autocorrelation error relative to its OWN energy.  acf = acf_scaled * 4^e.
#pragma unroll
        for (int fr = 0; fr < 2; ++fr) {
            const float* f = s.samples + fr * hop;
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 16; ++a) acc = fmaf(f[lt + 128 * a], f[lt + 128 * a], acc);
            acc = warp_sum(acc);
            if (lane == 0) s.en_part[fr][warp] = acc;
        }
        __syncthreads();
        int fexp[2];
        bool fzero[2];
#pragma unroll
        for (int fr = 0; fr < 2; ++fr) {
            const float en = (s.en_part[fr][0] + s.en_part[fr][1]) + (s.en_part[fr][2] + s.en_part[fr][3]);
            fexp[fr] = (en > 0.f) ? (ilogbf(en) >> 1) : 0;
            fzero[fr] = !(en > 0.f);  // digital silence: acf must be exactly 0, not the partner's noise
        }

        // ---- three transforms through ONE copy of the FFT code (the kernel is far larger than the instruction cache;
        // a second inlined copy for the inverse transform cost more in instruction fetch than the branches here):
        // it = 0, 1: spectra Z = FFT(frame + i*rev), rev[n] = frame[1024-n] (n < 1024), 0 otherwise, then the products;
        // it = 2:    one transform inverts both products: q[n] = conj(FFT(conj Q)[n]) / N = acf1[n] + i*acf2[n]
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
            {
                cf v[16];
                if (it < 2) {
                    const float* f = s.samples + it * hop;
                    const float sc = ldexpf(1.0f, -fexp[it]);
#pragma unroll
                    for (int a = 0; a < 16; ++a) {
                        const int n = lt + 128 * a;
                        v[a] = cf{f[n] * sc, (a < 8) ? f[FFT_N / 2 - n] * sc : 0.f};
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < 16; ++a) v[a] = s.Z[lt + 128 * a];
                }
                fft2048_pass1(lt, v, tw, s.buf);
                fft_finish_inplace(lt, v, tw, s.buf, s.Z);
            }
            if (it == 2) break;
            // P = (2A)(2B) with 2A = Z[k] + conj Z[N-k], 2B = (Z[k] - conj Z[N-k]) / i
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                const int k = lt + 128 * m;
                if (k <= FFT_N / 2) {
                    const int kn = (FFT_N - k) & (FFT_N - 1);
                    const cf zk = s.Z[k], zn = s.Z[kn];
                    const cf A2 = cf{zk.x + zn.x, zk.y - zn.y};
                    const cf B2 = cf{zk.y + zn.y, zn.x - zk.x};
                    const cf P = cmul(A2, B2);
                    if (it == 0) {
                        s.P1[k] = P;
                    } else {  // conj(Q), Q = P1 + i*P2 extended Hermitian-wise; in place over Z
                        const cf p1 = s.P1[k];
                        s.Z[k] = cf{p1.x - P.y, -(p1.y + P.x)};
                        if (kn != k) s.Z[kn] = cf{p1.x + P.y, p1.y - P.x};
                    }
                }
            }
            __syncthreads();
        }

        // ---- phase E: energies, difference function, cumulative mean, CMND (64 threads / frame)
        {
            const int fr = lt >> 6, u = lt & 63;
            const float* f = s.samples + fr * hop;
            double acc = 0.0;
            #pragma unroll 2
            for (int j = 1 + u; j <= FFT_N / 2; j += 64) acc += static_cast<double>(f[j] * f[j]);
            acc = warp_sum_d(acc);
            if (lane == 0) se.e0[fr][warp & 1] = acc;
            const int chunk = (maxp + 63) / 64;
            const int lo = 1 + u * chunk, hi = min(lo + chunk, maxp + 1);
            double run = 0.0;
            #pragma unroll 1
            for (int tau = lo; tau < hi; ++tau)
                run += static_cast<double>(f[FFT_N / 2 + tau] * f[FFT_N / 2 + tau]) - static_cast<double>(f[tau] * f[tau]);
            se.tot[fr][u] = run;
            __syncthreads();
            const double e0 = se.e0[fr][0] + se.e0[fr][1];
            double base = e0;
            #pragma unroll 2
            for (int v = 0; v < u; ++v) base += se.tot[fr][v];
            float e0f = static_cast<float>(e0);
            if (fabsf(e0f) < 1e-6f) e0f = 0.f;
            const float acf_scale = ldexpf(1.0f / (4.0f * FFT_N), 2 * fexp[fr]);
            run = 0.0;
            double dsum = 0.0;
            #pragma unroll 1
            for (int tau = lo; tau < hi; ++tau) {
                run += static_cast<double>(f[FFT_N / 2 + tau] * f[FFT_N / 2 + tau]) - static_cast<double>(f[tau] * f[tau]);
                float e = static_cast<float>(base + run);
                if (fabsf(e) < 1e-6f) e = 0.f;
                const cf F = s.Z[FFT_N / 2 + tau];
                float acf = (fr == 0 ? F.x : -F.y) * acf_scale;
                if (fabsf(acf) < 1e-6f || fzero[fr]) acf = 0.f;
                const float dv = (e0f + e) - 2.0f * acf;
                se.d[fr][tau] = dv;
                dsum += static_cast<double>(dv);
            }
            __syncthreads();  // tot[] fully consumed
            se.tot[fr][u] = dsum;
            __syncthreads();
            base = 0.0;
            #pragma unroll 2
            for (int v = 0; v < u; ++v) base += se.tot[fr][v];
            run = 0.0;
            #pragma unroll 1
            for (int tau = lo; tau < hi; ++tau) {
                run += static_cast<double>(se.d[fr][tau]);
                if (tau >= minp) {
                    const double yv = cmnd_value(se.d[fr][tau], static_cast<double>(static_cast<float>(base + run)), tau);
                    se.yin[fr][tau - minp] = yv;
                    if (p.cmnd_out != nullptr && t0 + fr < T)
                        p.cmnd_out[(static_cast<long long>(clip) * T + t0 + fr) * L + (tau - minp)] = yv;
                }
            }
        }
        __syncthreads();

        // ---- phase C: one warp per frame (warps 0 and 2)
        if ((warp & 1) == 0) {
            const int fr = warp >> 1;
            const int t = t0 + fr;
            if (t < T) {
                const YinTables tb{p.thresholds, p.beta_probs, p.beta_cumsum, p.boltz_fact, p.boltz_exp};
                yin_candidates_of_frame<false>(p, tb, se.yin[fr], L, minp, sc.tk[fr], sc.tq[fr], sc.tp[fr], YIN_MAX_TROUGHS,
                                        static_cast<long long>(clip) * T + t, lane);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// hop == 512: direct block-sum autocorrelation
// ------------------------------------------------------------------------------------------------------------------
constexpr int YD_THREADS = 256;                    // 8 warps = 8 block pairs
constexpr int YD_BLOCKS = 16;                      // blocks of 512 products per CTA
constexpr int YD_FRAMES = YD_BLOCKS - 1;           // frame i = block i + block i+1
// samples a run of 16 blocks reads with n_passes lag passes: m < 16 * 512 + 1 + 272 n_passes + 64 (the sliding windows look ahead)

struct YinDirectLayout {      // byte offsets into dynamic shared memory (computed on the host from max_period)
    int n_passes;             // lag passes of 272
    int span;                 // samples staged per run
    int b_pitch;              // floats per block row of B (272 * n_passes; lag tau at blocksum_index(tau))
    int off_b, off_warp;      // B sums; first per-warp scratch area
    int warp_bytes;           // per-warp scratch: yin (double), d (float), tp (double), tk (u16), tq (u8)
    int off_d, off_tp, off_tk, off_tq;   // inside a warp's area
    int max_troughs;
    int total_bytes;
};

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// inclusive warp scan of doubles (Hillis-Steele)
__device__ __forceinline__ double warp_scan_incl_d(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// Lane map of the block-sum kernel: lane = (lag group g = lane & 15, sample half h = lane >> 4).  A lane owns YB_WG = 17
// consecutive lags (16 groups x 17 = 272 lags per pass; E2 at 22.05 kHz needs 269, at 44.1 kHz 537 = two passes) and walks
// its half of the block's 512 samples with a 17-entry sliding window in registers: per sample one broadcast load of
// x[n], one load of the sample entering the window, 17 packed FMAs (the two halves of a register pair are the two blocks
// of the warp's block pair).  Nothing is reduced across lanes but the two halves (one shuffle per lag).  The halves are
// 272 / 240 samples, not 256 / 256: 272 = 16 * 17 is a whole number of window rotations and puts the upper half's
// loads 16 banks from the lower half's (17 g mod 32 covers 16 banks), so neither the broadcast nor the window loads
// conflict and the samples need no skewed layout.  (The first cut owned 16 samples x 16 lags per lane and reduced over
// all 32 lanes: 39 % of its instructions were the transpose-reduce and the pipe sat at 60 %.)
constexpr int YB_WG = 17;
constexpr int YB_PASS_LAGS = 16 * YB_WG;     // 272
constexpr int YB_LOWER = 16 * YB_WG;         // 272 samples for h = 0, 240 for h = 1
constexpr int YB_MAX_PASSES = 4;             // lags <= 1087 >= FFT_N / 2 - 1
constexpr int YS_SPAN_MAX = YD_BLOCKS * 512 + 1 + YB_MAX_PASSES * YB_PASS_LAGS + 64;
constexpr int YS_PHYS = ((YS_SPAN_MAX + 3) / 4) * 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
// 4-byte cp.async with zero fill: copies src_bytes (4 or 0) bytes and zero-fills the rest of the 4
__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gmem_src, int src_bytes) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}

// position of lag tau in a block-sum row: pass-major, then lag-in-group, then group (what a warp stores side by side)
__host__ __device__ __forceinline__ int blocksum_index(int tau) {
    const int pass = tau / YB_PASS_LAGS, r = tau - pass * YB_PASS_LAGS, g = r / YB_WG;
    return pass * YB_PASS_LAGS + 16 * (r - g * YB_WG) + g;
}

// one rotation of the window = 17 samples.  MASKED: the upper half-warp has run out of its 240 samples at sample `n0 + s`
template <bool MASKED>
__device__ __forceinline__ void blocksum_rotation(const float* __restrict__ pa, const float* __restrict__ pr, unsigned long long (&c)[YB_WG],
                                                  unsigned long long (&R)[YB_WG], const int n0, const bool upper) {
#pragma unroll
    for (int s_ = 0; s_ < YB_WG; ++s_) {
        float a0 = pa[s_], a1 = pa[512 + s_];
        if (MASKED) {
            const bool off = upper && (n0 + s_ >= 512 - YB_LOWER);
            a0 = off ? 0.f : a0;
            a1 = off ? 0.f : a1;
        }
        const unsigned long long a = pack2(a0, a1);
#pragma unroll
        for (int j = 0; j < YB_WG; ++j) c[j] = ffma2(a, R[(s_ + j) % YB_WG], c[j]);
        R[s_] = pack2(pr[YB_WG + s_], pr[512 + YB_WG + s_]);
    }
}

__global__ void __launch_bounds__(YD_THREADS, 2)
yin_direct_kernel(const aegis_yin_params p, const YinDirectLayout lay, const int runs_per_clip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);
    float* Bs = reinterpret_cast<float*>(smem_raw + lay.off_b);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = blockIdx.x / runs_per_clip;
    const int t0 = (blockIdx.x - clip * runs_per_clip) * YD_FRAMES;
    const int T = p.n_frames;
    const long long N = p.n_samples;
    const int maxp = p.max_period, minp = p.min_period;
    const int L = maxp - minp + 1;
    const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
    const long long g0 = static_cast<long long>(t0) * 512 - p.pad;   // clip sample of m = 0
    const int n_fr = min(YD_FRAMES, T - t0);                         // frames of this run
    const int m_end = 512 * (n_fr - 1) + 2048;                       // samples the frames of this run span

    // ---- phase A: samples -> shared (zeros outside the clip: center padding; zeros past the run's last frame)
    for (int m = tid; m < lay.span; m += YD_THREADS) {
        const long long gi = g0 + m;
        xs[m] = (m < m_end && gi >= 0 && gi < N) ? __ldg(yc + gi) : 0.f;
    }
    __syncthreads();

    // ---- phase B: block sums B[b][tau] = sum_{n in block b} x[n] x[n + tau], block b = samples m = 512 b + 1 .. 512 b + 512,
    // summed exactly as yin_blocksum_kernel sums them (same lane map, same order: the two forms of K2 give identical bits).
    // Warp w: blocks 2w (low words) and 2w+1 (high words)
    if (2 * warp < n_fr + 1) {
        const int g = lane & 15;
        const bool upper = (lane & 16) != 0;
        float* brow = Bs + (2 * warp + (upper ? 1 : 0)) * lay.b_pitch + g;
        const float* base = xs + 512 * (2 * warp) + 1 + (upper ? YB_LOWER : 0);
#pragma unroll 1
        for (int pass = 0; pass < lay.n_passes; ++pass) {
            const int lag0 = pass * YB_PASS_LAGS + YB_WG * g;
            const float* pa = base;
            const float* pr = base + lag0;
            unsigned long long c[YB_WG], R[YB_WG];
#pragma unroll
            for (int j = 0; j < YB_WG; ++j) {
                c[j] = 0ull;
                R[j] = pack2(pr[j], pr[512 + j]);
            }
            constexpr int FULL = (512 - YB_LOWER) / YB_WG;
#pragma unroll 1
            for (int it = 0; it < FULL; ++it, pa += YB_WG, pr += YB_WG) blocksum_rotation<false>(pa, pr, c, R, YB_WG * it, upper);
#pragma unroll 1
            for (int it = FULL; it < YB_LOWER / YB_WG; ++it, pa += YB_WG, pr += YB_WG) blocksum_rotation<true>(pa, pr, c, R, YB_WG * it, upper);
#pragma unroll
            for (int j = 0; j < YB_WG; ++j) {
                float lo, hi;
                unpack2(c[j], lo, hi);
                const float keep = upper ? hi : lo, send = upper ? lo : hi;
                brow[pass * YB_PASS_LAGS + 16 * j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
    }
    __syncthreads();

    // ---- phases E + C, one warp per frame: energies, difference function, cumulative mean, CMND; then the candidates
    unsigned char* wbase = smem_raw + lay.off_warp + warp * lay.warp_bytes;
    double* yin = reinterpret_cast<double*>(wbase);
    float* dbuf = reinterpret_cast<float*>(wbase + lay.off_d);
    double* tp = reinterpret_cast<double*>(wbase + lay.off_tp);
    unsigned short* tk = reinterpret_cast<unsigned short*>(wbase + lay.off_tk);
    unsigned char* tq = wbase + lay.off_tq;
    for (int fr = warp; fr < n_fr; fr += YD_THREADS / 32) {
        const int t = t0 + fr;
        const float* f = xs;                       // frame sample j = x[m = 512 fr + j]
        auto fs = [&](int j) -> float { return f[512 * fr + j]; };
        double acc = 0.0;
#pragma unroll 4
        for (int j = 1 + lane; j <= FFT_N / 2; j += 32) { const float v = fs(j); acc += static_cast<double>(v * v); }
        const double e0 = warp_sum_d(acc);
        const int chunk = (maxp + 31) / 32;
        const int lo = 1 + lane * chunk, hi = min(lo + chunk, maxp + 1);
        double run = 0.0;
#pragma unroll 1
        for (int tau = lo; tau < hi; ++tau) {
            const float u1 = fs(FFT_N / 2 + tau), u0 = fs(tau);
            run += static_cast<double>(u1 * u1) - static_cast<double>(u0 * u0);
        }
        double incl = warp_scan_incl_d(run, lane);
        double base = e0 + (incl - run);           // e0 + sum of the lower lanes' chunks
        float e0f = static_cast<float>(e0);
        if (fabsf(e0f) < 1e-6f) e0f = 0.f;
        const float* B0 = Bs + fr * lay.b_pitch;
        const float* B1 = B0 + lay.b_pitch;
        run = 0.0;
        double dsum = 0.0;
        int bj = lo % YB_PASS_LAGS, bi = lo - bj;   // position of lag `lo` in a block-sum row (blocksum_index), stepped with tau
        { const int bg = bj / YB_WG; bj -= bg * YB_WG; bi += bg; }
#pragma unroll 1
        for (int tau = lo; tau < hi; ++tau) {
            const float u1 = fs(FFT_N / 2 + tau), u0 = fs(tau);
            run += static_cast<double>(u1 * u1) - static_cast<double>(u0 * u0);
            float e = static_cast<float>(base + run);
            if (fabsf(e) < 1e-6f) e = 0.f;
            float acf = B0[bi + 16 * bj] + B1[bi + 16 * bj];
            if (++bj == YB_WG) {
                bj = 0;
                bi += ((bi & 15) == 15) ? YB_PASS_LAGS - 15 : 1;
            }
            if (fabsf(acf) < 1e-6f) acf = 0.f;
            const float dv = (e0f + e) - 2.0f * acf;
            dbuf[tau] = dv;
            dsum += static_cast<double>(dv);
        }
        incl = warp_scan_incl_d(dsum, lane);
        base = incl - dsum;
        run = 0.0;
        __syncwarp();
#pragma unroll 1
        for (int tau = lo; tau < hi; ++tau) {
            run += static_cast<double>(dbuf[tau]);
            if (tau >= minp) {
                const double yv = cmnd_value(dbuf[tau], static_cast<double>(static_cast<float>(base + run)), tau);
                yin[tau - minp] = yv;
                if (p.cmnd_out != nullptr) p.cmnd_out[(static_cast<long long>(clip) * T + t) * L + (tau - minp)] = yv;
            }
        }
        __syncwarp();
        const YinTables tb{p.thresholds, p.beta_probs, p.beta_cumsum, p.boltz_fact, p.boltz_exp};
        yin_candidates_of_frame<false>(p, tb, yin, L, minp, tk, tq, tp, lay.max_troughs, static_cast<long long>(clip) * T + t, lane);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------------------------
// hop == 512, two kernels (used when the caller provides the block-sum workspace): the FMA-bound block sums at 128
// registers per thread, then everything per frame at high occupancy (the energy / CMND / candidate stage is a chain of
// short dependent steps: it wants warps, not registers).
// ------------------------------------------------------------------------------------------------------------------
// block sums of blocks [16 run, 16 run + 16) of a clip -> bsum[clip][block][b_pitch] (lag tau at blocksum_index(tau)).
// Persistent CTAs (two per SM) walk the runs; the samples of the NEXT run are requested with cp.async while this run's
// sums are being accumulated, so the FMA pipe does not idle behind the fill at the start of every run.
#ifndef AEGIS_YB_MINB
#define AEGIS_YB_MINB 2
#endif
__global__ void __launch_bounds__(YD_THREADS, AEGIS_YB_MINB)
yin_blocksum_kernel(const aegis_yin_params p, const int n_passes, const int b_pitch, const int runs_per_clip, const int n_blocks,
                    const long long n_runs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const bufs = reinterpret_cast<float*>(smem_raw);                  // two sample buffers of YS_PHYS floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = p.n_samples;
    const int span = YD_BLOCKS * 512 + 1 + n_passes * YB_PASS_LAGS + 64;
    const int ys_phys = ((span + 3) / 4) * 4;                                // floats per sample buffer

    auto request = [&](long long run, float* xs) {      // samples of `run` -> xs, zeros outside the clip
        const int clip = static_cast<int>(run / runs_per_clip);
        const int k0 = static_cast<int>(run - static_cast<long long>(clip) * runs_per_clip) * YD_BLOCKS;
        const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
        const long long g0 = static_cast<long long>(k0) * 512 - p.pad;       // clip sample of m = 0; block b = m in [512 b + 1, 512 b + 512]
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(yc) & 15) == 0) && ((g0 & 3) == 0);
        for (int m = tid * 4; m < span; m += 4 * YD_THREADS) {
            const long long gi = g0 + m;
            if (vec_ok && gi >= 0 && gi + 3 < N) {
                cp_async16(&xs[m], yc + gi);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool in = gi + j >= 0 && gi + j < N;
                    cp_async4_zfill(&xs[m + j], yc + (in ? gi + j : 0), in ? 4 : 0);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int g = lane & 15;
    const bool upper = (lane & 16) != 0;
    long long run = blockIdx.x;
    int cur = 0;
    if (run < n_runs) request(run, bufs);
    for (; run < n_runs; run += gridDim.x, cur ^= 1) {
        const long long next = run + gridDim.x;
        if (next < n_runs) {
            request(next, bufs + (cur ^ 1) * ys_phys);
            asm volatile("cp.async.wait_group 1;" ::: "memory");           // this run's samples have landed (the next run's may be in flight)
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* xs = bufs + cur * ys_phys;
        const int clip = static_cast<int>(run / runs_per_clip);
        const int k0 = static_cast<int>(run - static_cast<long long>(clip) * runs_per_clip) * YD_BLOCKS;
        if (k0 + 2 * warp < n_blocks) {
            const int kb = k0 + 2 * warp + (upper ? 1 : 0);                     // the block this lane stores
            float* brow = p.block_sums + (static_cast<long long>(clip) * n_blocks + min(kb, n_blocks - 1)) * b_pitch + g;
            const bool store = kb < n_blocks;
            const float* base = xs + 512 * (2 * warp) + 1 + (upper ? YB_LOWER : 0);   // first sample of this lane's half; block 2w+1 is 512 further
#pragma unroll 1
            for (int pass = 0; pass < n_passes; ++pass) {
                const int lag0 = pass * YB_PASS_LAGS + YB_WG * g;
                const float* pa = base;
                const float* pr = base + lag0;
                unsigned long long c[YB_WG], R[YB_WG];
#pragma unroll
                for (int j = 0; j < YB_WG; ++j) {
                    c[j] = 0ull;
                    R[j] = pack2(pr[j], pr[512 + j]);
                }
                constexpr int FULL = (512 - YB_LOWER) / YB_WG;                   // rotations both halves run unmasked (14)
#pragma unroll 1
                for (int it = 0; it < FULL; ++it, pa += YB_WG, pr += YB_WG) blocksum_rotation<false>(pa, pr, c, R, YB_WG * it, upper);
#pragma unroll 1
                for (int it = FULL; it < YB_LOWER / YB_WG; ++it, pa += YB_WG, pr += YB_WG) blocksum_rotation<true>(pa, pr, c, R, YB_WG * it, upper);
                // the other half's sum: the lower lanes keep block 2w (low words), the upper lanes block 2w+1 (high words)
#pragma unroll
                for (int j = 0; j < YB_WG; ++j) {
                    float lo, hi;
                    unpack2(c[j], lo, hi);
                    const float keep = upper ? hi : lo, send = upper ? lo : hi;
                    const float v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    if (store && lag0 + j <= p.max_period) brow[pass * YB_PASS_LAGS + 16 * j] = v;
                }
            }
        }
        __syncthreads();   // every warp is done with this buffer before the run after next is requested into it
    }
}

#ifndef AEGIS_YF_WARPS
#define AEGIS_YF_WARPS 7
#endif
constexpr int YF_WARPS = AEGIS_YF_WARPS;   // frames per CTA of the per-frame kernel (7: 53 KB of shared memory, four CTAs per SM; 8: 60 KB, three)

struct YinFrameLayout {
    int b_pitch, n_blocks;
    int span;                 // samples staged per CTA: 7 * 512 + 1025 + max_period
    int off_b, off_tab, off_warp, warp_bytes, off_d, off_tp, off_tk, off_tq, max_troughs, total_bytes;
};

__global__ void __launch_bounds__(32 * YF_WARPS)
yin_frame_kernel(const aegis_yin_params p, const YinFrameLayout lay, const int groups_per_clip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);
    float* Bs = reinterpret_cast<float*>(smem_raw + lay.off_b);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = blockIdx.x / groups_per_clip;
    const int t0 = (blockIdx.x - clip * groups_per_clip) * YF_WARPS;
    const int T = p.n_frames;
    const long long N = p.n_samples;
    const int maxp = p.max_period, minp = p.min_period;
    const int L = maxp - minp + 1;
    const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
    const long long g0 = static_cast<long long>(t0) * 512 - p.pad;
    const int n_fr = min(YF_WARPS, T - t0);
    const int m_end = 512 * (n_fr - 1) + 1025 + maxp;
    // everything the CTA reads is requested up front with asynchronous 16- / 8-byte copies (one round trip; the
    // register-staged loops this replaces took five, 21 % of the kernel's stall samples)
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(yc) & 15) == 0) && ((g0 & 3) == 0);
    for (int m = tid * 4; m < m_end; m += 4 * 32 * YF_WARPS) {
        const long long gi = g0 + m;
        if (vec_ok && gi >= 0 && gi + 3 < N) {
            cp_async16(&xs[m], yc + gi);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) xs[m + j] = (gi + j >= 0 && gi + j < N) ? __ldg(yc + gi + j) : 0.f;
        }
    }
    {   // block sums of blocks t0 .. t0 + n_fr (rows of b_pitch = 16 k floats: 16-byte pieces)
        const float* src = p.block_sums + (static_cast<long long>(clip) * lay.n_blocks + t0) * lay.b_pitch;
        const int cnt = (n_fr + 1) * lay.b_pitch;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = tid * 4; i < cnt; i += 4 * 32 * YF_WARPS) cp_async16(&Bs[i], src + i);
        } else {
            for (int i = tid; i < cnt; i += 32 * YF_WARPS) Bs[i] = __ldg(src + i);
        }
    }
    // the prior tables, copied once per CTA: thresholds | beta_probs | boltz_fact | boltz_exp
    double* tabs = reinterpret_cast<double*>(smem_raw + lay.off_tab);
    const int nth = p.n_thresholds, nbz = lay.max_troughs + 1;
    for (int i = tid; i < nth; i += 32 * YF_WARPS) { cp_async8(&tabs[i], p.thresholds + i); cp_async8(&tabs[nth + i], p.beta_probs + i); }
    for (int i = tid; i < nbz; i += 32 * YF_WARPS) { cp_async8(&tabs[2 * nth + i], p.boltz_fact + i); cp_async8(&tabs[2 * nth + nbz + i], p.boltz_exp + i); }
    const YinTables tb{tabs, tabs + nth, p.beta_cumsum, tabs + 2 * nth, tabs + 2 * nth + nbz};
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp >= n_fr) return;
    unsigned char* wbase = smem_raw + lay.off_warp + warp * lay.warp_bytes;
    double* yin = reinterpret_cast<double*>(wbase);
    float* dbuf = reinterpret_cast<float*>(wbase + lay.off_d);
    double* tp = reinterpret_cast<double*>(wbase + lay.off_tp);
    unsigned short* tk = reinterpret_cast<unsigned short*>(wbase + lay.off_tk);
    unsigned char* tq = wbase + lay.off_tq;
    const int fr = warp, t = t0 + fr;
    const float* f = xs + 512 * fr;                       // frame sample j = f[j]
    double acc = 0.0;
#pragma unroll 4
    for (int j = 1 + lane; j <= FFT_N / 2; j += 32) { const float v = f[j]; acc += static_cast<double>(v * v); }
    const double e0 = warp_sum_d(acc);
    const int chunk = (maxp + 31) / 32;
    const int lo = 1 + lane * chunk, hi = min(lo + chunk, maxp + 1);
    double run = 0.0;
#pragma unroll 1
    for (int tau = lo; tau < hi; ++tau) {
        const float u1 = f[FFT_N / 2 + tau], u0 = f[tau];
        run += static_cast<double>(u1 * u1) - static_cast<double>(u0 * u0);
    }
    double incl = warp_scan_incl_d(run, lane);
    double base = e0 + (incl - run);
    float e0f = static_cast<float>(e0);
    if (fabsf(e0f) < 1e-6f) e0f = 0.f;
    const float* B0 = Bs + fr * lay.b_pitch;
    const float* B1 = B0 + lay.b_pitch;
    run = 0.0;
    double dsum = 0.0;
    // position of lag `lo` in a block-sum row (blocksum_index), stepped along with tau
    int bj = lo % YB_PASS_LAGS, bi = lo - bj;
    { const int bg = bj / YB_WG; bj -= bg * YB_WG; bi += bg; }
#pragma unroll 1
    for (int tau = lo; tau < hi; ++tau) {
        const float u1 = f[FFT_N / 2 + tau], u0 = f[tau];
        run += static_cast<double>(u1 * u1) - static_cast<double>(u0 * u0);
        float e = static_cast<float>(base + run);
        if (fabsf(e) < 1e-6f) e = 0.f;
        float acf = B0[bi + 16 * bj] + B1[bi + 16 * bj];
        if (++bj == YB_WG) {                       // next group; after the 16th, the next pass
            bj = 0;
            bi += ((bi & 15) == 15) ? YB_PASS_LAGS - 15 : 1;
        }
        if (fabsf(acf) < 1e-6f) acf = 0.f;
        const float dv = (e0f + e) - 2.0f * acf;
        dbuf[tau] = dv;
        dsum += static_cast<double>(dv);
    }
    incl = warp_scan_incl_d(dsum, lane);
    base = incl - dsum;
    run = 0.0;
#pragma unroll 1
    for (int tau = lo; tau < hi; ++tau) {
        run += static_cast<double>(dbuf[tau]);
        if (tau >= minp) {
            const double yv = cmnd_value(dbuf[tau], static_cast<double>(static_cast<float>(base + run)), tau);
            yin[tau - minp] = yv;
            if (p.cmnd_out != nullptr) p.cmnd_out[(static_cast<long long>(clip) * T + t) * L + (tau - minp)] = yv;
        }
    }
    __syncwarp();
    yin_candidates_of_frame<true>(p, tb, yin, L, minp, tk, tq, tp, lay.max_troughs, static_cast<long long>(clip) * T + t, lane);
}

}  // namespace aegis

extern "C" int aegis_yin_candidates(const aegis_yin_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->y && p->twiddle, "aegis_yin_candidates: y / twiddle must be set");
    AEGIS_REQUIRE(p->hop >= 1 && p->hop <= YIN_MAX_HOP, "aegis_yin_candidates: hop=%d unsupported (1..512)", p->hop);
    AEGIS_REQUIRE(p->min_period >= 1 && p->max_period > p->min_period && p->max_period <= FFT_N / 2 - 1,
                  "aegis_yin_candidates: periods [%d, %d] out of range", p->min_period, p->max_period);
    AEGIS_REQUIRE(p->n_thresholds >= 1 && p->n_thresholds <= 128, "aegis_yin_candidates: n_thresholds must be 1..128");
    AEGIS_REQUIRE(p->thresholds && p->beta_probs && p->beta_cumsum && p->boltz_fact && p->boltz_exp,
                  "aegis_yin_candidates: prior tables missing");
    AEGIS_REQUIRE(p->cand_bin && p->cand_prob && p->cand_count && p->voiced_prob && p->overflow && p->max_cand >= 1,
                  "aegis_yin_candidates: outputs missing");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->hop == 512 && p->block_sums != nullptr) {   // the reference's hop, workspace given: block sums, then the per-frame stage
        const int L = p->max_period - p->min_period + 1;
        const int n_passes = (p->max_period + 1 + YB_PASS_LAGS - 1) / YB_PASS_LAGS;
        AEGIS_REQUIRE(n_passes <= YB_MAX_PASSES, "aegis_yin_candidates: max_period=%d needs more block-sum passes than compiled", p->max_period);
        const int b_pitch = n_passes * YB_PASS_LAGS;
        const int n_blocks = p->n_frames + 1;
        const int runs_per_clip = (n_blocks + YD_BLOCKS - 1) / YD_BLOCKS;
        const long long n_runs = static_cast<long long>(runs_per_clip) * p->n_clips;
        {
            const int span = YD_BLOCKS * 512 + 1 + n_passes * YB_PASS_LAGS + 64;
            const int smem = 2 * (((span + 3) / 4) * 4) * static_cast<int>(sizeof(float));
            cudaError_t e = cudaFuncSetAttribute(yin_blocksum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) {
                set_error("aegis_yin_candidates: cannot reserve %d B shared memory: %s", smem, cudaGetErrorString(e));
                return 2;
            }
            const long long max_grid = static_cast<long long>(AEGIS_YB_MINB) * sm_count();
            const unsigned grid = static_cast<unsigned>(n_runs < max_grid ? n_runs : max_grid);
            yin_blocksum_kernel<<<grid, YD_THREADS, smem, st>>>(*p, n_passes, b_pitch, runs_per_clip, n_blocks, n_runs);
        }
        if (int rc = check_launch("aegis_yin_candidates(block sums)")) return rc;
        YinFrameLayout lay{};
        lay.b_pitch = b_pitch;
        lay.n_blocks = n_blocks;
        lay.span = 512 * (YF_WARPS - 1) + 1025 + p->max_period;
        lay.off_b = ((lay.span * 4 + 15) / 16) * 16;
        lay.max_troughs = L / 2 + 1 < YIN_MAX_TROUGHS ? L / 2 + 1 : YIN_MAX_TROUGHS;
        lay.off_tab = lay.off_b + (((YF_WARPS + 1) * b_pitch * 4 + 15) / 16) * 16;
        lay.off_warp = lay.off_tab + (((2 * p->n_thresholds + 2 * (lay.max_troughs + 1)) * 8 + 15) / 16) * 16;
        // per warp: the CMND row, then the difference row d[] -- dead once the CMND is written -- sharing its bytes with the
        // trough lists (first written after that): 56 KB per CTA instead of 65, four CTAs per SM instead of three
        int o = ((L * 8 + 15) / 16) * 16;
        const int d_bytes = (((p->max_period + 1) * 4 + 15) / 16) * 16;
        lay.off_d = o;
        lay.off_tp = o;
        int q = o + ((lay.max_troughs * 8 + 15) / 16) * 16;
        lay.off_tk = q;      q += ((lay.max_troughs * 2 + 15) / 16) * 16;
        lay.off_tq = q;      q += ((lay.max_troughs + 15) / 16) * 16;
        o = (q - o > d_bytes) ? q : o + d_bytes;
        lay.warp_bytes = o;
        lay.total_bytes = lay.off_warp + YF_WARPS * lay.warp_bytes;
        AEGIS_REQUIRE(lay.total_bytes <= 227 * 1024, "aegis_yin_candidates: %d B shared memory needed for max_period=%d", lay.total_bytes, p->max_period);
        cudaError_t e = cudaFuncSetAttribute(yin_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total_bytes);
        if (e != cudaSuccess) {
            set_error("aegis_yin_candidates: cannot reserve %d B shared memory: %s", lay.total_bytes, cudaGetErrorString(e));
            return 2;
        }
        const int groups_per_clip = (p->n_frames + YF_WARPS - 1) / YF_WARPS;
        const long long n_groups_total = static_cast<long long>(groups_per_clip) * p->n_clips;
        AEGIS_REQUIRE(n_groups_total < (1ll << 31), "aegis_yin_candidates: too many frame groups for one launch");
        yin_frame_kernel<<<static_cast<unsigned>(n_groups_total), 32 * YF_WARPS, lay.total_bytes, st>>>(*p, lay, groups_per_clip);
        return check_launch("aegis_yin_candidates(frames)");
    }
    if (p->hop == 512) {   // no workspace: one kernel does both stages (15 frames per CTA)
        YinDirectLayout lay{};
        const int L = p->max_period - p->min_period + 1;
        lay.n_passes = (p->max_period + 1 + YB_PASS_LAGS - 1) / YB_PASS_LAGS;
        AEGIS_REQUIRE(lay.n_passes <= YB_MAX_PASSES, "aegis_yin_candidates: max_period=%d needs more block-sum passes than compiled", p->max_period);
        lay.span = YD_BLOCKS * 512 + 1 + lay.n_passes * YB_PASS_LAGS + 64;
        lay.b_pitch = lay.n_passes * YB_PASS_LAGS;
        lay.off_b = ((lay.span * 4 + 15) / 16) * 16;
        lay.off_warp = lay.off_b + ((YD_BLOCKS * lay.b_pitch * 4 + 15) / 16) * 16;
        lay.max_troughs = L / 2 + 1 < YIN_MAX_TROUGHS ? L / 2 + 1 : YIN_MAX_TROUGHS;
        int o = ((L * 8 + 15) / 16) * 16;
        lay.off_d = o;       o += (((p->max_period + 1) * 4 + 15) / 16) * 16;
        lay.off_tp = o;      o += ((lay.max_troughs * 8 + 15) / 16) * 16;
        lay.off_tk = o;      o += ((lay.max_troughs * 2 + 15) / 16) * 16;
        lay.off_tq = o;      o += ((lay.max_troughs + 15) / 16) * 16;
        lay.warp_bytes = o;
        lay.total_bytes = lay.off_warp + (YD_THREADS / 32) * lay.warp_bytes;
        AEGIS_REQUIRE(lay.total_bytes <= 227 * 1024, "aegis_yin_candidates: %d B shared memory needed for max_period=%d", lay.total_bytes, p->max_period);
        cudaError_t e = cudaFuncSetAttribute(yin_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total_bytes);
        if (e != cudaSuccess) {
            set_error("aegis_yin_candidates: cannot reserve %d B shared memory: %s", lay.total_bytes, cudaGetErrorString(e));
            return 2;
        }
        const int runs_per_clip = (p->n_frames + YD_FRAMES - 1) / YD_FRAMES;
        const long long n_runs = static_cast<long long>(runs_per_clip) * p->n_clips;
        AEGIS_REQUIRE(n_runs < (1ll << 31), "aegis_yin_candidates: too many frame runs for one launch");
        yin_direct_kernel<<<static_cast<unsigned>(n_runs), YD_THREADS, lay.total_bytes, st>>>(*p, lay, runs_per_clip);
        return check_launch("aegis_yin_candidates(direct)");
    }
    const int pairs_per_clip = (p->n_frames + 1) / 2;
    const long long n_pairs = static_cast<long long>(pairs_per_clip) * p->n_clips;
    cudaError_t e = cudaFuncSetAttribute(yin_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(YinSmem)));
    if (e != cudaSuccess) {
        set_error("aegis_yin_candidates: cannot reserve %zu B shared memory: %s", sizeof(YinSmem), cudaGetErrorString(e));
        return 2;
    }
    const long long max_grid = static_cast<long long>(sm_count()) * 4;
    const int grid = static_cast<int>(n_pairs < max_grid ? n_pairs : max_grid);
    yin_fft_kernel<<<grid, YIN_THREADS, sizeof(YinSmem), st>>>(*p, pairs_per_clip, n_pairs);
    return check_launch("aegis_yin_candidates");
}

// bytes of the block-sum workspace of aegis_yin_candidates (hop 512): (n_frames + 1) rows of 272 * ceil((max_period + 1) / 272) floats per clip
extern "C" long long aegis_yin_workspace_bytes(int n_clips, int n_frames, int max_period) {
    if (n_clips <= 0 || n_frames <= 0 || max_period < 0) return 0;
    const long long pitch = static_cast<long long>(aegis::YB_PASS_LAGS) * ((max_period + 1 + aegis::YB_PASS_LAGS - 1) / aegis::YB_PASS_LAGS);
    return static_cast<long long>(n_clips) * (n_frames + 1) * pitch * 4;
}
