// K4: mel power -> dB, rake-noise mask, spectral-flux onset envelope, onset peak picking.
//
// Replaces librosa.power_to_db(S, ref=np.max) (aegis_engine.py:26), detect_rake_patterns
// (aegis_engine_core/vision.py:3-38) and librosa.onset.onset_strength / util.peak_pick (no
// reference call site; named by BASELINE north_star, defined in SURVEY.md Appendix A.6).
//
// One thread per spectrogram column (coalesced across the time axis), CTAs of 256 columns of
// which the outer 32 on each side are halo (rake run-length gate looks at neighbours).  Every
// 10*log10 is evaluated once per (mel, column) and serves S_dB, the column test and the flux.
#include "common.cuh"

namespace aegis {

// 10 log10(x) for x >= 1e-10 through the hardware log2 (MUFU.LG2, ~2^-22 relative): the kernel evaluates one
// logarithm per spectrogram cell and was bound by the ~40-instruction accurate log10f (ncu: issue 80 %, 0.70 ms
// for 1024 x 30 s); the dB image is compared at 5e-3 and the same function serves the cell, the column maximum and
// the clip maximum, so max(S_dB) is still exactly 0.
__device__ __forceinline__ float db10(float x) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));   // arguments are >= 1e-10: no denormal handling needed
    // rounded product, never contracted into a following add: the cell, the column maximum, the clip maximum and the
    // floors derived from it must all see the SAME value of dB(x), in every kernel that uses this function
    return __fmul_rn(3.0102999566398120f, l);
}

constexpr int MP_THREADS = 256;
constexpr int MP_HALO = 32;
constexpr int MP_OWN = MP_THREADS - 2 * MP_HALO;  // 192

__global__ void __launch_bounds__(MP_THREADS)
mel_post_kernel(const aegis_melpost_params p) {
    __shared__ unsigned char flag[MP_THREADS];
    const int clip = blockIdx.y;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int T = p.n_frames;
    // the 32-column halos only serve the run-length gate of the rake mask; without it every thread owns its column
    const int halo = p.rake_mask != nullptr ? MP_HALO : 0;
    const int own_n = MP_THREADS - 2 * halo;
    const int t = blockIdx.x * own_n - halo + tid;
    const bool in_clip = (t >= 0 && t < T);
    const bool own = in_clip && tid >= halo && tid < halo + own_n;

    const float* __restrict__ mel = p.mel + static_cast<long long>(clip) * p.mel_clip_stride;
    const float amin = 1e-10f;
    const bool is_db = p.input_is_db != 0;
    const float max_db = is_db ? 0.f : db10(fmaxf(amin, __ldg(p.mel_max + clip)));
    const float ref_db = is_db ? 0.f : (p.ref_power ? db10(fmaxf(amin, __ldg(p.ref_power + clip))) : max_db);
    const float db_floor = is_db ? -INFINITY : (max_db - ref_db) - 80.0f;  // max(log_spec) - top_db
    const float onset_floor = max_db - 80.0f;  // power_to_db(ref=1.0): max(log_spec) - top_db

    // pass A: column maximum (dB is monotone in power, so dB(max) == max(dB)); only the rake test needs it
    const bool want_rake = p.rake_mask != nullptr;
    float col_max_db = -80.0f;
    if (in_clip && want_rake) {
        if (is_db) {
            float cmax = -INFINITY;
#pragma unroll 8
            for (int m = 0; m < p.n_mels; ++m) cmax = fmaxf(cmax, __ldg(mel + static_cast<long long>(m) * p.mel_row_stride + t));
            col_max_db = cmax;
        } else {
            float cmax = 0.f;
#pragma unroll 8
            for (int m = 0; m < p.n_mels; ++m) cmax = fmaxf(cmax, __ldg(mel + static_cast<long long>(m) * p.mel_row_stride + t));
            col_max_db = fmaxf(db10(fmaxf(amin, cmax)) - ref_db, db_floor);
        }
    }
    // pass B: dB per cell -> optional store, broadband count, positive flux against column t-1.  Four mel rows per
    // step: the loads are issued together, and lane 0 -- whose left neighbour belongs to another warp -- fetches its
    // four previous-column values in ONE divergent region (per row it cost as much as the rest of the loop body).
    int active = 0;
    float flux = 0.f;
    const float thr = col_max_db - 20.0f;
    const bool want_flux = p.onset_env != nullptr && !is_db;
    float* __restrict__ sdb = p.s_db ? p.s_db + static_cast<long long>(clip) * p.sdb_clip_stride : nullptr;
    const bool need_db = want_rake || sdb != nullptr;
    const bool fix0 = want_flux && lane == 0 && in_clip && t >= 1;
    constexpr int MB = 4;
    for (int m0 = 0; m0 < p.n_mels; m0 += MB) {
        float raw[MB], L[MB], Dp[MB];
#pragma unroll
        for (int i = 0; i < MB; ++i)
            raw[i] = (in_clip && m0 + i < p.n_mels) ? __ldg(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t) : (is_db ? 0.f : amin);
#pragma unroll
        for (int i = 0; i < MB; ++i) L[i] = in_clip ? (is_db ? raw[i] : db10(fmaxf(amin, raw[i]))) : 0.f;
        if (need_db) {
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                if (m0 + i < p.n_mels) {
                    const float db = fmaxf(L[i] - ref_db, db_floor);
                    if (own && sdb) sdb[static_cast<long long>(m0 + i) * p.sdb_row_stride + t] = db;
                    active += (db > thr) ? 1 : 0;
                }
            }
        }
        if (want_flux) {
#pragma unroll
            for (int i = 0; i < MB; ++i) Dp[i] = __shfl_up_sync(0xffffffffu, fmaxf(L[i], onset_floor), 1);
            if (fix0) {
#pragma unroll
                for (int i = 0; i < MB; ++i)
                    if (m0 + i < p.n_mels)
                        Dp[i] = fmaxf(db10(fmaxf(amin, __ldg(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t - 1))), onset_floor);
            }
#pragma unroll
            for (int i = 0; i < MB; ++i)
                if (m0 + i < p.n_mels) flux += fmaxf(0.0f, fmaxf(L[i], onset_floor) - Dp[i]);
        }
    }
    bool is_rake = false;
    if (in_clip && !(col_max_db < -60.0f)) {
        is_rake = (static_cast<double>(active) / static_cast<double>(p.n_mels)) > p.rake_ratio;
    }
    flag[tid] = is_rake ? 1 : 0;
    __syncthreads();

    if (own && p.rake_mask) {
        bool keep = false;
        if (is_rake) {
            const int maxf = p.rake_max_frames;
            int left = 0;   // rake columns directly before t
            while (left <= maxf && tid - left - 1 >= 0 && t - left - 1 >= 0 && flag[tid - left - 1]) ++left;
            int right = 0;  // rake columns directly after t
            while (right <= maxf && tid + right + 1 < MP_THREADS && t + right + 1 < T && flag[tid + right + 1]) ++right;
            const int len = left + right + 1;
            const bool closed = (t + right + 1) < T;  // a run still open at the last column is never emitted
            keep = closed && left <= maxf && right <= maxf && len >= p.rake_min_frames && len <= maxf;
        }
        p.rake_mask[static_cast<long long>(clip) * T + t] = keep ? 1 : 0;
    }

    if (want_flux) {
        float* __restrict__ env = p.onset_env + static_cast<long long>(clip) * T;
        float vmin = __int_as_float(0x7f800000), vmax = 0.f;
        if (own) {
            if (t < p.onset_pad) {  // leading zeros of the padded envelope
                env[t] = 0.f;
                vmin = 0.f;
            }
            const int e = t + p.onset_pad - 1;  // flux(t) = raw[t-1] lands at raw index + pad
            if (t >= 1 && e < T) {
                const float v = flux * (1.0f / static_cast<float>(p.n_mels));
                env[e] = v;
                vmin = fminf(vmin, v);
                vmax = fmaxf(vmax, v);
            }
        }
        if (p.env_minmax) {
            vmin = warp_min(vmin);
            vmax = warp_max(vmax);
            if (lane == 0) {
                if (vmin < __int_as_float(0x7f800000)) atomic_min_nonneg(p.env_minmax + 2 * clip, vmin);
                atomic_max_nonneg(p.env_minmax + 2 * clip + 1, vmax);
            }
        }
    }
}

// Onset envelope only (no dB image, no rake mask: the cfg2 step).  Same arithmetic per cell as mel_post_kernel -- one
// hardware log2, the floor at max - 80 dB, positive differences against column t - 1 summed over the mel rows in
// ascending order -- but a thread owns FOUR consecutive columns and reads them with one 16-byte load per row, four
// rows in flight: the one-column-per-thread kernel moved 0.68 GB in 0.31 ms (2.2 TB/s) because a warp's request per
// row was only 128 bytes.  The left neighbour of a thread's first column comes from the previous lane's fourth
// column by shuffle; lane 0 reads it from memory.
constexpr int OF_THREADS = 128;

__global__ void __launch_bounds__(OF_THREADS)
onset_flux4_kernel(const aegis_melpost_params p, int groups_per_clip, int blocks_per_clip) {
    const int clip = blockIdx.x / blocks_per_clip;
    const int grp = (blockIdx.x - clip * blocks_per_clip) * OF_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int T = p.n_frames;
    const int t0 = 4 * grp;
    const bool live = grp < groups_per_clip;             // whole warps stay in the loop for the shuffles
    const float* __restrict__ mel = p.mel + static_cast<long long>(clip) * p.mel_clip_stride;
    const float amin = 1e-10f;
    const float onset_floor = db10(fmaxf(amin, __ldg(p.mel_max + clip))) - 80.0f;   // power_to_db(ref=1.0): max - top_db
    const bool fix0 = lane == 0 && live && t0 >= 1;
    float fx = 0.f, fy = 0.f, fz = 0.f, fw = 0.f;
    constexpr int MB = 4;
    for (int m0 = 0; m0 < p.n_mels; m0 += MB) {
        float4 raw[MB];
        float left[MB];
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const bool ok = live && m0 + i < p.n_mels;
            raw[i] = ok ? __ldg(reinterpret_cast<const float4*>(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t0))
                        : make_float4(amin, amin, amin, amin);
            left[i] = (fix0 && m0 + i < p.n_mels) ? __ldg(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t0 - 1) : amin;
        }
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const float lx = fmaxf(db10(fmaxf(amin, raw[i].x)), onset_floor), ly = fmaxf(db10(fmaxf(amin, raw[i].y)), onset_floor);
            const float lz = fmaxf(db10(fmaxf(amin, raw[i].z)), onset_floor), lw = fmaxf(db10(fmaxf(amin, raw[i].w)), onset_floor);
            float prev = __shfl_up_sync(0xffffffffu, lw, 1);
            if (lane == 0) prev = fmaxf(db10(fmaxf(amin, left[i])), onset_floor);
            if (m0 + i < p.n_mels) {
                fx += fmaxf(0.0f, lx - prev);
                fy += fmaxf(0.0f, ly - lx);
                fz += fmaxf(0.0f, lz - ly);
                fw += fmaxf(0.0f, lw - lz);
            }
        }
    }
    float* __restrict__ env = p.onset_env + static_cast<long long>(clip) * T;
    float vmin = __int_as_float(0x7f800000), vmax = 0.f;
    if (live) {
        const float fl[4] = {fx, fy, fz, fw};
        const float inv = 1.0f / static_cast<float>(p.n_mels);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + j;
            if (t >= T) break;
            if (t < p.onset_pad) {       // leading zeros of the padded envelope
                env[t] = 0.f;
                vmin = 0.f;
            }
            const int e = t + p.onset_pad - 1;   // flux(t) = raw[t-1] lands at raw index + pad
            if (t >= 1 && e < T) {
                const float v = fl[j] * inv;
                env[e] = v;
                vmin = fminf(vmin, v);
                vmax = fmaxf(vmax, v);
            }
        }
    }
    if (p.env_minmax) {
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        if (lane == 0) {
            if (vmin < __int_as_float(0x7f800000)) atomic_min_nonneg(p.env_minmax + 2 * clip, vmin);
            atomic_max_nonneg(p.env_minmax + 2 * clip + 1, vmax);
        }
    }
}

// Rake mask only (no dB image stored, no onset envelope: the transcription step).  Same arithmetic per cell as
// mel_post_kernel -- column maximum first, then one hardware log2 per cell, the broadband count against max - 20 dB, the
// run-length gate over the flags of a 256-column CTA with 32-column halos -- but a thread owns FOUR consecutive columns
// and reads them with one 16-byte load per mel row (the one-column kernel's warp request per row is 128 bytes).
constexpr int MR_THREADS = MP_THREADS / 4;   // 64 threads = 256 columns, of which 2 x 32 are halo

__global__ void __launch_bounds__(MR_THREADS)
mel_rake4_kernel(const aegis_melpost_params p) {
    __shared__ unsigned char flag[MP_THREADS];
    const int clip = blockIdx.y;
    const int tid = threadIdx.x;
    const int T = p.n_frames;
    const int t0 = blockIdx.x * MP_OWN - MP_HALO + 4 * tid;          // first of this thread's four columns (a multiple of 4)
    const bool any_in = t0 >= 0 && t0 < T;                           // t0 < 0: all four columns lie before the clip
    const bool own = 4 * tid >= MP_HALO && 4 * tid < MP_HALO + MP_OWN;
    const float* __restrict__ mel = p.mel + static_cast<long long>(clip) * p.mel_clip_stride;
    const float amin = 1e-10f;
    const float max_db = db10(fmaxf(amin, __ldg(p.mel_max + clip)));
    const float ref_db = p.ref_power ? db10(fmaxf(amin, __ldg(p.ref_power + clip))) : max_db;
    const float db_floor = (max_db - ref_db) - 80.0f;
    constexpr int MB = 4;
    float4 cmax = make_float4(0.f, 0.f, 0.f, 0.f);
    if (any_in) {
        for (int m0 = 0; m0 < p.n_mels; m0 += MB) {
            float4 raw[MB];
#pragma unroll
            for (int i = 0; i < MB; ++i)
                raw[i] = (m0 + i < p.n_mels) ? __ldg(reinterpret_cast<const float4*>(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t0))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                cmax.x = fmaxf(cmax.x, raw[i].x); cmax.y = fmaxf(cmax.y, raw[i].y);
                cmax.z = fmaxf(cmax.z, raw[i].z); cmax.w = fmaxf(cmax.w, raw[i].w);
            }
        }
    }
    float col_max_db[4], thr[4];
    int active[4] = {0, 0, 0, 0};
    {
        const float cm[4] = {cmax.x, cmax.y, cmax.z, cmax.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            col_max_db[j] = (any_in && t0 + j < T) ? fmaxf(db10(fmaxf(amin, cm[j])) - ref_db, db_floor) : -80.0f;
            thr[j] = col_max_db[j] - 20.0f;
        }
    }
    if (any_in) {
        for (int m0 = 0; m0 < p.n_mels; m0 += MB) {
            float4 raw[MB];
#pragma unroll
            for (int i = 0; i < MB; ++i)
                raw[i] = (m0 + i < p.n_mels) ? __ldg(reinterpret_cast<const float4*>(mel + static_cast<long long>(m0 + i) * p.mel_row_stride + t0))
                                             : make_float4(amin, amin, amin, amin);
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                if (m0 + i < p.n_mels) {
                    const float v[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float db = fmaxf(db10(fmaxf(amin, v[j])) - ref_db, db_floor);
                        active[j] += (db > thr[j]) ? 1 : 0;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bool is_rake = false;
        if (any_in && t0 + j < T && !(col_max_db[j] < -60.0f))
            is_rake = (static_cast<double>(active[j]) / static_cast<double>(p.n_mels)) > p.rake_ratio;
        flag[4 * tid + j] = is_rake ? 1 : 0;
    }
    __syncthreads();
    if (own) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + j, c = 4 * tid + j;
            if (t >= T) break;
            bool keep = false;
            if (flag[c]) {
                const int maxf = p.rake_max_frames;
                int left = 0;   // rake columns directly before t
                while (left <= maxf && c - left - 1 >= 0 && t - left - 1 >= 0 && flag[c - left - 1]) ++left;
                int right = 0;  // rake columns directly after t
                while (right <= maxf && c + right + 1 < MP_THREADS && t + right + 1 < T && flag[c + right + 1]) ++right;
                const int len = left + right + 1;
                const bool closed = (t + right + 1) < T;  // a run still open at the last column is never emitted
                keep = closed && left <= maxf && right <= maxf && len >= p.rake_min_frames && len <= maxf;
            }
            p.rake_mask[static_cast<long long>(clip) * T + t] = keep ? 1 : 0;
        }
    }
}

// candidate test of librosa.util.peak_pick on the normalised envelope, one thread per frame
__global__ void __launch_bounds__(256)
peak_candidates_kernel(const aegis_peaks_params p) {
    const int clip = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = p.n_frames;
    if (n >= N) return;
    const float* __restrict__ env = p.onset_env + static_cast<long long>(clip) * N;
    const float mn = p.env_minmax[2 * clip], mx = p.env_minmax[2 * clip + 1];
    bool is_cand = false;
    if (mx > 0.f) {  // an all-zero envelope has no onsets
        const float shift = p.normalize ? mn : 0.f;
        const float scale = p.normalize ? ((mx - mn) + 1.17549435e-38f) : 1.f;
        auto x = [&](int i) -> float {
            const float v = __ldg(env + i) - shift;
            return p.normalize ? __fdiv_rn(v, scale) : v;
        };
        const float xn = x(n);
        float wmax = xn;
        const int lo = max(0, n - p.pre_max), hi = min(n + p.post_max, N);
        for (int i = lo; i < hi; ++i) wmax = fmaxf(wmax, x(i));
        if (xn == wmax) {
            const int alo = max(0, n - p.pre_avg), ahi = min(n + p.post_avg, N);
            double acc = 0.0;
            for (int i = alo; i < ahi; ++i) acc += static_cast<double>(x(i));
            const double avg = acc / static_cast<double>(ahi - alo);
            is_cand = static_cast<double>(xn) >= avg + p.delta;
        }
    }
    p.cand[static_cast<long long>(clip) * N + n] = is_cand ? 1 : 0;
}

// greedy left-to-right selection with the `wait` dead time (n = 0; while n < N: if cand[n] { take n; n += wait + 1 }
// else ++n), one warp per clip.  The candidates of 32 frames are gathered into one ballot word per step and every lane
// walks the set bits redundantly: the loop runs once per ACCEPTED peak instead of once per frame (a lane-0 walk over
// dependent byte loads took 90 us for 1292 frames, however small the batch).
__global__ void __launch_bounds__(128)
peak_select_kernel(const aegis_peaks_params p) {
    const int clip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= p.n_clips) return;
    const int lane = threadIdx.x & 31;
    const int N = p.n_frames;
    const unsigned char* __restrict__ cand = p.cand + static_cast<long long>(clip) * N;
    unsigned char* __restrict__ peaks = p.peaks + static_cast<long long>(clip) * N;
    int count = 0;
    long long next_ok = 0;   // first frame that may be taken again
    for (int base = 0; base < N; base += 32) {
        const int n = base + lane;
        unsigned m = __ballot_sync(0xffffffffu, n < N && cand[n] != 0);
        unsigned taken = 0;
        while (m) {
            if (next_ok - base >= 32) break;
            if (next_ok > base) m &= ~0u << static_cast<int>(next_ok - base);   // inside the dead time
            if (!m) break;
            const int b = __ffs(m) - 1;
            taken |= 1u << b;
            ++count;
            next_ok = static_cast<long long>(base) + b + p.wait + 1;
            m &= ~(1u << b);
        }
        if (n < N) peaks[n] = (taken >> lane) & 1u;
    }
    if (lane == 0) p.n_peaks[clip] = count;
}

}  // namespace aegis

extern "C" int aegis_mel_post(const aegis_melpost_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->mel && (p->mel_max || p->input_is_db), "aegis_mel_post: mel / mel_max must be set");
    AEGIS_REQUIRE(p->n_mels > 0 && p->n_clips >= 0 && p->n_frames >= 0, "aegis_mel_post: bad sizes");
    AEGIS_REQUIRE(p->rake_max_frames >= 0 && p->rake_max_frames <= MP_HALO - 2,
                  "aegis_mel_post: rake_max_frames=%d exceeds the %d-column halo", p->rake_max_frames, MP_HALO - 2);
    AEGIS_REQUIRE(p->onset_env == nullptr || p->onset_pad >= 1, "aegis_mel_post: onset_pad must be >= 1");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    // onset envelope alone, rows that can be read in 16-byte pieces (pitch >= T rounded up to 4): four columns per thread
    const bool flux_only = p->onset_env != nullptr && p->rake_mask == nullptr && p->s_db == nullptr && !p->input_is_db &&
                           p->ref_power == nullptr && (p->mel_row_stride % 4) == 0 && (p->mel_clip_stride % 4) == 0 &&
                           (reinterpret_cast<uintptr_t>(p->mel) % 16) == 0 && p->mel_row_stride >= ((p->n_frames + 3) / 4) * 4;
    if (flux_only) {
        const int groups = (p->n_frames + 3) / 4;
        const int blocks_per_clip = (groups + OF_THREADS - 1) / OF_THREADS;
        AEGIS_REQUIRE(static_cast<long long>(blocks_per_clip) * p->n_clips < (1LL << 31), "aegis_mel_post: too many blocks for one launch");
        onset_flux4_kernel<<<static_cast<unsigned>(blocks_per_clip) * p->n_clips, OF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p, groups, blocks_per_clip);
        return check_launch("aegis_mel_post(onset flux)");
    }
    const bool rows16 = (p->mel_row_stride % 4) == 0 && (p->mel_clip_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(p->mel) % 16) == 0 &&
                        p->mel_row_stride >= ((p->n_frames + 3) / 4) * 4;
    if (p->rake_mask != nullptr && p->s_db == nullptr && p->onset_env == nullptr && !p->input_is_db && rows16) {   // rake mask alone
        dim3 grid4((p->n_frames + MP_OWN - 1) / MP_OWN, p->n_clips);
        mel_rake4_kernel<<<grid4, MR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p);
        return check_launch("aegis_mel_post(rake mask)");
    }
    const int own_n = p->rake_mask != nullptr ? MP_OWN : MP_THREADS;
    dim3 grid((p->n_frames + own_n - 1) / own_n, p->n_clips);
    mel_post_kernel<<<grid, MP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p);
    return check_launch("aegis_mel_post");
}

extern "C" int aegis_onset_peaks(const aegis_peaks_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->onset_env && p->env_minmax && p->cand && p->peaks && p->n_peaks,
                  "aegis_onset_peaks: null pointer");
    AEGIS_REQUIRE(p->pre_max >= 0 && p->post_max >= 1 && p->pre_avg >= 0 && p->post_avg >= 1 && p->wait >= 0,
                  "aegis_onset_peaks: bad window parameters");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((p->n_frames + 255) / 256, p->n_clips);
    peak_candidates_kernel<<<grid, 256, 0, st>>>(*p);
    if (int rc = check_launch("aegis_onset_peaks(candidates)")) return rc;
    peak_select_kernel<<<(p->n_clips + 3) / 4, 128, 0, st>>>(*p);
    return check_launch("aegis_onset_peaks(select)");
}
