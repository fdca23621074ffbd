// K6: electric-guitar filters of the v2 engine on the perception outputs.
//
// Replaces apply_guitar_filters (aegis_engine_core_v2/guitar_specific.py:240-277, called from
// aegis_engine_financial.py:132-147): sub-E2 octave correction (:24-61), enhanced rake mask (:112-151),
// palm-mute mask (:63-110), distortion class (:209-233).
//
// One thread per spectrogram column (coalesced along time), CTAs of 256 columns of which the outer 32 on each
// side are halo: the run-length gate of the mute mask and the look-back of the rake enhancement reach at most
// 31 frames.  The reference's arithmetic is float32 numpy: column means are sequential row sums (np.mean over
// axis 0 adds row after row) divided by the row count, and the mean of the `rake_frames` energy differences is
// numpy's reduction (first element + pairwise sum of the rest); both are restated operation for operation
// (no FMA contraction: __fadd_rn / __fdiv_rn), so the masks are bit-exact.  The distortion class compares a ratio
// of two global means with 0.25 / 0.4; the sums are accumulated in double per column block (fixed order) and
// finished by a second kernel.
#include "common.cuh"

namespace aegis {

constexpr int GF_THREADS = 256;
constexpr int GF_HALO = 32;
constexpr int GF_OWN = GF_THREADS - 2 * GF_HALO;  // 192

// numpy's float32 add.reduce over a contiguous 1-D slice a[0..n), n <= 128: pairwise_sum's leaf (eight running
// sums combined as a tree, the remainder added one by one; fewer than eight elements are added in order)
__device__ __forceinline__ float numpy_sum_f32(const float* a, int n) {
    float res;
    if (n < 8) {
        res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
    } else {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], a[i + j]);
        }
        res = __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3])),
                        __fadd_rn(__fadd_rn(acc[4], acc[5]), __fadd_rn(acc[6], acc[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    }
    return res;
}

__global__ void __launch_bounds__(GF_THREADS)
guitar_columns_kernel(const aegis_guitar_params p) {
    __shared__ float total[GF_THREADS + 1];     // mean dB of column t (index 0: column before the block)
    __shared__ float diff[GF_THREADS];          // energy_diff[t] = total[t] - total[t-1]
    __shared__ unsigned char mute_col[GF_THREADS];
    __shared__ unsigned char trigger[GF_THREADS];
    __shared__ double red[2][GF_THREADS / 32];
    const int clip = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.n_frames;
    const int t = blockIdx.x * GF_OWN - GF_HALO + tid;
    const bool in_clip = (t >= 0 && t < T);
    const bool own = in_clip && tid >= GF_HALO && tid < GF_HALO + GF_OWN;
    const long long fidx = static_cast<long long>(clip) * T + t;

    // ---- sub-E2 filter (guitar_specific.py:24-61): elementwise on the owned columns
    if (own && p.f0_out != nullptr) {
        const double f = p.f0[fidx];
        double fo = f;
        bool vo = p.voiced != nullptr ? p.voiced[fidx] != 0 : !(f != f);
        if (f < p.fmin_hz) {                       // false for NaN
            const double c = f * 2;
            if (p.fmin_hz <= c && c < p.fmin_hz * 4) {
                fo = c;
                vo = true;
            } else {
                fo = __longlong_as_double(0x7ff8000000000000LL);
                vo = false;
            }
        }
        p.f0_out[fidx] = fo;
        if (p.voiced_out != nullptr) p.voiced_out[fidx] = vo ? 1 : 0;
    }
    if (p.s_db == nullptr) return;

    // ---- column sums, float32, row after row (np.mean(S_dB[...], axis=0))
    const float* __restrict__ img = p.s_db + static_cast<long long>(clip) * p.sdb_clip_stride;
    const int n_mels = p.n_mels, mid = n_mels / 2, hi0 = static_cast<int>(n_mels * 0.7);
    float s_low = 0.f, s_high = 0.f, s_all = 0.f;
    double d_all = 0.0, d_hi = 0.0;
    if (in_clip) {
        for (int m = 0; m < n_mels; ++m) {
            const float v = __ldg(img + static_cast<long long>(m) * p.sdb_row_stride + t);
            if (m == 0) s_all = v; else s_all = __fadd_rn(s_all, v);
            if (m < mid) { if (m == 0) s_low = v; else s_low = __fadd_rn(s_low, v); }
            else { if (m == mid) s_high = v; else s_high = __fadd_rn(s_high, v); }
            if (own) {
                d_all += static_cast<double>(v);
                if (m >= hi0) d_hi += static_cast<double>(v);
            }
        }
    }
    // palm-mute column test (:84-94)
    bool mute = false;
    if (in_clip && mid > 0) {
        const float low = __fdiv_rn(s_low, static_cast<float>(mid));
        const float high = __fdiv_rn(s_high, static_cast<float>(n_mels - mid));
        mute = __fdiv_rn(low, __fadd_rn(high, 1e-6f)) > 2.0f;
    }
    mute_col[tid] = mute ? 1 : 0;
    const float tot = in_clip ? __fdiv_rn(s_all, static_cast<float>(n_mels)) : 0.f;
    total[tid + 1] = tot;
    if (tid == 0) {   // column before the block (for the first difference)
        float prev = 0.f;
        if (t - 1 >= 0 && t - 1 < T) {
            float sa = 0.f;
            for (int m = 0; m < n_mels; ++m) {
                const float v = __ldg(img + static_cast<long long>(m) * p.sdb_row_stride + t - 1);
                sa = (m == 0) ? v : __fadd_rn(sa, v);
            }
            prev = __fdiv_rn(sa, static_cast<float>(n_mels));
        }
        total[0] = prev;
    }
    __syncthreads();
    // energy_diff = np.diff(total, prepend=total[0]) (:134)
    diff[tid] = (in_clip && t >= 1) ? __fadd_rn(total[tid + 1], -total[tid]) : 0.f;
    __syncthreads();

    // ---- rake enhancement (:136-149): trigger(i) marks [i, i + n)
    const int nr = p.rake_frames;
    bool trig = false;
    if (in_clip && t >= 1 && nr > 0 && diff[tid] > 10.0f && t + nr < T && tid + nr <= GF_THREADS) {
        const float mean = __fdiv_rn(numpy_sum_f32(&diff[tid], nr), static_cast<float>(nr));
        trig = mean < 0.f;
    }
    trigger[tid] = trig ? 1 : 0;
    __syncthreads();

    if (own) {
        if (p.rake_out != nullptr) {
            bool r = p.rake_in != nullptr ? p.rake_in[fidx] != 0 : false;
            for (int k = 0; k < nr && !r; ++k)
                if (tid - k >= 0 && trigger[tid - k]) r = true;
            p.rake_out[fidx] = r ? 1 : 0;
        }
        if (p.mute_out != nullptr) {   // closed runs of at most mute_max_frames columns (:100-110)
            bool keep = false;
            if (mute) {
                const int maxf = p.mute_max_frames;
                int left = 0;
                while (left <= maxf && tid - left - 1 >= 0 && t - left - 1 >= 0 && mute_col[tid - left - 1]) ++left;
                int right = 0;
                while (right <= maxf && tid + right + 1 < GF_THREADS && t + right + 1 < T && mute_col[tid + right + 1]) ++right;
                const bool closed = (t + right + 1) < T;   // a run still open at the last column is never emitted
                keep = closed && left <= maxf && right <= maxf && (left + right + 1) <= maxf;
            }
            p.mute_out[fidx] = keep ? 1 : 0;
        }
    }
    // ---- distortion: per-block partial sums in a fixed order
    if (p.distortion != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d_all += __shfl_xor_sync(0xffffffffu, d_all, o);
            d_hi += __shfl_xor_sync(0xffffffffu, d_hi, o);
        }
        if (lane == 0) { red[0][warp] = d_all; red[1][warp] = d_hi; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, h = 0.0;
            for (int w = 0; w < GF_THREADS / 32; ++w) { a += red[0][w]; h += red[1][w]; }
            double* dst = p.dist_work + (static_cast<long long>(clip) * gridDim.x + blockIdx.x) * 2;
            dst[0] = a;
            dst[1] = h;
        }
    }
}

// distortion class of every clip from the block sums (:209-233)
__global__ void guitar_distortion_kernel(const aegis_guitar_params p, const int n_blocks) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= p.n_clips) return;
    double a = 0.0, h = 0.0;
    for (int b = 0; b < n_blocks; ++b) {
        a += p.dist_work[(static_cast<long long>(clip) * n_blocks + b) * 2];
        h += p.dist_work[(static_cast<long long>(clip) * n_blocks + b) * 2 + 1];
    }
    const int hi0 = static_cast<int>(p.n_mels * 0.7);
    const double total = a / (static_cast<double>(p.n_mels) * p.n_frames);
    const double high = h / (static_cast<double>(p.n_mels - hi0) * p.n_frames);
    const double ratio = high / (total + 1e-6);
    p.distortion[clip] = ratio > 0.4 ? 2 : (ratio > 0.25 ? 1 : 0);
}

}  // namespace aegis

extern "C" int aegis_guitar_blocks(int n_frames) {
    return n_frames <= 0 ? 0 : (n_frames + aegis::GF_OWN - 1) / aegis::GF_OWN;
}

extern "C" int aegis_guitar_filters(const aegis_guitar_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "aegis_guitar_filters: null params");
    AEGIS_REQUIRE(p->n_clips >= 0 && p->n_frames >= 0, "aegis_guitar_filters: negative size");
    AEGIS_REQUIRE(p->f0_out == nullptr || p->f0 != nullptr, "aegis_guitar_filters: f0_out needs f0");
    AEGIS_REQUIRE(p->voiced_out == nullptr || p->f0_out != nullptr, "aegis_guitar_filters: voiced_out needs f0_out");
    const bool want_img = p->rake_out || p->mute_out || p->distortion;
    AEGIS_REQUIRE(!want_img || p->s_db != nullptr, "aegis_guitar_filters: rake_out / mute_out / distortion need s_db");
    if (want_img) {
        AEGIS_REQUIRE(p->n_mels >= 2 && p->sdb_row_stride >= p->n_frames, "aegis_guitar_filters: bad dB image shape");
        AEGIS_REQUIRE(p->mute_max_frames >= 0 && p->mute_max_frames <= GF_HALO - 2 && p->rake_frames >= 0 && p->rake_frames <= GF_HALO - 2,
                      "aegis_guitar_filters: mute_max_frames=%d / rake_frames=%d unsupported (<= %d: hop too small for the 32-column halo)",
                      p->mute_max_frames, p->rake_frames, GF_HALO - 2);
        AEGIS_REQUIRE(p->distortion == nullptr || p->dist_work != nullptr, "aegis_guitar_filters: distortion needs dist_work");
    }
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    aegis_guitar_params q = *p;
    if (!want_img) q.s_db = nullptr;
    const int n_blocks = aegis_guitar_blocks(p->n_frames);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    guitar_columns_kernel<<<dim3(n_blocks, p->n_clips), GF_THREADS, 0, st>>>(q);
    if (int rc = check_launch("aegis_guitar_filters(columns)")) return rc;
    if (want_img && p->distortion != nullptr) {
        guitar_distortion_kernel<<<(p->n_clips + 127) / 128, 128, 0, st>>>(q, n_blocks);
        return check_launch("aegis_guitar_filters(distortion)");
    }
    return 0;
}
