// K1: fused frame gather + Hann window + FP32 real FFT -> |X| (+ mel power, frame RMS).
//
// Replaces librosa.stft / feature.melspectrogram / feature.rms as called from
// aegis_engine.py:25,70 and aegis_engine_financial.py:46-50,154 (see include/aegis_b200.h).
//
// The kernel is bound by FP32 lanes, issue slots and shared-memory bandwidth together (SURVEY H3:
// ~9 FLOP per HBM byte), so the design minimises all three per frame:
//   * one WARP transforms TWO frames at once in packed f32x2 arithmetic (rfft2048x2.cuh): half the
//     issue slots of scalar code, no cross-talk between the frames, no pre-scaling;
//   * each frame is ONE 1024-point complex transform (even/odd packing) done as 32 x 32 in
//     registers with a single exchange through shared memory, inside the warp (no CTA barrier);
//   * window, twiddles and mel weights are read once per frame PAIR.
// Work unit ("tile") = 8 consecutive frames of one clip = 4 warps ("group").  One persistent CTA per
// SM holds two groups that run out of phase: while one streams its magnitudes to HBM and projects
// them on the mel triangles, the other transforms.  Per tile and group:
//   1. (7*hop + 2048) samples in shared memory (cp.async prefetch issued during the previous tile;
//      zeros outside the clip = centre padding)
//   2. each warp: samples * window -> registers (sum of squares for the RMS on the way), pass 1,
//      exchange, pass 2, Z -> shared, conjugate-pair split, |X| for its two frames -> shared
//   3. the group writes |X| in librosa's [1025, T] layout, one 32-byte sector (8 frames) per row,
//      and reduces |X|^2 over the mel triangles (each bin read once: rise/fall partial sums per
//      band-edge segment).
// HBM traffic per frame: hop*4 B read (+ halo, L2-served) and 1025*4 B written.
#include "common.cuh"
#include "rfft2048x2.cuh"

namespace aegis {

constexpr int TILE_F = 8;
constexpr int GROUP_WARPS = TILE_F / 2;
constexpr int GROUP_THREADS = GROUP_WARPS * 32;            // 128
constexpr int STFT_GROUPS = 2;
constexpr int STFT_THREADS = STFT_GROUPS * GROUP_THREADS;  // 256
constexpr int MAX_HOP = 512;
constexpr int SAMPLES_MAX = (TILE_F - 1) * MAX_HOP + RF_N;  // 5632
constexpr int MEL_MAX = 128;
constexpr int MEL_PART = (MEL_MAX + 1) * TILE_F;            // floats per rise / fall partial array

// One warp's exchange buffer (re / im planes of its two frames, rfft2048x2.cuh).  After pass 2 has loaded them the
// same memory holds the warp's two magnitude columns as p2 mag[1025] (8200 B) and, above MEL_PART_OFFSET, one of
// the group's mel partial-sum arrays.  The tail pads the stride to 8 banks (mod 32): the 8 frames of one bin sit
// in 8 distinct banks.
struct alignas(16) WarpBuf {
    float e[RF_WARP_WORDS];
    float skew[24];
};
static_assert((sizeof(WarpBuf) / 4) % 32 == 8, "consecutive warp buffers must be skewed by 8 banks");
constexpr int MEL_PART_OFFSET = 8448;  // bytes; >= 1025 * 8
static_assert(MEL_PART_OFFSET >= RF_BINS * 8 && MEL_PART_OFFSET + MEL_PART * 4 <= RF_WARP_WORDS * 4, "mel partials alias the warp buffer");

struct StftSmem {
    WarpBuf wb[STFT_GROUPS * GROUP_WARPS];
    float samples[STFT_GROUPS][SAMPLES_MAX];
    float window[RF_N];              // 0.5 * analysis window (the split produces 2 X)
    cf32 tw1[32 * 32];               // [b][lane] W1024^{lane b}
    cf32 tw2[RF_M];                  // W2048^k
    float4 mel_rf[RF_BINS];          // (rise, rise, fall, fall) weight of every FFT bin
    int mel_seg[MEL_MAX + 2];        // first bin of every segment between mel band edges
};
static_assert(sizeof(StftSmem) <= 227 * 1024, "shared memory budget of one SM");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }


// synchronous tile fill with bounds checks (clip edges, unaligned rows)
__device__ __forceinline__ void fill_samples(float* smp, const float* __restrict__ yc, long long g0, long long N, int n_buf, int gt) {
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(yc) & 15) == 0) && ((g0 & 3) == 0);
    for (int i = gt * 4; i < n_buf; i += GROUP_THREADS * 4) {
        const long long gi = g0 + i;
        float4 v;
        if (vec_ok && gi >= 0 && gi + 3 < N) {
            v = __ldg(reinterpret_cast<const float4*>(yc + gi));
        } else {
            v.x = (gi >= 0 && gi < N) ? __ldg(yc + gi) : 0.f;
            v.y = (gi + 1 >= 0 && gi + 1 < N) ? __ldg(yc + gi + 1) : 0.f;
            v.z = (gi + 2 >= 0 && gi + 2 < N) ? __ldg(yc + gi + 2) : 0.f;
            v.w = (gi + 3 >= 0 && gi + 3 < N) ? __ldg(yc + gi + 3) : 0.f;
        }
        *reinterpret_cast<float4*>(&smp[i]) = v;
    }
}

__global__ void __launch_bounds__(STFT_THREADS, 1)
stft_fused_kernel(const aegis_stft_params p, const int tiles_per_clip, const long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = tid >> 7, gt = tid & 127, lane = tid & 31, wg = gt >> 5;
    const int T = p.n_frames;
    const int hop = p.hop;
    const long long N = p.n_samples;

    {   // constant tables (once per CTA)
        const cf32* tab = reinterpret_cast<const cf32*>(p.twiddle);
        for (int i = tid; i < RF_N; i += STFT_THREADS) s.window[i] = 0.5f * p.window[i];
        for (int i = tid; i < 32 * 32; i += STFT_THREADS) s.tw1[i] = tab[(2 * (i & 31) * (i >> 5)) & (RF_N - 1)];
        for (int i = tid; i < RF_M; i += STFT_THREADS) s.tw2[i] = tab[i];
        if (p.mel != nullptr) {
            for (int i = tid; i < RF_BINS; i += STFT_THREADS) {
                const float2 rf = reinterpret_cast<const float2*>(p.mel_rise_fall)[i];
                s.mel_rf[i] = make_float4(rf.x, rf.x, rf.y, rf.y);
            }
            for (int i = tid; i < p.n_mels + 2; i += STFT_THREADS) s.mel_seg[i] = p.mel_seg_start[i];
        }
    }
    __syncthreads();

    const int n_buf = (TILE_F - 1) * hop + RF_N;
    const bool do_fft = (p.mag != nullptr) || (p.mel != nullptr);
    float* const smp = s.samples[g];
    WarpBuf* const gbuf = &s.wb[g * GROUP_WARPS];
    float* const xbuf = gbuf[wg].e;
    float* const rise = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(gbuf[0].e) + MEL_PART_OFFSET);
    float* const fall = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(gbuf[1].e) + MEL_PART_OFFSET);
    const long long tile_step = static_cast<long long>(gridDim.x) * STFT_GROUPS;
    bool prefetched = false;  // this tile's samples were requested with cp.async during the previous tile

    for (long long tile = static_cast<long long>(blockIdx.x) * STFT_GROUPS + g; tile < n_tiles; tile += tile_step) {
        const int clip = static_cast<int>(tile / tiles_per_clip);
        const int t0 = static_cast<int>(tile - static_cast<long long>(clip) * tiles_per_clip) * TILE_F;
        const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;

        if (prefetched) cp_async_wait_all();
        named_barrier(1 + g, GROUP_THREADS);  // previous tile's buffers are free; prefetched samples will be complete
        if (!prefetched) {
            fill_samples(smp, yc, static_cast<long long>(t0) * hop - p.pad, N, n_buf, gt);
            named_barrier(1 + g, GROUP_THREADS);
        }

        // ---- 2a. this warp's frame pair -> registers (windowed), sums of squares for the RMS
        c2 v[32];
        {
            const float* fa = smp + (2 * wg) * hop + 2 * lane;
            const float* wp = s.window + 2 * lane;
            float sa, sb;
            if (hop == 512) {
                // frame B = frame A shifted by 8 x 64 samples: lane reads 40 sample pairs instead of 64, each window
                // pair once; the sum of squares of the 24 shared pairs is shared too
                float2 wv[32];
                p2 head = p2{0.f, 0.f}, mid = p2{0.f, 0.f}, tail = p2{0.f, 0.f};
#pragma unroll
                for (int a = 0; a < 40; ++a) {
                    const float2 x = *reinterpret_cast<const float2*>(fa + 64 * a);
                    const p2 xs = p2{x.x, x.y};
                    if (a < 8) head = pfma(xs, xs, head);
                    else if (a < 32) mid = pfma(xs, xs, mid);
                    else tail = pfma(xs, xs, tail);
                    if (a < 32) {
                        wv[a] = *reinterpret_cast<const float2*>(wp + 64 * a);
                        v[a].re.x = x.x * wv[a].x;
                        v[a].im.x = x.y * wv[a].y;
                    }
                    if (a >= 8) {
                        v[a - 8].re.y = x.x * wv[a - 8].x;
                        v[a - 8].im.y = x.y * wv[a - 8].y;
                    }
                }
                const p2 ta = head + mid, tb = mid + tail;
                sa = ta.x + ta.y;
                sb = tb.x + tb.y;
            } else {
                const float* fb = fa + hop;
                p2 ssa = p2{0.f, 0.f}, ssb = p2{0.f, 0.f};
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float2 xa = *reinterpret_cast<const float2*>(fa + 64 * a);
                    const float2 xb = *reinterpret_cast<const float2*>(fb + 64 * a);
                    const float2 w = *reinterpret_cast<const float2*>(wp + 64 * a);
                    ssa = pfma(p2{xa.x, xa.y}, p2{xa.x, xa.y}, ssa);
                    ssb = pfma(p2{xb.x, xb.y}, p2{xb.x, xb.y}, ssb);
                    v[a] = c2{p2{xa.x * w.x, xb.x * w.x}, p2{xa.y * w.y, xb.y * w.y}};
                }
                sa = ssa.x + ssa.y;
                sb = ssb.x + ssb.y;
            }
            if (p.rms != nullptr) {
                sa = warp_sum(sa);
                sb = warp_sum(sb);
                const int t = t0 + 2 * wg + lane;
                if (lane < 2 && t < T)
                    p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t] = sqrtf((lane == 0 ? sa : sb) * (1.0f / RF_N));
            }
        }
        // every warp of the group has read its samples: request the next tile's samples now, so their DRAM latency
        // hides behind the transforms, the stores and the mel projection of this tile
        named_barrier(1 + g, GROUP_THREADS);
        {
            prefetched = false;
            const long long nt = tile + tile_step;
            if (nt < n_tiles) {
                const int nclip = static_cast<int>(nt / tiles_per_clip);
                const int nt0 = static_cast<int>(nt - static_cast<long long>(nclip) * tiles_per_clip) * TILE_F;
                const float* nyc = p.y + static_cast<long long>(nclip) * p.clip_stride;
                const long long ng0 = static_cast<long long>(nt0) * hop - p.pad;
                if (ng0 >= 0 && ng0 + n_buf <= N && ((ng0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(nyc) & 15) == 0)) {
                    for (int i = gt * 4; i < n_buf; i += GROUP_THREADS * 4) cp_async16(&smp[i], nyc + ng0 + i);
                    prefetched = true;
                }
            }
            cp_async_commit();
        }
        if (!do_fft) continue;  // RMS-only call (librosa.feature.rms)

        // ---- 2b. two real 2048-point transforms, inside the warp
        rfft_pass1(lane, v, s.tw1, xbuf);
        __syncwarp();
        rfft_pass2_load(lane, xbuf, v);
        __syncwarp();  // the planes are dead: the magnitudes go over them
        fft32(v);
        {
            float* const mcol = xbuf + (lane >> 4);  // p2 mag[k]: word 2k + frame
            rfft_split_emit(lane, v, s.tw2, [&](int k, float pw) { mcol[2 * k] = sqrt_approx(pw); });
        }
        named_barrier(1 + g, GROUP_THREADS);  // the group's 8 magnitude columns are complete

        // ---- 3a. |X| -> HBM, librosa layout
        const bool full_tile = (t0 + TILE_F <= T);
        if (p.mag != nullptr) {
            float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0;
            const bool vec_store = ((p.mag_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 15) == 0);
            if (vec_store) {  // one float4 (4 frames) per lane, two lanes per spectrogram row
                const int h = gt & 1, kq = gt >> 1;
                const p2* m0p = reinterpret_cast<const p2*>(gbuf[2 * h].e) + kq;
                const p2* m1p = reinterpret_cast<const p2*>(gbuf[2 * h + 1].e) + kq;
                float* dst = mo + static_cast<long long>(kq) * p.mag_row_stride + 4 * h;
                const long long dstep = 64LL * p.mag_row_stride;
                if (full_tile) {
#pragma unroll 8
                    for (int it = 0; it < 16; ++it) {
                        const p2 m0 = m0p[64 * it], m1 = m1p[64 * it];
                        *reinterpret_cast<float4*>(dst) = make_float4(m0.x, m0.y, m1.x, m1.y);
                        dst += dstep;
                    }
                    if (gt < 2) {
                        const p2 m0 = m0p[1024], m1 = m1p[1024];
                        *reinterpret_cast<float4*>(dst) = make_float4(m0.x, m0.y, m1.x, m1.y);
                    }
                } else {
                    for (int k = kq; k < RF_BINS; k += 64) {
                        const p2 m0 = m0p[k - kq], m1 = m1p[k - kq];
                        const float st[4] = {m0.x, m0.y, m1.x, m1.y};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (t0 + 4 * h + j < T) dst[j] = st[j];
                        dst += dstep;
                    }
                }
            } else {          // each warp store = 4 rows x 8 frames
                const int f = lane & 7;
                if (t0 + f < T) {
                    const float* col = gbuf[f >> 1].e + (f & 1);
                    for (int k = wg * 4 + (lane >> 3); k < RF_BINS; k += GROUP_WARPS * 4)
                        mo[static_cast<long long>(k) * p.mag_row_stride + f] = col[2 * k];
                }
            }
        }
        // ---- 3b. mel[b] = sum_{k in seg b} rise[k] |X_k|^2 + sum_{k in seg b+1} fall[k] |X_k|^2
        if (p.mel != nullptr) {
            {   // lane = (segment slot, warp buffer): a frame pair per lane, packed accumulation
                const int b4 = lane & 3;
                const p2* col = reinterpret_cast<const p2*>(gbuf[b4].e);
                auto body = [&](int k, p2& r, p2& fl) {
                    const p2 m = col[k];
                    const float4 w = s.mel_rf[k];
                    const p2 pw = m * m;
                    r = pfma(pw, p2{w.x, w.y}, r);
                    fl = pfma(pw, p2{w.z, w.w}, fl);
                };
                for (int j = wg * 8 + (lane >> 2); j <= p.n_mels; j += GROUP_WARPS * 8) {
                    int k = s.mel_seg[j];
                    const int k1 = s.mel_seg[j + 1];
                    p2 r = p2{0.f, 0.f}, fl = p2{0.f, 0.f};
                    for (; k + 4 <= k1; k += 4) {
                        body(k, r, fl);
                        body(k + 1, r, fl);
                        body(k + 2, r, fl);
                        body(k + 3, r, fl);
                    }
                    for (; k < k1; ++k) body(k, r, fl);
                    *reinterpret_cast<p2*>(&rise[j * TILE_F + 2 * b4]) = r;
                    *reinterpret_cast<p2*>(&fall[j * TILE_F + 2 * b4]) = fl;
                }
            }
            named_barrier(1 + g, GROUP_THREADS);
            float* __restrict__ me = p.mel + static_cast<long long>(clip) * p.mel_clip_stride + t0;
            float vmax = 0.f;
            if (full_tile && p.n_mels == MEL_MAX) {
                const int f = gt & 7;
                float* dst = me + static_cast<long long>(gt >> 3) * p.mel_row_stride + f;
                const long long dstep = 16LL * p.mel_row_stride;
#pragma unroll
                for (int it = 0; it < MEL_MAX / 16; ++it) {
                    const float val = rise[gt + 128 * it] + fall[gt + 128 * it + TILE_F];
                    *dst = val;
                    dst += dstep;
                    vmax = fmaxf(vmax, val);
                }
            } else {
                for (int idx = gt; idx < p.n_mels * TILE_F; idx += GROUP_THREADS) {
                    const int b = idx >> 3, f = idx & 7;
                    const float val = rise[b * TILE_F + f] + fall[(b + 1) * TILE_F + f];
                    if (t0 + f < T) {
                        me[static_cast<long long>(b) * p.mel_row_stride + f] = val;
                        vmax = fmaxf(vmax, val);
                    }
                }
            }
            if (p.mel_max != nullptr) {
                vmax = warp_max(vmax);
                if (lane == 0) atomic_max_nonneg(p.mel_max + clip, vmax);
            }
        }
    }
    cp_async_wait_all();
}

}  // namespace aegis

extern "C" int aegis_stft_fused(const aegis_stft_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "aegis_stft_fused: null params");
    AEGIS_REQUIRE(p->y && p->window && p->twiddle, "aegis_stft_fused: y/window/twiddle must be set");
    AEGIS_REQUIRE(p->hop >= 4 && p->hop <= MAX_HOP && p->hop % 4 == 0,
                  "aegis_stft_fused: hop=%d unsupported (4..512, multiple of 4)", p->hop);
    AEGIS_REQUIRE(p->n_clips >= 0 && p->n_frames >= 0 && p->n_samples >= 0, "aegis_stft_fused: negative size");
    AEGIS_REQUIRE(p->pad >= 0 && p->pad % 4 == 0, "aegis_stft_fused: pad must be a non-negative multiple of 4");
    AEGIS_REQUIRE(p->mag == nullptr || p->mag_row_stride >= p->n_frames, "aegis_stft_fused: mag_row_stride < n_frames");
    if (p->mel != nullptr) {
        AEGIS_REQUIRE(p->mel_seg_start && p->mel_rise_fall && p->n_mels > 0 && p->n_mels <= MEL_MAX,
                      "aegis_stft_fused: mel tables missing or n_mels > 128");
        AEGIS_REQUIRE(p->mel_row_stride >= p->n_frames, "aegis_stft_fused: mel_row_stride < n_frames");
    }
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    const int tiles_per_clip = (p->n_frames + TILE_F - 1) / TILE_F;
    const long long n_tiles = static_cast<long long>(tiles_per_clip) * p->n_clips;
    {
        cudaError_t e = cudaFuncSetAttribute(stft_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(StftSmem)));
        if (e != cudaSuccess) {
            set_error("aegis_stft_fused: cannot reserve %zu B shared memory: %s", sizeof(StftSmem), cudaGetErrorString(e));
            return 2;
        }
    }
    const long long ctas_needed = (n_tiles + STFT_GROUPS - 1) / STFT_GROUPS;
    const long long max_grid = sm_count();
    const int grid = static_cast<int>(ctas_needed < max_grid ? ctas_needed : max_grid);
    stft_fused_kernel<<<grid, STFT_THREADS, sizeof(StftSmem), static_cast<cudaStream_t>(stream)>>>(*p, tiles_per_clip, n_tiles);
    return check_launch("aegis_stft_fused");
}
