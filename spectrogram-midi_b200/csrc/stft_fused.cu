// K1: fused frame gather + Hann window + FP32 real FFT -> |X| (+ mel power, frame RMS).
//
// Replaces librosa.stft / feature.melspectrogram / feature.rms as called from
// aegis_engine.py:25,70 and aegis_engine_financial.py:46-50,154 (see include/aegis_b200.h).
//
// The kernel is bound by FP32 lanes, issue slots and shared-memory bandwidth together (SURVEY H3:
// ~9 FLOP per HBM byte), so the design minimises all three per frame and keeps them busy at once:
//   * one WARP transforms TWO frames at once in packed f32x2 arithmetic (rfft2048x2.cuh): half the
//     issue slots of scalar code, no cross-talk between the frames, no pre-scaling;
//   * each frame is ONE 1024-point complex transform (even/odd packing) done as 32 x 32 in
//     registers with a single exchange through shared memory, inside the warp (no CTA barrier);
//     pass 2 packs a column with its conjugate-partner column, so the real-spectrum split is done
//     in registers;
//   * window, twiddles and mel weights are read once per frame PAIR; with hop 512 the two frames
//     share 3/4 of their sample loads;
//   * warp specialisation: one persistent CTA per SM = two COMPUTE groups (4 warps each, 232
//     registers, one 8-frame tile of one clip at a time, running out of phase) + one STORE group
//     (4 warps, 40 registers) that streams finished magnitudes to HBM in librosa's [1025, T] layout
//     (one 32-byte sector = 8 frames per row) and reduces |X|^2 over the mel triangles (each bin
//     read once: rise / fall partial sums per band-edge segment) while the compute groups are
//     already transforming their next tiles.  Producer / consumer hand-over with named barriers
//     (bar.arrive / bar.sync), registers rebalanced with setmaxnreg.
// HBM traffic per frame: hop*4 B read (+ halo, L2-served) and 1025*4 B written.
#include "common.cuh"
#include "rfft2048x2.cuh"

namespace aegis {

constexpr int TILE_F = 8;
constexpr int GROUP_WARPS = TILE_F / 2;
constexpr int GROUP_THREADS = GROUP_WARPS * 32;            // 128
constexpr int STFT_GROUPS = 2;                             // compute groups
constexpr int COMPUTE_THREADS = STFT_GROUPS * GROUP_THREADS;
constexpr int STFT_THREADS = COMPUTE_THREADS + GROUP_THREADS;  // + the store group = 384
constexpr int MAX_HOP = 512;
constexpr int SAMPLES_MAX = (TILE_F - 1) * MAX_HOP + RF_N;  // 5632
constexpr int MEL_MAX = 128;
constexpr int MEL_PART = (MEL_MAX + 1) * TILE_F;            // floats per rise / fall partial array
constexpr int REGS_COMPUTE = 224, REGS_STORE = 56;
static_assert(COMPUTE_THREADS * REGS_COMPUTE + GROUP_THREADS * REGS_STORE <= STFT_THREADS * 168, "setmaxnreg trades registers inside the pool the CTA was launched with (168 per thread)");

// named barriers (0 is __syncthreads)
constexpr int BAR_FULL = 1;     // + g: compute group g arrives, store group waits   (magnitudes of a tile are complete)
constexpr int BAR_EMPTY = 3;    // + g: store group arrives, compute group g waits   (magnitudes have been consumed)
constexpr int BAR_GROUP = 5;    // + g: among the 128 threads of compute group g
constexpr int BAR_STORE = 7;    // among the 128 threads of the store group

__device__ __forceinline__ void bar_arrive(int id, int n_threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// One compute warp's shared memory: the exchange area of its transforms and the magnitude columns of its two frames
// (p2 mag[k] = frames A, B).  Separate areas: the store group reads a tile's magnitudes while the warp already
// exchanges the next tile.  The stride is 8 banks (mod 32): the 8 frames of one bin sit in 8 distinct banks.
struct alignas(16) WarpBuf {
    float x[RF_XCHG_WORDS];
    float mag[2 * RF_BINS + 22];
};
static_assert((sizeof(WarpBuf) / 4) % 32 == 8 && sizeof(WarpBuf) % 16 == 0, "consecutive warp buffers must be skewed by 8 banks");

struct StftSmem {
    WarpBuf wb[STFT_GROUPS * GROUP_WARPS];
    float samples[STFT_GROUPS][SAMPLES_MAX];
    float window[RF_N];              // 0.5 * analysis window (the split produces 2 X)
    cf32 tw1[32 * 32];               // [b][lane] W1024^{lane b}
    cf32 tw2[RF_M];                  // W2048^k
    float2 mel_rf[RF_BINS];          // (rise, fall) weight of every FFT bin
    float rise[MEL_PART], fall[MEL_PART];
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

// Tiles are dealt round-robin: tile = first + i * step.  (clip, tile-in-clip) advance without divisions.
struct TileWalk {
    int clip, tin;          // current tile = clip * tiles_per_clip + tin
    int step_clip, step_tin, tiles_per_clip;
    __device__ TileWalk(long long first, long long step, int tpc) {
        tiles_per_clip = tpc;
        clip = static_cast<int>(first / tpc);
        tin = static_cast<int>(first - static_cast<long long>(clip) * tpc);
        step_clip = static_cast<int>(step / tpc);
        step_tin = static_cast<int>(step - static_cast<long long>(step_clip) * tpc);
    }
    __device__ __forceinline__ void next(int& c, int& t) const {  // the tile after (c, t)
        c += step_clip;
        t += step_tin;
        if (t >= tiles_per_clip) {
            t -= tiles_per_clip;
            ++c;
        }
    }
    __device__ __forceinline__ void advance() { next(clip, tin); }
};

// ------------------------------------------------------------------------------------------------------------
// compute group: samples -> windowed frame pair -> two real transforms -> |X| columns in shared memory
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void compute_group(const aegis_stft_params& p, StftSmem& s, const int g, const int gt,
                                              const int tiles_per_clip, const long long n_tiles) {
    const int lane = gt & 31, wg = gt >> 5;
    const int T = p.n_frames;
    const int hop = p.hop;
    const long long N = p.n_samples;
    const int n_buf = (TILE_F - 1) * hop + RF_N;
    const bool do_fft = (p.mag != nullptr) || (p.mel != nullptr);
    float* const smp = s.samples[g];
    WarpBuf& wb = s.wb[g * GROUP_WARPS + wg];
    const long long tile_step = static_cast<long long>(gridDim.x) * STFT_GROUPS;
    const long long first = static_cast<long long>(blockIdx.x) * STFT_GROUPS + g;
    TileWalk walk(first < n_tiles ? first : 0, tile_step, tiles_per_clip);
    bool prefetched = false;  // this tile's samples were requested with cp.async during the previous tile
    bool handed_over = false; // a tile's magnitudes are with the store group

    for (long long tile = first; tile < n_tiles; tile += tile_step, walk.advance()) {
        const int clip = walk.clip;
        const int t0 = walk.tin * TILE_F;

        if (prefetched) cp_async_wait_all();
        named_barrier(BAR_GROUP + g, GROUP_THREADS);  // prefetched samples are complete and visible
        if (!prefetched) {
            const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
            fill_samples(smp, yc, static_cast<long long>(t0) * hop - p.pad, N, n_buf, gt);
            named_barrier(BAR_GROUP + g, GROUP_THREADS);
        }

        // ---- this warp's frame pair -> registers (windowed), sums of squares for the RMS
        c2 v[32];
        {
            const float* fa = smp + (2 * wg) * hop + 2 * lane;
            const float* wp = s.window + 2 * lane;
            float sa, sb;
            if (hop == 512) {
                // frame B = frame A shifted by 8 x 64 samples: lane reads 40 sample pairs instead of 64, each window
                // pair once; the sum of squares of the 24 shared pairs is shared too
                // loads are issued one block of 8 register indices ahead of their use (shared-memory latency)
                float2 xv[40], wv[32];
                p2 head = p2{0.f, 0.f}, mid = p2{0.f, 0.f}, tail = p2{0.f, 0.f};
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    xv[a] = *reinterpret_cast<const float2*>(fa + 64 * a);
                    wv[a] = *reinterpret_cast<const float2*>(wp + 64 * a);
                }
#pragma unroll
                for (int blk = 0; blk < 5; ++blk) {
                    if (blk < 4) {
#pragma unroll
                        for (int a = 8 * blk + 8; a < 8 * blk + 16; ++a) {
                            xv[a] = *reinterpret_cast<const float2*>(fa + 64 * a);
                            if (a < 32) wv[a] = *reinterpret_cast<const float2*>(wp + 64 * a);
                        }
                    }
#pragma unroll
                    for (int a = 8 * blk; a < 8 * blk + 8; ++a) {
                        const float2 x = xv[a];
                        const p2 xs = p2{x.x, x.y};
                        if (a < 8) head = pfma(xs, xs, head);
                        else if (a < 32) mid = pfma(xs, xs, mid);
                        else tail = pfma(xs, xs, tail);
                        if (a < 32) {
                            v[a].re.x = x.x * wv[a].x;
                            v[a].im.x = x.y * wv[a].y;
                        }
                        if (a >= 8) {
                            v[a - 8].re.y = x.x * wv[a - 8].x;
                            v[a - 8].im.y = x.y * wv[a - 8].y;
                        }
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
        // hides behind the transforms of this tile
        named_barrier(BAR_GROUP + g, GROUP_THREADS);
        {
            prefetched = false;
            if (tile + tile_step < n_tiles) {
                int nclip = clip, ntin = walk.tin;
                walk.next(nclip, ntin);
                const float* nyc = p.y + static_cast<long long>(nclip) * p.clip_stride;
                const long long ng0 = static_cast<long long>(ntin) * TILE_F * hop - p.pad;
                if (ng0 >= 0 && ng0 + n_buf <= N && ((ng0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(nyc) & 15) == 0)) {
                    const float* src = nyc + ng0;
                    for (int i = gt * 4; i < n_buf; i += GROUP_THREADS * 4) cp_async16(&smp[i], src + i);
                    prefetched = true;
                }
            }
            cp_async_commit();
        }
        if (!do_fft) continue;  // RMS-only call (librosa.feature.rms)

        // ---- two real 2048-point transforms, inside the warp; the exchange in two rounds (re, im)
        rfft_pass1(lane, v, s.tw1);
        rfft_xstore<false>(lane, v, wb.x);
        __syncwarp();
        p2 nre[32];
        rfft_xload(lane, wb.x, nre);
        __syncwarp();
        rfft_xstore<true>(lane, v, wb.x);
        __syncwarp();
        {
            p2 nim[32];
            rfft_xload(lane, wb.x, nim);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = c2{nre[j], nim[j]};
        }
        fft32(v);
        // the store group must have consumed the previous tile's magnitudes before they are overwritten
        if (handed_over) named_barrier(BAR_EMPTY + g, 2 * GROUP_THREADS);
        {
            float* const mcol = wb.mag + (lane >> 4);  // p2 mag[k]: word 2k + frame
            rfft_split_emit(lane, v, s.tw2, [&](int k, float pw) { mcol[2 * k] = sqrt_approx(pw); });
        }
        bar_arrive(BAR_FULL + g, 2 * GROUP_THREADS);
        handed_over = true;
    }
    cp_async_wait_all();
    if (handed_over) named_barrier(BAR_EMPTY + g, 2 * GROUP_THREADS);  // pairs with the store group's last arrive
}

// ------------------------------------------------------------------------------------------------------------
// store group: |X| columns of a finished tile -> HBM, mel projection
// ------------------------------------------------------------------------------------------------------------
// |X| only (no mel projection requested): one float4 (4 frames) per lane, two lanes per spectrogram row
__device__ __forceinline__ void store_mag_only(const aegis_stft_params& p, const WarpBuf* gbuf, const int gt, const int clip, const int t0) {
    const int lane = gt & 31, wg = gt >> 5;
    const int T = p.n_frames;
    float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0;
    const bool vec_store = ((p.mag_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 15) == 0);
    if (vec_store) {
        const int h = gt & 1, kq = gt >> 1;
        const p2* m0p = reinterpret_cast<const p2*>(gbuf[2 * h].mag);
        const p2* m1p = reinterpret_cast<const p2*>(gbuf[2 * h + 1].mag);
        float* dst = mo + static_cast<long long>(kq) * p.mag_row_stride + 4 * h;
        const long long dstep = 64LL * p.mag_row_stride;
        for (int k = kq; k < RF_BINS; k += 64) {
            const p2 m0 = m0p[k], m1 = m1p[k];
            if (t0 + 4 * h + 3 < T) {
                *reinterpret_cast<float4*>(dst) = make_float4(m0.x, m0.y, m1.x, m1.y);
            } else {
                const float st[4] = {m0.x, m0.y, m1.x, m1.y};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (t0 + 4 * h + j < T) dst[j] = st[j];
            }
            dst += dstep;
        }
    } else {          // each warp store = 4 rows x 8 frames
        const int f = lane & 7;
        if (t0 + f < T) {
            const float* col = gbuf[f >> 1].mag + (f & 1);
            for (int k = wg * 4 + (lane >> 3); k < RF_BINS; k += GROUP_WARPS * 4)
                mo[static_cast<long long>(k) * p.mag_row_stride + f] = col[2 * k];
        }
    }
}

// One pass over the tile's magnitudes: every (bin, frame pair) is read from shared memory ONCE, written to HBM and
// accumulated into the rise / fall partial sums of its mel segment.  lane = (segment slot, frame pair): the four
// lanes of a slot write one 32-byte sector of a spectrogram row.
// MODE 0: no |X| output; 1: 8-byte stores (full tile, even row stride, aligned base); 2: guarded scalar stores.
template <int MODE>
__device__ __forceinline__ void store_mel_pass(const aegis_stft_params& p, StftSmem& s, const WarpBuf* gbuf, const int gt,
                                               float* __restrict__ mo, const int frames_left) {
    const int lane = gt & 31, wg = gt >> 5;
    const int b4 = lane & 3;
    const p2* col = reinterpret_cast<const p2*>(gbuf[b4].mag);
    const long long rs = p.mag_row_stride;
    float* const dcol = mo + 2 * b4;
    const bool ok0 = 2 * b4 < frames_left, ok1 = 2 * b4 + 1 < frames_left;
    auto put = [&](int k, p2 m) {
        if (MODE == 1) {
            *reinterpret_cast<p2*>(dcol + k * rs) = m;
        } else if (MODE == 2) {
            if (ok0) dcol[k * rs] = m.x;
            if (ok1) dcol[k * rs + 1] = m.y;
        }
    };
    auto acc = [&](p2 m, float2 w, p2& r, p2& fl) {
        const p2 pw = m * m;
        r = p2{fmaf(pw.x, w.x, r.x), fmaf(pw.y, w.x, r.y)};
        fl = p2{fmaf(pw.x, w.y, fl.x), fmaf(pw.y, w.y, fl.y)};
    };
    for (int j = wg * 8 + (lane >> 2); j <= p.n_mels; j += GROUP_WARPS * 8) {
        int k = s.mel_seg[j];
        const int k1 = s.mel_seg[j + 1];
        p2 r = p2{0.f, 0.f}, fl = p2{0.f, 0.f};
        for (; k + 4 <= k1; k += 4) {   // all eight loads first: one shared-memory latency per four bins
            const p2 m0 = col[k], m1 = col[k + 1], m2 = col[k + 2], m3 = col[k + 3];
            const float2 w0 = s.mel_rf[k], w1 = s.mel_rf[k + 1], w2 = s.mel_rf[k + 2], w3 = s.mel_rf[k + 3];
            put(k, m0);
            put(k + 1, m1);
            put(k + 2, m2);
            put(k + 3, m3);
            acc(m0, w0, r, fl);
            acc(m1, w1, r, fl);
            acc(m2, w2, r, fl);
            acc(m3, w3, r, fl);
        }
        if (k < k1) {                   // up to three bins left: predicated, still all loads first
            const bool h1 = k + 1 < k1, h2 = k + 2 < k1;
            const p2 m0 = col[k], m1 = col[h1 ? k + 1 : k], m2 = col[h2 ? k + 2 : k];
            const float2 w0 = s.mel_rf[k];
            float2 w1 = s.mel_rf[h1 ? k + 1 : k], w2 = s.mel_rf[h2 ? k + 2 : k];
            if (!h1) w1 = make_float2(0.f, 0.f);
            if (!h2) w2 = make_float2(0.f, 0.f);
            put(k, m0);
            if (h1) put(k + 1, m1);
            if (h2) put(k + 2, m2);
            acc(m0, w0, r, fl);
            acc(m1, w1, r, fl);
            acc(m2, w2, r, fl);
        }
        *reinterpret_cast<p2*>(&s.rise[j * TILE_F + 2 * b4]) = r;
        *reinterpret_cast<p2*>(&s.fall[j * TILE_F + 2 * b4]) = fl;
    }
}

__device__ __forceinline__ void store_tile(const aegis_stft_params& p, StftSmem& s, const int g, const int gt, const int clip, const int t0) {
    const int lane = gt & 31;
    const int T = p.n_frames;
    const WarpBuf* gbuf = &s.wb[g * GROUP_WARPS];
    if (p.mel == nullptr) {
        store_mag_only(p, gbuf, gt, clip, t0);
        bar_arrive(BAR_EMPTY + g, 2 * GROUP_THREADS);
        return;
    }
    const bool full_tile = (t0 + TILE_F <= T);
    // mel[b] = sum_{k in seg b} rise[k] |X_k|^2 + sum_{k in seg b+1} fall[k] |X_k|^2
    if (p.mag == nullptr) {
        store_mel_pass<0>(p, s, gbuf, gt, nullptr, 0);
    } else {
        float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0;
        if (full_tile && ((p.mag_row_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 7) == 0))
            store_mel_pass<1>(p, s, gbuf, gt, mo, TILE_F);
        else
            store_mel_pass<2>(p, s, gbuf, gt, mo, T - t0);
    }
    bar_arrive(BAR_EMPTY + g, 2 * GROUP_THREADS);   // the magnitudes are consumed: the compute group may overwrite them
    named_barrier(BAR_STORE, GROUP_THREADS);        // partial sums complete
    float* __restrict__ me = p.mel + static_cast<long long>(clip) * p.mel_clip_stride + t0;
    float vmax = 0.f;
    if (full_tile && p.n_mels == MEL_MAX) {
        const int f = gt & 7;
        float* dst = me + static_cast<long long>(gt >> 3) * p.mel_row_stride + f;
        const long long dstep = 16LL * p.mel_row_stride;
#pragma unroll
        for (int it = 0; it < MEL_MAX / 16; ++it) {
            const float val = s.rise[gt + 128 * it] + s.fall[gt + 128 * it + TILE_F];
            *dst = val;
            dst += dstep;
            vmax = fmaxf(vmax, val);
        }
    } else {
        for (int idx = gt; idx < p.n_mels * TILE_F; idx += GROUP_THREADS) {
            const int b = idx >> 3, f = idx & 7;
            const float val = s.rise[b * TILE_F + f] + s.fall[(b + 1) * TILE_F + f];
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
    named_barrier(BAR_STORE, GROUP_THREADS);        // partial sums are free for the next tile
}

__device__ __forceinline__ void store_group(const aegis_stft_params& p, StftSmem& s, const int gt,
                                            const int tiles_per_clip, const long long n_tiles) {
    if (p.mag == nullptr && p.mel == nullptr) return;
    const long long tile_step = static_cast<long long>(gridDim.x) * STFT_GROUPS;
    const long long first = static_cast<long long>(blockIdx.x) * STFT_GROUPS;
    TileWalk w0(first < n_tiles ? first : 0, tile_step, tiles_per_clip);
    TileWalk w1(first + 1 < n_tiles ? first + 1 : 0, tile_step, tiles_per_clip);
    for (long long tile = first; tile < n_tiles; tile += tile_step) {  // the two compute groups finish tiles alternately
        named_barrier(BAR_FULL + 0, 2 * GROUP_THREADS);
        store_tile(p, s, 0, gt, w0.clip, w0.tin * TILE_F);
        w0.advance();
        if (tile + 1 < n_tiles) {
            named_barrier(BAR_FULL + 1, 2 * GROUP_THREADS);
            store_tile(p, s, 1, gt, w1.clip, w1.tin * TILE_F);
            w1.advance();
        }
    }
}

__global__ void __launch_bounds__(STFT_THREADS, 1)
stft_fused_kernel(const aegis_stft_params p, const int tiles_per_clip, const long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x;
    {   // constant tables (once per CTA)
        const cf32* tab = reinterpret_cast<const cf32*>(p.twiddle);
        for (int i = tid; i < RF_N; i += STFT_THREADS) s.window[i] = 0.5f * p.window[i];
        for (int i = tid; i < 32 * 32; i += STFT_THREADS) s.tw1[i] = tab[(2 * (i & 31) * (i >> 5)) & (RF_N - 1)];
        for (int i = tid; i < RF_M; i += STFT_THREADS) s.tw2[i] = tab[i];
        if (p.mel != nullptr) {
            for (int i = tid; i < RF_BINS; i += STFT_THREADS) s.mel_rf[i] = reinterpret_cast<const float2*>(p.mel_rise_fall)[i];
            for (int i = tid; i < p.n_mels + 2; i += STFT_THREADS) s.mel_seg[i] = p.mel_seg_start[i];
        }
    }
    __syncthreads();
    if (tid < COMPUTE_THREADS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
        compute_group(p, s, tid >> 7, tid & 127, tiles_per_clip, n_tiles);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_STORE));
        store_group(p, s, tid & 127, tiles_per_clip, n_tiles);
    }
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
