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
//   * warp specialisation: one persistent CTA of 512 threads per SM = two COMPUTE groups (4 warps each, 200
//     registers, one 8-frame tile of one clip at a time) + two STORE groups (4 warps each, 56 registers), one per
//     compute group.  A compute group leaves |X| of its tile in a shared-memory stage laid out as a TMA box
//     (CU_TENSOR_MAP_SWIZZLE_32B) and goes on with its next tile; its store group has the TMA unit write the stage
//     into librosa's [1025, T] layout (cp.async.bulk.tensor, one thread, no LSU traffic) and meanwhile reduces |X|^2
//     over the mel triangles (each bin read once: rise / fall partial sums per band-edge segment).  Producer /
//     consumer hand-over with named barriers (bar.arrive / bar.sync), registers rebalanced with setmaxnreg.
//   * the hot loop is ~2800 straight-line instructions (45 KB) against a 32 KB L1.5 instruction cache: the two
//     compute groups are kept in step (one barrier per tile) so that they fetch the same code at the same time
//     (ncu: stall_no_instruction was the top stall when they drifted apart; 4.3 ms without the barrier, 3.7 with).
// What bounds it now (ncu, profiles/): the FP32 pipe during the transform phases (both compute warps of an SM
// sub-partition in step), shared-memory latency in the store groups.
// HBM traffic per frame: hop*4 B read (+ halo, L2-served) and 1025*4 B written.
#include <cstddef>
#include <cstring>
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include "rfft2048x2.cuh"

namespace aegis {

constexpr int TILE_F = 8;
constexpr int GROUP_WARPS = TILE_F / 2;
constexpr int GROUP_THREADS = GROUP_WARPS * 32;            // 128
constexpr int STFT_GROUPS = 2;                             // compute groups
constexpr int COMPUTE_THREADS = STFT_GROUPS * GROUP_THREADS;
constexpr int STFT_THREADS = 2 * COMPUTE_THREADS;             // + one store group per compute group = 512
constexpr int MAX_HOP = 512;
constexpr int SAMPLES_MAX = (TILE_F - 1) * MAX_HOP + RF_N;  // 5632
constexpr int MEL_MAX = 128;
constexpr int MEL_PART = (MEL_MAX + 1) * TILE_F;            // floats per rise / fall partial array
constexpr int REGS_LAUNCH = 65536 / STFT_THREADS;            // 128: what __launch_bounds__(512, 1) gives every thread
constexpr int REGS_COMPUTE = 200, REGS_STORE = 2 * REGS_LAUNCH - REGS_COMPUTE;   // 200 / 56 (A/B timed: 176..216 within 5 %)
static_assert(COMPUTE_THREADS * (REGS_COMPUTE + REGS_STORE) <= STFT_THREADS * REGS_LAUNCH,
              "setmaxnreg trades registers inside the pool the CTA was launched with");

// named barriers (0 is __syncthreads)
constexpr int BAR_FULL = 1;     // + g: compute group g arrives, store group waits   (magnitudes of a tile are complete)
constexpr int BAR_EMPTY = 3;    // + g: store group arrives, compute group g waits   (magnitudes have been consumed)
constexpr int BAR_GROUP = 5;    // + g: among the 128 threads of compute group g
constexpr int BAR_LOCKSTEP = 7; // among the 256 threads of both compute groups
constexpr int BAR_STORE = 8;    // + g: among the 128 threads of store group g

__device__ __forceinline__ void bar_arrive(int id, int n_threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// Magnitude stage of one compute group: |X| of the tile as [1025 bins][8 frames], 32 bytes per bin, in the layout
// of a TMA box with CU_TENSOR_MAP_SWIZZLE_32B: the two 16-byte halves of a row swap when bit 7 of the byte address
// (= bit 2 of the bin index, the stage being 256-byte aligned) is set.  The compute warps scatter single floats into it
// (2-way bank conflicts instead of 4-way for the plain layout), the store group reads whole 16-byte halves, and one
// thread hands the stage to the TMA unit, which writes the [1025, 8] block into librosa's [1025, T] layout by itself.
constexpr int STAGE_WORDS = RF_BINS * TILE_F;            // 8200
constexpr int STAGE_BOX_ROWS = 256;                      // TMA box = 8 frames x 256 bins; 5 boxes (the last one clipped)
__host__ __device__ constexpr int stage_word(int k, int f) { return 8 * k + 4 * ((f >> 2) ^ ((k >> 2) & 1)) + (f & 3); }

struct StftSmem {
    alignas(256) float stage[STFT_GROUPS][STAGE_WORDS + 56];   // 33024 B each: keeps the second stage 256-byte aligned
    float xchg[STFT_GROUPS * GROUP_WARPS][RF_XCHG_WORDS];      // one exchange area per compute warp
    float samples[STFT_GROUPS][SAMPLES_MAX];
    float window[RF_N];              // 0.5 * analysis window (the split produces 2 X)
    cf32 tw1[32 * 32];               // [b][lane] W1024^{lane b}
    tw4 tw2p[16 * 16];               // twiddles of the packed split, [iteration c][column pair q]
    cf32 tw512;                      // W2048^512
    float2 mel_rf[RF_BINS];          // (rise, fall) weight of every FFT bin
    alignas(16) float rise[STFT_GROUPS][MEL_PART];   // per store group
    alignas(16) float fall[STFT_GROUPS][MEL_PART];
    int mel_seg[MEL_MAX + 2];        // first bin of every segment between mel band edges
};
static_assert(sizeof(StftSmem) <= 227 * 1024, "shared memory budget of one SM");
static_assert(((STAGE_WORDS + 56) * 4) % 256 == 0, "stage alignment");
// the clipped last TMA box still addresses 256 rows of shared memory: they must lie inside the allocation
static_assert(offsetof(StftSmem, stage) + sizeof(float) * ((STFT_GROUPS - 1) * (STAGE_WORDS + 56) + (4 * STAGE_BOX_ROWS + STAGE_BOX_ROWS) * TILE_F) <= sizeof(StftSmem), "TMA box overruns shared memory");

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
    float* const xbuf = s.xchg[g * GROUP_WARPS + wg];
    const long long tile_step = static_cast<long long>(gridDim.x) * STFT_GROUPS;
    const long long first = static_cast<long long>(blockIdx.x) * STFT_GROUPS + g;
    TileWalk walk(first < n_tiles ? first : 0, tile_step, tiles_per_clip);
    bool prefetched = false;  // this tile's samples were requested with cp.async during the previous tile
    bool handed_over = false; // a tile's magnitudes are with the store group

    // The two compute groups start every tile together (one more barrier): they then run the same instructions at
    // about the same time, which halves the instruction-cache traffic.  Group 0 never has fewer tiles than group 1;
    // a group without a tile in the last round still joins the barrier.
    for (long long tile = first; tile - g < n_tiles; tile += tile_step, walk.advance()) {
        named_barrier(BAR_LOCKSTEP, COMPUTE_THREADS);
        if (tile >= n_tiles) break;
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
        fft32(v);
        rfft_twiddle1(lane, v, s.tw1);
        rfft_xstore<false>(lane, v, xbuf);
        __syncwarp();
        {
            p2 nre[32], nim[32];
            rfft_xload(lane, xbuf, nre);
            __syncwarp();
            rfft_xstore<true>(lane, v, xbuf);
            __syncwarp();
            rfft_xload(lane, xbuf, nim);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = c2{nre[j], nim[j]};
        }
        fft32(v);
        // the store group must have consumed the previous tile's magnitudes before they are overwritten
        if (handed_over) named_barrier(BAR_EMPTY + g, 2 * GROUP_THREADS);
        {   // this thread's frame is f = 2 wg + (lane >> 4); bins k = q + 32c and 1024 - k (rfft_split_emit): the swizzle
            // bit (bin >> 2) & 1 is the same for all of a thread's "k" bins and for all of its "1024 - k" bins
            const int f = 2 * wg + (lane >> 4), q = lane & 15;
            float* const st = s.stage[g];
            const int off_k = stage_word(q, f) - 8 * q, off_n = stage_word(RF_M - q, f) - 8 * (RF_M - q);
            rfft_split_emit(lane, v, s.tw2p, s.tw512, [&](int k, float pw, bool is_k) { st[8 * k + (is_k ? off_k : off_n)] = sqrt_approx(pw); });
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the TMA unit (async proxy) reads these writes
        bar_arrive(BAR_FULL + g, 2 * GROUP_THREADS);
        handed_over = true;
    }
    cp_async_wait_all();
    if (handed_over) named_barrier(BAR_EMPTY + g, 2 * GROUP_THREADS);  // pairs with the store group's last arrive
}

// ------------------------------------------------------------------------------------------------------------
// store group: |X| columns of a finished tile -> HBM, mel projection
// ------------------------------------------------------------------------------------------------------------
// |X| of a finished tile -> HBM without the TMA unit (row pitch or base address not 16-byte aligned, or no tensor map):
// one 16-byte half row (4 frames) per lane, two lanes per spectrogram row
__device__ __forceinline__ void store_mag_lsu(const aegis_stft_params& p, const float* st, const int gt, const int clip, const int t0) {
    const int T = p.n_frames;
    float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0;
    const bool vec_store = ((p.mag_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 15) == 0);
    const int h = gt & 1;
    float* dst = mo + static_cast<long long>(gt >> 1) * p.mag_row_stride + 4 * h;
    const long long dstep = 64LL * p.mag_row_stride;
    for (int k = gt >> 1; k < RF_BINS; k += 64) {
        const float4 m = *reinterpret_cast<const float4*>(st + stage_word(k, 4 * h));
        if (vec_store && t0 + 4 * h + 3 < T) {
            *reinterpret_cast<float4*>(dst) = m;
        } else {
            const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (t0 + 4 * h + j < T) dst[j] = mv[j];
        }
        dst += dstep;
    }
}

// rise / fall partial sums of the mel projection: lane = (segment slot, 4-frame half); every (bin, half) is read from
// the stage once with one 16-byte load
__device__ __forceinline__ void mel_partials(const aegis_stft_params& p, const StftSmem& s, float* rise, float* fall, const float* st, const int gt) {
    const int h = gt & 1;
    // neighbouring lanes take neighbouring segments (similar lengths: little divergence inside a warp)
    const int slot = gt >> 1;
    struct q4 { p2 a, b; };   // four frames as two packed pairs
    auto acc = [&](float4 m, float2 w, q4& r, q4& fl) {
        const p2 pa = p2{m.x, m.y} * p2{m.x, m.y}, pb = p2{m.z, m.w} * p2{m.z, m.w};
        r.a = pfma(pa, w.x, r.a);
        r.b = pfma(pb, w.x, r.b);
        fl.a = pfma(pa, w.y, fl.a);
        fl.b = pfma(pb, w.y, fl.b);
    };
    auto ld = [&](int k) { return *reinterpret_cast<const float4*>(st + stage_word(k, 4 * h)); };
    constexpr int NB = 8;   // bins per round: all loads of a round are issued before the first use (4, 6, 8 time alike)
    for (int j = slot; j <= p.n_mels; j += GROUP_THREADS / 2) {
        int k = s.mel_seg[j];
        const int k1 = s.mel_seg[j + 1];
        q4 r = q4{p2{0.f, 0.f}, p2{0.f, 0.f}}, fl = r;
        // The four segment slots of a quarter-warp read four different stage rows with one 16-byte load each: they
        // are conflict-free iff the rows differ modulo 4 (a row is 32 B).  Each slot therefore walks the bins of a
        // round in rotated order, so that at load i its row is congruent to slot + i (ncu: 1.7x excess wavefronts
        // before).  Bins past the end of the segment are predicated off (no wavefront) and get zero weights.
        const int rot = (slot - k) & 3;
        for (; k < k1; k += NB) {
            float4 m[NB];
            float2 w[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int kk = k + ((i + rot) & (NB - 1));
                m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                w[i] = make_float2(0.f, 0.f);
                if (kk < k1) {
                    m[i] = ld(kk);
                    w[i] = s.mel_rf[kk];
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) acc(m[i], w[i], r, fl);
        }
        *reinterpret_cast<float4*>(&rise[j * TILE_F + 4 * h]) = make_float4(r.a.x, r.a.y, r.b.x, r.b.y);
        *reinterpret_cast<float4*>(&fall[j * TILE_F + 4 * h]) = make_float4(fl.a.x, fl.a.y, fl.b.x, fl.b.y);
    }
}

__device__ __forceinline__ void store_tile(const aegis_stft_params& p, const CUtensorMap* tmap, const bool use_tma, StftSmem& s,
                                           const int g, const int gt, const int clip, const int t0) {
    float* const rise = s.rise[g];
    float* const fall = s.fall[g];
    const int lane = gt & 31;
    const int T = p.n_frames;
    const float* st = s.stage[g];
    if (p.mag != nullptr) {
        if (use_tma) {
            if (gt == 0) {   // five boxes of 256 bins x 8 frames; frames >= T and bins > 1024 are clipped by the tensor map
                const unsigned src = static_cast<unsigned>(__cvta_generic_to_shared(st));
#pragma unroll
                for (int bx = 0; bx < 5; ++bx)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                                     reinterpret_cast<unsigned long long>(tmap)),
                                 "r"(src + bx * STAGE_BOX_ROWS * TILE_F * 4), "r"(t0), "r"(bx * STAGE_BOX_ROWS), "r"(clip)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            store_mag_lsu(p, st, gt, clip, t0);
        }
    }
    if (p.mel != nullptr) mel_partials(p, s, rise, fall, st, gt);
    if (p.mag != nullptr && use_tma && gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // stage has been read
    bar_arrive(BAR_EMPTY + g, 2 * GROUP_THREADS);   // the magnitudes are consumed: the compute group may overwrite them
    if (p.mel == nullptr) return;
    named_barrier(BAR_STORE + g, GROUP_THREADS);        // partial sums complete (and every thread is done with the other buffer)
    const bool full_tile = (t0 + TILE_F <= T);
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
    named_barrier(BAR_STORE + g, GROUP_THREADS);        // partial sums are free for the next tile
}

__device__ __forceinline__ void store_group(const aegis_stft_params& p, const CUtensorMap* tmap, const bool use_tma, StftSmem& s,
                                            const int g, const int gt, const int tiles_per_clip, const long long n_tiles) {
    if (p.mag == nullptr && p.mel == nullptr) return;
    const long long tile_step = static_cast<long long>(gridDim.x) * STFT_GROUPS;
    const long long first = static_cast<long long>(blockIdx.x) * STFT_GROUPS + g;
    TileWalk w(first < n_tiles ? first : 0, tile_step, tiles_per_clip);
    for (long long tile = first; tile < n_tiles; tile += tile_step, w.advance()) {
        named_barrier(BAR_FULL + g, 2 * GROUP_THREADS);
        store_tile(p, tmap, use_tma, s, g, gt, w.clip, w.tin * TILE_F);
    }
    if (use_tma && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every bulk store has landed before the CTA retires
}

__global__ void __launch_bounds__(STFT_THREADS, 1)
stft_fused_kernel(const aegis_stft_params p, const __grid_constant__ CUtensorMap mag_map, const int use_tma,
                  const int tiles_per_clip, const long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x;
    {   // constant tables (once per CTA)
        const cf32* tab = reinterpret_cast<const cf32*>(p.twiddle);
        for (int i = tid; i < RF_N; i += STFT_THREADS) s.window[i] = 0.5f * p.window[i];
        for (int i = tid; i < 32 * 32; i += STFT_THREADS) s.tw1[i] = tab[(2 * (i & 31) * (i >> 5)) & (RF_N - 1)];
        for (int i = tid; i < 16 * 16; i += STFT_THREADS) {
            const cf32 lo = tab[rfft_split_klo(i & 15, i >> 4)], hi = tab[rfft_split_khi(i & 15, i >> 4)];
            s.tw2p[i] = tw4{lo.x, hi.x, lo.y, hi.y};
        }
        if (tid == 0) s.tw512 = tab[RF_M / 2];
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
        store_group(p, &mag_map, use_tma != 0, s, (tid - COMPUTE_THREADS) >> 7, tid & 127, tiles_per_clip, n_tiles);
    }
}

}  // namespace aegis

namespace aegis {
// Tensor map of the |X| output, [n_clips][1025][T] with the caller's row / clip pitch, boxes of 8 frames x 256 bins.
// Returns false when the layout cannot be described (pitches or base not 16-byte aligned, driver entry point missing):
// the kernel then stores through the LSU path.
static bool make_mag_map(const aegis_stft_params* p, CUtensorMap* map) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const encode_fn encode = []() -> encode_fn {   // resolved once (thread-safe static initialisation)
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<encode_fn>(fn);
        return nullptr;
    }();
    if (encode == nullptr || p->mag == nullptr) return false;
    if ((reinterpret_cast<uintptr_t>(p->mag) & 15) || (p->mag_row_stride & 3) || (p->mag_clip_stride & 3)) return false;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(p->n_frames), static_cast<cuuint64_t>(AEGIS_N_BINS), static_cast<cuuint64_t>(p->n_clips)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(p->mag_row_stride) * 4, static_cast<cuuint64_t>(p->mag_clip_stride) * 4};
    const cuuint32_t box[3] = {TILE_F, STAGE_BOX_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p->mag, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
    alignas(64) CUtensorMap mag_map;
    memset(&mag_map, 0, sizeof(mag_map));
    const int use_tma = make_mag_map(p, &mag_map) ? 1 : 0;
    stft_fused_kernel<<<grid, STFT_THREADS, sizeof(StftSmem), static_cast<cudaStream_t>(stream)>>>(*p, mag_map, use_tma, tiles_per_clip, n_tiles);
    return check_launch("aegis_stft_fused");
}
