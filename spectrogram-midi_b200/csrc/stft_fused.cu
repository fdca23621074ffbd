// K1: fused frame gather + Hann window + FP32 2048-point FFT -> |X| (+ mel power, frame RMS).
//
// Replaces librosa.stft / feature.melspectrogram / feature.rms as called from
// aegis_engine.py:25,70 and aegis_engine_financial.py:46-50,154 (see include/aegis_b200.h).
//
// Work unit ("tile") = 8 consecutive frames of one clip.  Persistent CTAs of 256 threads (2 FFT
// groups x 128 threads), two resident per SM so one CTA's load / store phases overlap the other's
// transforms.  Tiles are walked with stride gridDim.x: neighbouring CTAs work on neighbouring tiles
// of the same clip at the same time (halo samples and partially written output sectors meet in L2).
// Per tile:
//   1. (7*hop + 2048) samples -> shared, float4 loads, zeros outside the clip (centre padding)
//   2. two rounds; in each, a group packs two real frames (re = frame a, im = frame b) into one
//      complex transform.  Each frame is first scaled by an exact power of two ~ 1/||w x|| so the
//      rounding error of the shared transform stays relative to the frame's OWN norm (a quiet frame
//      next to an onset keeps its accuracy; an all-zero frame yields exact zeros).  The sums of
//      squares needed for that scale and for the RMS output come from the values pass 1 loads anyway.
//      The three radix passes exchange through ONE shared buffer per group (in place).
//   3. X_a[k], X_b[k] are separated from Z[k], conj(Z[N-k]); magnitudes go to an [1025][8] stage
//   4. the stage is written out in librosa's [1025, T] layout (float4 per half row when the row
//      stride allows, else 8 scalar lanes per row); the mel projection reads each stage bin once (every
//      FFT bin lies in at most two triangles: rise/fall partial sums per band-edge segment).
// HBM traffic per frame: hop*4 B read (+ halo, L2-served) and 1025*4 B written.
#include "common.cuh"
#include "fft2048.cuh"

namespace aegis {

constexpr int TILE_F = 8;
constexpr int STFT_THREADS = 256;
constexpr int STFT_GROUPS = STFT_THREADS / FFT_THREADS;  // 2
constexpr int MAX_HOP = 512;
constexpr int SAMPLES_MAX = (TILE_F - 1) * MAX_HOP + FFT_N;  // 5632
constexpr int STAGE_PITCH = TILE_F + 1;
constexpr int MEL_MAX = 128;

struct StftSmem {
    float samples[SAMPLES_MAX];
    float window[FFT_N];
    float part[STFT_GROUPS][4][4];   // per warp: sum x_a^2, x_b^2, (w x_a)^2, (w x_b)^2 (float4 rows: keep 16 B aligned)
    cf buf[STFT_GROUPS][BUFA_SIZE];
    float stage[AEGIS_N_BINS * STAGE_PITCH];
    // mel tables live in shared memory: with ~220 KB of the SM carved out only a few KB of L1 are left, so
    // weights read through L1 would thrash
    float2 mel_rf[AEGIS_N_BINS];     // (rise, fall) weight of every FFT bin
    int mel_seg[MEL_MAX + 2];        // first bin of every segment between mel band edges
};
static_assert(2 * (MEL_MAX + 1) * TILE_F * sizeof(float) <= sizeof(cf) * BUFA_SIZE, "mel partial sums alias buf[0]");
static_assert((SAMPLES_MAX * 4) % 16 == 0 && (FFT_N * 4) % 16 == 0, "part[] must stay 16-byte aligned");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int pack_exponent(float windowed_energy) {
    return (windowed_energy > 0.f) ? (ilogbf(windowed_energy) >> 1) : 0;
}

__global__ void __launch_bounds__(STFT_THREADS, 2)
stft_fused_kernel(const aegis_stft_params p, const int tiles_per_clip, const long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = tid >> 7, lt = tid & 127, lane = tid & 31, warp = tid >> 5, gw = warp & 3;
    const int T = p.n_frames;
    const int hop = p.hop;
    const long long N = p.n_samples;

    for (int i = tid; i < FFT_N; i += STFT_THREADS) s.window[i] = p.window[i];
    if (p.mel != nullptr) {
        for (int i = tid; i < AEGIS_N_BINS; i += STFT_THREADS) s.mel_rf[i] = reinterpret_cast<const float2*>(p.mel_rise_fall)[i];
        for (int i = tid; i < p.n_mels + 2; i += STFT_THREADS) s.mel_seg[i] = p.mel_seg_start[i];
    }
    FftTwiddles tw;
    fft2048_load_twiddles(lt, reinterpret_cast<const cf*>(p.twiddle), tw);
    const int n_buf = (TILE_F - 1) * hop + FFT_N;
    const bool do_fft = (p.mag != nullptr) || (p.mel != nullptr);
    cf* const buf = s.buf[g];
    constexpr int N_ROUNDS = TILE_F / (2 * STFT_GROUPS);
    bool prefetched = false;  // the samples of this tile were requested with cp.async during the previous one

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int clip = static_cast<int>(tile / tiles_per_clip);
        const int t0 = static_cast<int>(tile - static_cast<long long>(clip) * tiles_per_clip) * TILE_F;
        const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
        const long long g0 = static_cast<long long>(t0) * hop - p.pad;

        __syncthreads();  // everybody is done with the previous tile's stage (and samples, if not prefetched)
        if (prefetched) {
            cp_async_wait_all();
        } else {
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(yc) & 15) == 0) && ((g0 & 3) == 0);
            for (int i = tid * 4; i < n_buf; i += STFT_THREADS * 4) {
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
                *reinterpret_cast<float4*>(&s.samples[i]) = v;
            }
        }
        __syncthreads();

#pragma unroll 1
        for (int round = 0; round < N_ROUNDS; ++round) {
            const int fa_idx = 2 * (round * STFT_GROUPS + g);  // this group's frame pair (fa_idx, fa_idx + 1)
            const float* fa = s.samples + fa_idx * hop;
            const float* fb = fa + hop;
            cf v[16];
            float ea2 = 0.f, eb2 = 0.f, wa2 = 0.f, wb2 = 0.f;
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                const int n = lt + 128 * a;
                const float w = s.window[n], xa = fa[n], xb = fb[n];
                const float pa = xa * w, pb = xb * w;
                ea2 = fmaf(xa, xa, ea2);
                eb2 = fmaf(xb, xb, eb2);
                wa2 = fmaf(pa, pa, wa2);
                wb2 = fmaf(pb, pb, wb2);
                v[a] = cf{pa, pb};
            }
            ea2 = warp_sum(ea2);
            eb2 = warp_sum(eb2);
            wa2 = warp_sum(wa2);
            wb2 = warp_sum(wb2);
            if (lane == 0) *reinterpret_cast<float4*>(s.part[g][gw]) = make_float4(ea2, eb2, wa2, wb2);
            if (round == N_ROUNDS - 1) {
                // last read of this tile's samples is behind every thread: request the next tile's samples now,
                // so their DRAM latency hides behind the rest of this tile (passes 2-3, split, stores, mel)
                __syncthreads();
                prefetched = false;
                const long long nt = tile + gridDim.x;
                if (nt < n_tiles) {
                    const int nclip = static_cast<int>(nt / tiles_per_clip);
                    const int nt0 = static_cast<int>(nt - static_cast<long long>(nclip) * tiles_per_clip) * TILE_F;
                    const float* nyc = p.y + static_cast<long long>(nclip) * p.clip_stride;
                    const long long ng0 = static_cast<long long>(nt0) * hop - p.pad;
                    if (ng0 >= 0 && ng0 + n_buf <= N && ((ng0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(nyc) & 15) == 0)) {
                        for (int i = tid * 4; i < n_buf; i += STFT_THREADS * 4) cp_async16(&s.samples[i], nyc + ng0 + i);
                        prefetched = true;
                    }
                }
                cp_async_commit();
            } else {
                named_barrier(1 + g, FFT_THREADS);
            }
            float4 tot = *reinterpret_cast<const float4*>(s.part[g][0]);
#pragma unroll
            for (int w4 = 1; w4 < 4; ++w4) {
                const float4 q = *reinterpret_cast<const float4*>(s.part[g][w4]);
                tot.x += q.x; tot.y += q.y; tot.z += q.z; tot.w += q.w;
            }
            if (p.rms != nullptr && lt < 2 && t0 + fa_idx + lt < T) {
                p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t0 + fa_idx + lt] =
                    sqrtf((lt == 0 ? tot.x : tot.y) * (1.0f / FFT_N));
            }
            if (!do_fft) {  // RMS-only call (librosa.feature.rms)
                named_barrier(1 + g, FFT_THREADS);  // s.part is rewritten next round
                continue;
            }
            const int ea = pack_exponent(tot.z), eb = pack_exponent(tot.w);
            {
                const float sa = ldexpf(1.0f, -ea), sb = ldexpf(1.0f, -eb);
#pragma unroll
                for (int a = 0; a < 16; ++a) v[a] = cf{v[a].x * sa, v[a].y * sb};
            }
            fft2048_pass1(lt, v, tw, buf);
            named_barrier(1 + g, FFT_THREADS);
            fft2048_pass2_load(lt, buf, v);
            named_barrier(1 + g, FFT_THREADS);
            fft2048_pass2_store(lt, v, tw, buf);
            named_barrier(1 + g, FFT_THREADS);
            fft2048_pass3_load(lt, buf, v);
            named_barrier(1 + g, FFT_THREADS);
            fft2048_pass3_store(lt, v, buf);
            named_barrier(1 + g, FFT_THREADS);
            {   // split the packed spectrum: X_a = (Z[k] + conj Z[N-k]) / 2, X_b = (Z[k] - conj Z[N-k]) / 2i
                // undo the packing scale (and the /2); an all-zero frame yields exact zeros
                const float ua = tot.z > 0.f ? ldexpf(0.5f, ea) : 0.f, ub = tot.w > 0.f ? ldexpf(0.5f, eb) : 0.f;
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    const int k = lt + 128 * m;
                    if (k <= FFT_N / 2) {
                        const cf zk = buf[k];
                        const cf zn = buf[(FFT_N - k) & (FFT_N - 1)];
                        const float ar = zk.x + zn.x, ai = zk.y - zn.y;
                        const float br = zk.y + zn.y, bi = zn.x - zk.x;
                        s.stage[k * STAGE_PITCH + fa_idx] = ua * sqrt_approx(fmaf(ar, ar, ai * ai));
                        s.stage[k * STAGE_PITCH + fa_idx + 1] = ub * sqrt_approx(fmaf(br, br, bi * bi));
                    }
                }
            }
            named_barrier(1 + g, FFT_THREADS);  // buf / s.part are rewritten by the next round
        }
        if (!do_fft) continue;
        __syncthreads();

        if (p.mag != nullptr) {
            float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0;
            const bool vec_store = ((p.mag_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 15) == 0);
            if (vec_store) {  // one float4 (4 frames) per lane, two lanes per spectrogram row
                for (int idx = tid; idx < 2 * AEGIS_N_BINS; idx += STFT_THREADS) {
                    const int k = idx >> 1, h = (idx & 1) * 4;
                    const float* st = &s.stage[k * STAGE_PITCH + h];
                    float* dst = mo + static_cast<long long>(k) * p.mag_row_stride + h;
                    if (t0 + h + 3 < T) {
                        *reinterpret_cast<float4*>(dst) = make_float4(st[0], st[1], st[2], st[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (t0 + h + j < T) dst[j] = st[j];
                    }
                }
            } else {          // each warp store = 4 rows x 8 frames
                const int f = lane & 7;
                if (t0 + f < T) {
                    for (int k = warp * 4 + (lane >> 3); k < AEGIS_N_BINS; k += (STFT_THREADS / 32) * 4)
                        mo[static_cast<long long>(k) * p.mag_row_stride + f] = s.stage[k * STAGE_PITCH + f];
                }
            }
        }
        if (p.mel != nullptr) {
            // mel[b] = sum_{k in seg b} rise[k] |X_k|^2 + sum_{k in seg b+1} fall[k] |X_k|^2  (each bin read once).
            // lane = (frame, segment-in-quad): a warp reads 4 stage rows x 8 frames per step, near conflict free.
            float* const rise = reinterpret_cast<float*>(s.buf[0]);          // [n_mels + 1][8], buf is free here
            float* const fall = rise + (MEL_MAX + 1) * TILE_F;
            {
                const int f = lane & 7;
                for (int j = warp * 4 + (lane >> 3); j <= p.n_mels; j += (STFT_THREADS / 32) * 4) {
                    const int k0 = s.mel_seg[j], k1 = s.mel_seg[j + 1];
                    float r = 0.f, fl = 0.f;
                    for (int k = k0; k < k1; ++k) {
                        const float m = s.stage[k * STAGE_PITCH + f];
                        const float2 w = s.mel_rf[k];
                        const float pw = m * m;
                        r = fmaf(w.x, pw, r);
                        fl = fmaf(w.y, pw, fl);
                    }
                    rise[j * TILE_F + f] = r;
                    fall[j * TILE_F + f] = fl;
                }
            }
            __syncthreads();
            float* __restrict__ me = p.mel + static_cast<long long>(clip) * p.mel_clip_stride + t0;
            float vmax = 0.f;
            for (int idx = tid; idx < p.n_mels * TILE_F; idx += STFT_THREADS) {
                const int b = idx >> 3, f = idx & 7;
                const float v = rise[b * TILE_F + f] + fall[(b + 1) * TILE_F + f];
                if (t0 + f < T) {
                    me[static_cast<long long>(b) * p.mel_row_stride + f] = v;
                    vmax = fmaxf(vmax, v);
                }
            }
            if (p.mel_max != nullptr) {
                vmax = warp_max(vmax);
                if (lane == 0) atomic_max_nonneg(p.mel_max + clip, vmax);
            }
        }
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
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_fused_kernel, STFT_THREADS, sizeof(StftSmem)) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const long long max_grid = static_cast<long long>(sm_count()) * per_sm;
    const int grid = static_cast<int>(n_tiles < max_grid ? n_tiles : max_grid);
    stft_fused_kernel<<<grid, STFT_THREADS, sizeof(StftSmem), static_cast<cudaStream_t>(stream)>>>(*p, tiles_per_clip, n_tiles);
    return check_launch("aegis_stft_fused");
}
