// K1: fused frame gather + Hann window + FP32 2048-point FFT -> |X| (+ mel power, frame RMS).
//
// Replaces librosa.stft / feature.melspectrogram / feature.rms as called from
// aegis_engine.py:25,70 and aegis_engine_financial.py:46-50,154 (see include/aegis_b200.h).
//
// Work unit ("tile") = 8 consecutive frames of one clip.  A persistent CTA of 512 threads
// (4 FFT groups x 128 threads) walks tiles with stride gridDim.x, so neighbouring CTAs work on
// neighbouring tiles of the same clip at the same time (halo samples and partially written
// output sectors meet in L2).  Per tile:
//   1. (7*hop + 2048) samples -> shared, float4 loads, zeros outside the clip (centre padding)
//   2. each group packs two real frames (re = frame 2g, im = frame 2g+1) into one complex FFT
//   3. X1[k], X2[k] are separated from Z[k], conj(Z[N-k]); magnitudes go to an [1025][8] stage
//   4. the stage is written out as rows of 8 consecutive frames (one 32 B sector per row and
//      tile) in librosa's [1025, T] layout; mel triangles and the clip maximum are taken from
//      the stage, RMS from the raw samples.
// HBM traffic per frame: hop*4 B read (+ halo, L2-served) and 1025*4 B written.
#include "common.cuh"
#include "fft2048.cuh"

namespace aegis {

constexpr int TILE_F = 8;
constexpr int STFT_THREADS = 512;
constexpr int MAX_HOP = 512;
constexpr int SAMPLES_MAX = (TILE_F - 1) * MAX_HOP + FFT_N;  // 5632
constexpr int STAGE_PITCH = TILE_F + 1;

struct StftSmem {
    float samples[SAMPLES_MAX];
    float window[FFT_N];
    cf bufA[4][BUFA_SIZE];
    cf bufB[4][BUFB_SIZE];
    float stage[AEGIS_N_BINS * STAGE_PITCH];
    float rms_part[16];
};

__global__ void __launch_bounds__(STFT_THREADS, 1)
stft_fused_kernel(const aegis_stft_params p, const int tiles_per_clip, const long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = tid >> 7, lt = tid & 127, lane = tid & 31, warp = tid >> 5;
    const int T = p.n_frames;
    const int hop = p.hop;
    const long long N = p.n_samples;

    for (int i = tid; i < FFT_N; i += STFT_THREADS) s.window[i] = p.window[i];
    FftTwiddles tw;
    fft2048_load_twiddles(lt, reinterpret_cast<const cf*>(p.twiddle), tw);
    const int n_buf = (TILE_F - 1) * hop + FFT_N;
    const bool do_fft = (p.mag != nullptr) || (p.mel != nullptr);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int clip = static_cast<int>(tile / tiles_per_clip);
        const int t0 = static_cast<int>(tile - static_cast<long long>(clip) * tiles_per_clip) * TILE_F;
        const float* __restrict__ yc = p.y + static_cast<long long>(clip) * p.clip_stride;
        const long long g0 = static_cast<long long>(t0) * hop - p.pad;

        __syncthreads();  // everybody is done with the previous tile's samples / stage
        {
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

        if (p.rms != nullptr) {  // sum of squares of each frame, two warps per frame
            const float* fr = s.samples + (warp >> 1) * hop + (warp & 1) * (FFT_N / 2);
            float acc = 0.f;
#pragma unroll 8
            for (int i = lane; i < FFT_N / 2; i += 32) acc = fmaf(fr[i], fr[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) s.rms_part[warp] = acc;
        }

        if (!do_fft) {  // RMS-only call (librosa.feature.rms): no transform needed
            __syncthreads();
            if (p.rms != nullptr && tid < TILE_F && t0 + tid < T) {
                p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t0 + tid] =
                    sqrtf((s.rms_part[2 * tid] + s.rms_part[2 * tid + 1]) * (1.0f / FFT_N));
            }
            continue;
        }
        {   // pass 1: window + pack two real frames
            const float* fa = s.samples + (2 * g) * hop;
            const float* fb = fa + hop;
            cf v[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                const int n = lt + 128 * a;
                const float w = s.window[n];
                v[a] = cf{fa[n] * w, fb[n] * w};
            }
            fft2048_pass1(lt, v, tw, s.bufA[g]);
        }
        named_barrier(1 + g, FFT_THREADS);
        fft2048_pass2(lt, tw, s.bufA[g], s.bufB[g]);
        named_barrier(1 + g, FFT_THREADS);
        fft2048_pass3(lt, s.bufB[g], s.bufA[g]);
        named_barrier(1 + g, FFT_THREADS);

        {   // split the packed spectrum: X1 = (Z[k] + conj Z[N-k]) / 2, X2 = (Z[k] - conj Z[N-k]) / 2i
            const cf* Z = s.bufA[g];
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                const int k = lt + 128 * m;
                if (k <= FFT_N / 2) {
                    const cf zk = Z[k];
                    const cf zn = Z[(FFT_N - k) & (FFT_N - 1)];
                    const float ar = zk.x + zn.x, ai = zk.y - zn.y;
                    const float br = zk.y + zn.y, bi = zn.x - zk.x;
                    s.stage[k * STAGE_PITCH + 2 * g] = 0.5f * sqrt_approx(fmaf(ar, ar, ai * ai));
                    s.stage[k * STAGE_PITCH + 2 * g + 1] = 0.5f * sqrt_approx(fmaf(br, br, bi * bi));
                }
            }
        }
        __syncthreads();

        if (p.rms != nullptr && tid < TILE_F && t0 + tid < T) {
            p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t0 + tid] =
                sqrtf((s.rms_part[2 * tid] + s.rms_part[2 * tid + 1]) * (1.0f / FFT_N));
        }

        const int f = lane & 7;
        const bool f_ok = (t0 + f) < T;
        if (p.mag != nullptr) {  // each warp store = 4 rows x 8 frames
            float* __restrict__ mo = p.mag + static_cast<long long>(clip) * p.mag_clip_stride + t0 + f;
            for (int k = warp * 4 + (lane >> 3); k < AEGIS_N_BINS; k += (STFT_THREADS / 32) * 4) {
                if (f_ok) mo[static_cast<long long>(k) * p.mag_row_stride] = s.stage[k * STAGE_PITCH + f];
            }
        }
        if (p.mel != nullptr) {  // sparse triangles over |X|^2
            float* __restrict__ me = p.mel + static_cast<long long>(clip) * p.mel_clip_stride + t0 + f;
            float vmax = 0.f;
            for (int band = tid >> 3; band < p.n_mels; band += STFT_THREADS / 8) {
                const int ks = __ldg(p.mel_start + band), len = __ldg(p.mel_len + band);
                const float* __restrict__ w = p.mel_w + __ldg(p.mel_off + band);
                float acc = 0.f;
                for (int i = 0; i < len; ++i) {
                    const float m = s.stage[(ks + i) * STAGE_PITCH + f];
                    acc = fmaf(__ldg(w + i), m * m, acc);
                }
                if (f_ok) {
                    me[static_cast<long long>(band) * p.mel_row_stride] = acc;
                    vmax = fmaxf(vmax, acc);
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
    AEGIS_REQUIRE(p->mag == nullptr || p->mag_row_stride >= p->n_frames, "aegis_stft_fused: mag_row_stride < n_frames");
    if (p->mel != nullptr) {
        AEGIS_REQUIRE(p->mel_start && p->mel_len && p->mel_off && p->mel_w && p->n_mels > 0,
                      "aegis_stft_fused: mel tables missing");
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
    const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
    stft_fused_kernel<<<grid, STFT_THREADS, sizeof(StftSmem), static_cast<cudaStream_t>(stream)>>>(*p, tiles_per_clip, n_tiles);
    return check_launch("aegis_stft_fused");
}
