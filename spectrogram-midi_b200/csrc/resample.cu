// K9: audio ingest -- PCM to float, channel mix-down and polyphase rate conversion in one pass.
//
// Replaces the step before the hot path: librosa.load(path, sr=engine_rate) = soundfile's int16 -> float32 / 32768,
// librosa.to_mono (mean over channels), librosa.resample (aegis_engine.py:24, aegis_engine_financial.py:45).  The rate
// conversion is librosa's res_type='polyphase', i.e. scipy.signal.resample_poly(y, up, down): a Kaiser(5.0) windowed-sinc
// low-pass of 2 * 10 * max(up, down) + 1 taps (designed on the host with scipy.signal.firwin, scaled by `up`), applied as
// upfirdn with the filter centred on the output samples.  (librosa's default res_type, soxr_hq, lives in libsoxr, which
// is not in this image and cannot be pinned; see DESIGN.md.)
//
// out[m] = sum_k h[c(m) - k * up] * x[k],   c(m) = (m + n_pre_remove) * down - n_pre_pad
// Each output touches ~ n_taps / up inputs.  A CTA owns a tile of consecutive outputs of one clip, stages the input
// span it needs in shared memory once (coalesced, converted and mixed down on the way in) and keeps the filter in
// shared memory in polyphase order [phase][tap] so the inner loop is two shared loads and one multiply-add.  The sum
// runs over k ascending with separately rounded products and sums -- scipy's upfirdn accumulates in float32 in exactly
// that order -- so the result is bit-identical to scipy.signal.resample_poly on float32 input.
#include "common.cuh"

namespace aegis {

constexpr int RS_THREADS = 256;
constexpr int RS_PER_THREAD = 4;
constexpr int RS_TILE = RS_THREADS * RS_PER_THREAD;

__device__ __forceinline__ float load_mono(const aegis_resample_params& p, const void* clip_base, long long k) {
    // float32 mean over the interleaved channels, summed in channel order as np.mean(axis=0) does
    const int ch = p.n_channels;
    float acc;
    if (p.in_format == 1) {
        const short* s = static_cast<const short*>(clip_base) + k * ch;
        acc = __fmul_rn(static_cast<float>(s[0]), 1.0f / 32768.0f);
        for (int c = 1; c < ch; ++c) acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(s[c]), 1.0f / 32768.0f));
    } else {
        const float* s = static_cast<const float*>(clip_base) + k * ch;
        acc = s[0];
        for (int c = 1; c < ch; ++c) acc = __fadd_rn(acc, s[c]);
    }
    return ch > 1 ? __fdiv_rn(acc, static_cast<float>(ch)) : acc;
}

__global__ void __launch_bounds__(RS_THREADS)
resample_kernel(const aegis_resample_params p, int taps_per_phase, int taps_pitch, int span, int tiles_per_clip) {
    extern __shared__ float rs_smem[];
    float* taps_t = rs_smem;                                   // [up][taps_pitch]: taps_t[phase][j] = h[phase + j * up]
    float* xs = rs_smem + static_cast<size_t>(p.up) * taps_pitch;   // [span] inputs k_lo .. k_lo + span - 1
    const int tid = threadIdx.x;
    const int clip = blockIdx.x / tiles_per_clip;
    const long long m0 = static_cast<long long>(blockIdx.x - clip * tiles_per_clip) * RS_TILE;
    const long long c0 = (m0 + p.n_pre_remove) * p.down - p.n_pre_pad;      // >= 0: half_len >= 10 * down
    const long long k_lo = c0 / p.up - (taps_per_phase - 1);

    for (int idx = tid; idx < p.up * taps_per_phase; idx += RS_THREADS) {
        const int phase = idx / taps_per_phase, j = idx - phase * taps_per_phase;
        const int i = phase + j * p.up;
        taps_t[phase * taps_pitch + j] = i < p.n_taps ? p.taps[i] : 0.f;
    }
    const unsigned char* in_base = static_cast<const unsigned char*>(p.x) +
                                   static_cast<long long>(clip) * p.in_clip_stride * (p.in_format == 1 ? 2 : 4);
    for (int s = tid; s < span; s += RS_THREADS) {
        const long long k = k_lo + s;
        xs[s] = (k >= 0 && k < p.n_in) ? load_mono(p, in_base, k) : 0.f;
    }
    __syncthreads();

    float* out = p.out + static_cast<long long>(clip) * p.out_clip_stride;
#pragma unroll
    for (int r = 0; r < RS_PER_THREAD; ++r) {
        const long long m = m0 + tid + r * RS_THREADS;
        if (m >= p.n_out) continue;
        const long long c = c0 + static_cast<long long>(tid + r * RS_THREADS) * p.down;
        const long long q = c / p.up;                          // newest input that reaches this output
        const int phase = static_cast<int>(c - q * p.up);
        const float* h = taps_t + phase * taps_pitch;
        const float* x = xs + (q - k_lo);                      // x[-j] pairs with h[j]
        float acc = 0.f;
#pragma unroll 8
        for (int j = taps_per_phase - 1; j >= 0; --j)          // k ascending
            acc = __fadd_rn(acc, __fmul_rn(x[-j], h[j]));
        out[m] = acc;
    }
}

// Integer decimation (up == 1: 44.1 -> 22.05 kHz, 88.2 -> 22.05 kHz ...) with the default filter of 20 * DOWN + 1 taps.
// A thread owns R consecutive outputs: their inputs overlap almost completely, so it loads ONE window of
// 20 * DOWN + 1 + (R - 1) * DOWN samples into registers (float4 shared loads) and walks the taps once (broadcast
// shared loads), feeding R accumulation chains -- 0.6 shared wavefronts per output instead of 82, same sums in the
// same order.  Results leave through shared memory so the global stores are coalesced whatever the row alignment.
template <int DOWN, int R>
__global__ void __launch_bounds__(RS_THREADS)
decimate_kernel(const aegis_resample_params p, int span, int tiles_per_clip) {
    constexpr int J = 20 * DOWN + 1;
    constexpr int W = J + (R - 1) * DOWN;                      // window of one thread
    constexpr int W4 = (W + 3) / 4;
    constexpr int TILE = RS_THREADS * R;
    extern __shared__ float rs_smem[];
    float* taps = rs_smem;                                     // [J] padded to a multiple of 4
    float* xs = rs_smem + ((J + 3) & ~3);                      // [span] inputs k_lo ..; reused for the outputs
    const int tid = threadIdx.x;
    const int clip = blockIdx.x / tiles_per_clip;
    const long long m0 = static_cast<long long>(blockIdx.x - clip * tiles_per_clip) * TILE;
    const long long c0 = (m0 + p.n_pre_remove) * DOWN - p.n_pre_pad;
    const long long k_lo = c0 - (J - 1);
    for (int j = tid; j < J; j += RS_THREADS) taps[j] = p.taps[j];
    const unsigned char* in_base = static_cast<const unsigned char*>(p.x) +
                                   static_cast<long long>(clip) * p.in_clip_stride * (p.in_format == 1 ? 2 : 4);
    // the tile is stored in groups of 16 floats padded to 20: thread windows start 80 B apart, so the float4 loads of
    // a quarter warp fall in eight different 16-byte bank groups
    for (int s = tid; s < span; s += RS_THREADS) {
        const long long k = k_lo + s;
        xs[s + (s >> 4) * 4] = (k >= 0 && k < p.n_in) ? load_mono(p, in_base, k) : 0.f;
    }
    __syncthreads();
    static_assert((R * DOWN) % 16 == 0, "a thread's window must start on a group of 16");
    float xw[W4 * 4];
    const float* src = xs + tid * (R * DOWN / 16) * 20;
#pragma unroll
    for (int v = 0; v < W4; ++v) {
        const float4 q = *reinterpret_cast<const float4*>(src + (v / 4) * 20 + (v % 4) * 4);
        xw[4 * v] = q.x; xw[4 * v + 1] = q.y; xw[4 * v + 2] = q.z; xw[4 * v + 3] = q.w;
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {                         // input index ascending, as scipy's upfirdn
        const float h = taps[j];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(xw[(J - 1 - j) + r * DOWN], h));
    }
    __syncthreads();                                           // every window is in registers: reuse xs for the outputs
    static_assert(R % 4 == 0, "outputs are staged as float4");
#pragma unroll
    for (int r = 0; r < R; r += 4)
        *reinterpret_cast<float4*>(xs + tid * R + r) = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
    __syncthreads();
    float* out = p.out + static_cast<long long>(clip) * p.out_clip_stride;
    for (int i = tid; i < TILE; i += RS_THREADS) {
        const long long m = m0 + i;
        if (m < p.n_out) out[m] = xs[i];
    }
}

// equal rates: conversion and mix-down only
__global__ void __launch_bounds__(RS_THREADS)
convert_kernel(const aegis_resample_params p, int blocks_per_clip) {
    const int clip = blockIdx.x / blocks_per_clip;
    const unsigned char* in_base = static_cast<const unsigned char*>(p.x) +
                                   static_cast<long long>(clip) * p.in_clip_stride * (p.in_format == 1 ? 2 : 4);
    float* out = p.out + static_cast<long long>(clip) * p.out_clip_stride;
    const float gain = p.taps[0];
    const long long m0 = static_cast<long long>(blockIdx.x - clip * blocks_per_clip) * RS_TILE + threadIdx.x;
#pragma unroll
    for (int r = 0; r < RS_PER_THREAD; ++r) {
        const long long m = m0 + r * RS_THREADS;
        if (m < p.n_out) out[m] = __fadd_rn(0.f, __fmul_rn(m < p.n_in ? load_mono(p, in_base, m) : 0.f, gain));
    }
}

template <int DOWN, int R>
int launch_decimate(const aegis_resample_params* p, cudaStream_t st) {
    constexpr int J = 20 * DOWN + 1;
    constexpr int TILE = RS_THREADS * R;
    const int span = ((TILE - 1) * DOWN + J + 2 + 3 + 15) & ~15;   // + the float4 overshoot of the last window
    const size_t smem = (((J + 3) & ~3) + span / 16 * 20) * sizeof(float);   // groups of 16 padded to 20
    cudaError_t e = cudaFuncSetAttribute(decimate_kernel<DOWN, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
        set_error("aegis_resample_poly: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
        return 1;
    }
    const long long tiles_per_clip = (p->n_out + TILE - 1) / TILE;
    AEGIS_REQUIRE(tiles_per_clip * p->n_clips < (1LL << 31), "aegis_resample_poly: too many tiles for one launch");
    decimate_kernel<DOWN, R><<<static_cast<unsigned>(tiles_per_clip * p->n_clips), RS_THREADS, smem, st>>>(*p, span, static_cast<int>(tiles_per_clip));
    return check_launch("aegis_resample_poly(decimate)");
}

}  // namespace aegis

extern "C" int aegis_resample_poly(const aegis_resample_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->x && p->out && p->taps, "aegis_resample_poly: x / out / taps must be set");
    AEGIS_REQUIRE(p->n_clips >= 0 && p->n_in >= 0 && p->n_out >= 0, "aegis_resample_poly: negative size");
    AEGIS_REQUIRE(p->up >= 1 && p->down >= 1 && p->n_taps >= 1, "aegis_resample_poly: up, down and n_taps must be positive");
    AEGIS_REQUIRE(p->in_format == 0 || p->in_format == 1, "aegis_resample_poly: in_format is 0 (float32) or 1 (int16)");
    AEGIS_REQUIRE(p->n_channels >= 1 && p->n_channels <= 8, "aegis_resample_poly: 1..8 interleaved channels");
    AEGIS_REQUIRE(p->in_clip_stride >= p->n_in * p->n_channels && p->out_clip_stride >= p->n_out, "aegis_resample_poly: clip strides too small");
    AEGIS_REQUIRE(p->n_pre_pad >= 0 && p->n_pre_remove >= 0 &&
                  static_cast<long long>(p->n_pre_remove) * p->down >= p->n_pre_pad, "aegis_resample_poly: bad filter centring");
    if (p->n_clips == 0 || p->n_out == 0) return 0;
    if (p->up == 1 && p->down == 1 && p->n_taps == 1 && p->n_pre_pad == 0 && p->n_pre_remove == 0) {
        const long long blocks_per_clip = (p->n_out + RS_TILE - 1) / RS_TILE;
        AEGIS_REQUIRE(blocks_per_clip * p->n_clips < (1LL << 31), "aegis_resample_poly: too many tiles for one launch");
        convert_kernel<<<static_cast<unsigned>(blocks_per_clip * p->n_clips), RS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*p, static_cast<int>(blocks_per_clip));
        return check_launch("aegis_resample_poly(convert)");
    }
    if (p->up == 1 && p->down == 2 && p->n_taps == 41) return launch_decimate<2, 8>(p, static_cast<cudaStream_t>(stream));
    if (p->up == 1 && p->down == 4 && p->n_taps == 81) return launch_decimate<4, 4>(p, static_cast<cudaStream_t>(stream));
    const int taps_per_phase = (p->n_taps + p->up - 1) / p->up;
    const int taps_pitch = taps_per_phase | 1;                 // odd pitch: phases spread over the banks
    // inputs a tile can touch: from floor(c0 / up) - (J - 1) to floor((c0 + (TILE - 1) * down) / up)
    const int span = static_cast<int>((static_cast<long long>(RS_TILE - 1) * p->down) / p->up) + taps_per_phase + 2;
    const size_t smem = (static_cast<size_t>(p->up) * taps_pitch + span) * sizeof(float);
    AEGIS_REQUIRE(smem <= 200 * 1024, "aegis_resample_poly: filter of %d taps x tile span %d needs %zu B shared memory (> 200 KB); reduce the ratio's terms",
                  p->n_taps, span, smem);
    cudaError_t e = cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
        set_error("aegis_resample_poly: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
        return 1;
    }
    const long long tiles_per_clip = (p->n_out + RS_TILE - 1) / RS_TILE;
    AEGIS_REQUIRE(tiles_per_clip * p->n_clips < (1LL << 31), "aegis_resample_poly: too many tiles for one launch");
    resample_kernel<<<static_cast<unsigned>(tiles_per_clip * p->n_clips), RS_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        *p, taps_per_phase, taps_pitch, span, static_cast<int>(tiles_per_clip));
    return check_launch("aegis_resample_poly");
}
