// Device-side corpus synthesis: Karplus-Strong plucks and noise rakes (a restatement of the
// reference's fixture generator generate_test_signal.py:5-53 with a counter-based RNG), one thread
// per event.  Benchmark/test input generation only -- not part of the analysis path.
#include "common.cuh"

namespace aegis {

constexpr int SYNTH_MAX_PERIOD = 640;

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float u01(unsigned seed, unsigned i) {  // (0, 1)
    return (static_cast<float>(hash32(seed ^ (i * 0x9E3779B9U + 0x85ebca6bU)) >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(64)
synth_kernel(const aegis_synth_params p) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_events) return;
    float* out = p.out + static_cast<long long>(p.ev_clip[e]) * p.clip_stride + p.ev_start[e];
    const int len = p.ev_len[e], period = min(p.ev_period[e], SYNTH_MAX_PERIOD);
    const float amp = p.ev_amp[e];
    const unsigned seed = p.ev_seed[e];
    if (period > 0) {
        float buf[SYNTH_MAX_PERIOD];
        for (int i = 0; i < period; ++i) buf[i] = 2.0f * u01(seed, i) - 1.0f;
        float prev = buf[period - 1];
        int ptr = 0;
        for (int i = 0; i < len; ++i) {
            const float val = buf[ptr];
            out[i] = amp * val;
            prev = 0.5f * (val + prev) * p.decay;  // the reference reads buf[ptr-1] after overwriting it
            buf[ptr] = prev;
            ptr = (ptr + 1 == period) ? 0 : ptr + 1;
        }
    } else {
        const float inv = len > 1 ? 1.0f / static_cast<float>(len - 1) : 0.f;
        for (int i = 0; i < len; ++i) {
            const float u1 = u01(seed, 2 * i), u2 = u01(seed, 2 * i + 1);
            const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
            const float env = 1.0f - static_cast<float>(i) * inv;
            out[i] = amp * 0.8f * g * env * env;
        }
    }
}

}  // namespace aegis

extern "C" int aegis_synth_ks(const aegis_synth_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr && p->out, "aegis_synth_ks: out must be set");
    AEGIS_REQUIRE(p->n_events >= 0, "aegis_synth_ks: negative event count");
    if (p->n_events == 0) return 0;
    AEGIS_REQUIRE(p->ev_clip && p->ev_start && p->ev_len && p->ev_period && p->ev_amp && p->ev_seed, "aegis_synth_ks: event arrays missing");
    synth_kernel<<<(p->n_events + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(*p);
    return check_launch("aegis_synth_ks");
}
