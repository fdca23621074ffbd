// Pieces shared by the two logic filters (K7 notes.cu, K8 notes_fin.cu): librosa.amplitude_to_db(rms, ref=np.max) in
// numpy's float32 arithmetic and the MIDI velocity map (aegis_engine_core/midi_logic.py:51,71;
// aegis_engine_core_v2/midi_logic_financial.py:196,218).
#pragma once
#include "common.cuh"

namespace aegis {

__device__ __forceinline__ float db10_f32(float power, float amin) {
    // 10.0 * np.log10(np.maximum(amin, power)) in float32; log10 is evaluated in double and rounded once
    const float v = fmaxf(amin, power);
    return __fmul_rn(10.0f, static_cast<float>(log10(static_cast<double>(v))));
}

// amplitude_to_db(rms, ref=np.max)[t]: power_to_db(rms**2, ref=max**2, amin=1e-10, top_db=80) (librosa, float32)
__device__ __forceinline__ float rms_db_f32(float rms, float ref) {
    const float amin = 1e-10f;
    const float mag = fabsf(rms);
    const float ref_db = db10_f32(__fmul_rn(ref, ref), amin);
    const float e = __fadd_rn(db10_f32(__fmul_rn(mag, mag), amin), -ref_db);
    // top_db: max(log_spec, log_spec.max() - 80); the maximum is the reference frame itself
    const float top = __fadd_rn(__fadd_rn(ref_db, -ref_db), -80.0f);
    return fmaxf(e, top);
}

// int(np.clip((energy + 80) * 1.5, 0, 127)) on a float32 energy
__device__ __forceinline__ int velocity_from_db(float energy) {
    return static_cast<int>(fminf(fmaxf(__fmul_rn(__fadd_rn(energy, 80.0f), 1.5f), 0.0f), 127.0f));
}

// per clip: max of |rms| (the reference level of amplitude_to_db)
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rms_max_kernel(const float* __restrict__ rms, long long clip_stride, int n_frames, float* __restrict__ rms_max) {
    __shared__ float red[THREADS / 32];
    const int clip = blockIdx.x;
    const float* r = rms + static_cast<long long>(clip) * clip_stride;
    float m = 0.f;
    for (int t = threadIdx.x; t < n_frames; t += THREADS) m = fmaxf(m, fabsf(r[t]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < THREADS / 32; ++w) m = fmaxf(m, red[w]);
        rms_max[clip] = m;
    }
}

}  // namespace aegis
