// K3: pitch/voicing HMM Viterbi decode for pYIN (librosa.sequence.viterbi inside librosa.pyin;
// call sites aegis_engine.py:63,67,190,216, aegis_engine_core/worker.py:9-15,
// aegis_engine_financial.py:63-69; algorithm SURVEY.md Appendix A.5 steps 10-12).
//
// States: 2*n_bins (voiced bins 0..n-1, unvoiced n..2n-1).  librosa evaluates a dense
// [2n x 2n] max-plus product per frame in float64 with log(p + tiny) everywhere, so "impossible"
// transitions/observations cost a finite log(tiny) and still compete.  This kernel is exact with
// respect to that definition (same float64 adds, first-index argmax) but never touches the dense
// matrix:
//   * in-band sources (|b - b'| <= half_width, both voicings) are enumerated from the banded
//     log-transition table; rows that differ in the last ulp (librosa's pairwise row sums) are
//     kept as "variants", the few interior ones in shared memory, truncated edge rows in global;
//   * out-of-band sources all carry log(tiny): their best is a leftmost-max prefix/suffix query
//     (warp-shuffle scans over V, second level recomputed per warp), O(1) per destination.
// One CTA per clip, one thread per pitch bin (it owns the voiced and the unvoiced state of that
// bin, so every V[b'] load feeds four candidates); sequential over frames with one barrier per
// frame; V, scans and observations are ping-ponged in shared memory.  Back-pointers stream to
// global (uint16, coalesced); a second kernel walks them, one thread per clip.
#include <cfloat>
#include <cmath>
#include "common.cuh"

namespace aegis {

constexpr int VT_MAX_BINS = 512;
constexpr int VT_HALO = 64;            // >= half_width
constexpr int VT_MAX_W = 2 * VT_HALO + 1;
constexpr int VT_SMEM_VARIANTS = 6;
constexpr int VT_MAX_WARPS = VT_MAX_BINS / 32;

struct VitSmem {
    double V[2][2][VT_MAX_BINS + 2 * VT_HALO];     // [ping][voicing][halo | bins | halo]
    double pw_val[2][2][VT_MAX_BINS];              // within-warp leftmost prefix max
    double sw_val[2][2][VT_MAX_BINS];              // within-warp leftmost suffix max
    double obs_lp[2][VT_MAX_BINS];                 // log(obs + tiny) of the voiced states
    double lt[VT_SMEM_VARIANTS][2][VT_MAX_W];      // interior transition variants
    double seg_val[2][2][VT_MAX_WARPS];            // per-warp leftmost max
    short pw_arg[2][2][VT_MAX_BINS];
    short sw_arg[2][2][VT_MAX_BINS];
    short seg_arg[2][2][VT_MAX_WARPS];
    unsigned char rowvar[VT_MAX_BINS + 2 * VT_HALO];
};

struct VA {
    double v;
    int a;
};

__device__ __forceinline__ VA shfl_va(VA x, int src) {
    return VA{__shfl_sync(0xffffffffu, x.v, src), __shfl_sync(0xffffffffu, x.a, src)};
}

__global__ void __launch_bounds__(VT_MAX_BINS, 1)
viterbi_forward_kernel(const aegis_viterbi_params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VitSmem& s = *reinterpret_cast<VitSmem*>(smem_raw);
    const int clip = blockIdx.x;
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    const int n = p.n_pitch_bins, hw = p.half_width, W = 2 * hw + 1, T = p.n_frames;
    const int n_warps = (n + 31) >> 5;
    const int nsv = min(p.n_interior_variants, VT_SMEM_VARIANTS);
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;
    const bool live = b < n;

    // ---- one-time shared set-up
    for (int i = b; i < 2 * 2 * (VT_MAX_BINS + 2 * VT_HALO); i += blockDim.x) (&s.V[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * VT_MAX_BINS; i += blockDim.x) (&s.obs_lp[0][0])[i] = LOGTINY;
    for (int i = b; i < VT_MAX_BINS + 2 * VT_HALO; i += blockDim.x) {
        const int src = i - VT_HALO;
        s.rowvar[i] = (src >= 0 && src < n) ? static_cast<unsigned char>(__ldg(p.row_variant + src)) : 0;
    }
    for (int i = b; i < nsv * 2 * W; i += blockDim.x) {
        const int var = i / (2 * W), rem = i - var * 2 * W;
        s.lt[var][rem / W][rem % W] = __ldg(p.lt_variants + i);
    }
    __syncthreads();

    const long long f0idx = static_cast<long long>(clip) * T;
    const unsigned short* __restrict__ cbin = p.cand_bin + f0idx * p.max_cand;
    const double* __restrict__ cprob = p.cand_prob + f0idx * p.max_cand;
    const int* __restrict__ ccnt = p.cand_count + f0idx;
    const double* __restrict__ vprob = p.voiced_prob + f0idx;
    unsigned short* __restrict__ bp_out = p.backptr + f0idx * (2 * n);

    // scatter frame 0 observations
    {
        const int cnt = min(__ldg(ccnt), p.max_cand);
        if (b < cnt) s.obs_lp[0][cbin[b]] = log(cprob[b] + DBL_MIN);
    }
    __syncthreads();

    double vnew0 = NEG_INF, vnew1 = NEG_INF;   // this bin's voiced / unvoiced value
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, nxt = cur ^ 1;
        // prefetch the next frame's sparse observation
        int ncnt = 0, nbin = 0;
        double nprob = 0.0;
        if (t + 1 < T) {
            ncnt = min(__ldg(ccnt + t + 1), p.max_cand);
            if (b < ncnt) {
                nbin = __ldg(cbin + static_cast<long long>(t + 1) * p.max_cand + b);
                nprob = __ldg(cprob + static_cast<long long>(t + 1) * p.max_cand + b);
            }
        }
        const double vp = __ldg(vprob + t);
        const double lp_u = log((1.0 - vp) / static_cast<double>(n) + DBL_MIN);
        double lp_v = LOGTINY;
        if (live) {
            lp_v = s.obs_lp[cur][b];
            s.obs_lp[cur][b] = LOGTINY;  // reset for frame t+2
        }

        if (t == 0) {
            vnew0 = lp_v + LOGTINY;               // log(p_init = 0 + tiny)
            vnew1 = lp_u + p.log_init_unvoiced;   // log(1/n + tiny)
        } else {
            // second-level leftmost prefix / suffix maxima over warp segments (registers, per warp)
            VA segp[2], segs[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                VA x{NEG_INF, 0};
                if (lane < n_warps) x = VA{s.seg_val[cur][v][lane], s.seg_arg[cur][v][lane]};
                VA pf = x, sf = x;
#pragma unroll
                for (int o = 1; o < VT_MAX_WARPS; o <<= 1) {
                    const VA up = shfl_va(pf, max(lane - o, 0));
                    if (lane >= o && up.v >= pf.v) pf = up;
                    const VA dn = shfl_va(sf, min(lane + o, 31));
                    if (lane + o < 32 && dn.v > sf.v) sf = dn;
                }
                segp[v] = pf;
                segs[v] = sf;
            }
            // in-band candidates; [dest voicing][source voicing]
            double best[2][2] = {{NEG_INF, NEG_INF}, {NEG_INF, NEG_INF}};
            int arg[2][2] = {{0, 0}, {0, 0}};
            if (live) {
                const double* V0 = &s.V[cur][0][VT_HALO + b - hw];
                const double* V1 = &s.V[cur][1][VT_HALO + b - hw];
                const unsigned char* rv = &s.rowvar[VT_HALO + b - hw];
                for (int d = 0; d < W; ++d) {
                    const int var = rv[d];
                    const int o = W - 1 - d;  // dest offset inside the source row's band
                    double ls, lx;
                    if (var < nsv) {
                        ls = s.lt[var][0][o];
                        lx = s.lt[var][1][o];
                    } else {
                        const double* g = p.lt_variants + static_cast<long long>(var) * 2 * W + o;
                        ls = __ldg(g);
                        lx = __ldg(g + W);
                    }
                    const double x0 = V0[d], x1 = V1[d];
                    const int src = b - hw + d;
                    double c;
                    c = x0 + ls; if (c > best[0][0]) { best[0][0] = c; arg[0][0] = src; }
                    c = x1 + lx; if (c > best[0][1]) { best[0][1] = c; arg[0][1] = src; }
                    c = x0 + lx; if (c > best[1][0]) { best[1][0] = c; arg[1][0] = src; }
                    c = x1 + ls; if (c > best[1][1]) { best[1][1] = c; arg[1][1] = src; }
                }
            }
            // out-of-band competitors: every one costs log(tiny), whatever the voicing
            const int lo_idx = b - hw - 1, hi_idx = b + hw + 1;
#pragma unroll
            for (int sv = 0; sv < 2; ++sv) {
                const int li = min(max(lo_idx, 0), VT_MAX_BINS - 1), hi = min(max(hi_idx, 0), VT_MAX_BINS - 1);
                VA lowq{s.pw_val[cur][sv][li], s.pw_arg[cur][sv][li]};
                const VA pseg = shfl_va(segp[sv], max((li >> 5) - 1, 0));
                if ((li >> 5) > 0 && pseg.v >= lowq.v) lowq = pseg;
                VA highq{s.sw_val[cur][sv][hi], s.sw_arg[cur][sv][hi]};
                const VA sseg = shfl_va(segs[sv], min((hi >> 5) + 1, 31));
                if ((hi >> 5) + 1 < n_warps && sseg.v > highq.v) highq = sseg;
                if (live) {
#pragma unroll
                    for (int dv = 0; dv < 2; ++dv) {
                        double bv = best[dv][sv];
                        int ba = arg[dv][sv];
                        if (lo_idx >= 0) {  // lower indices than the band: wins ties
                            const double c = lowq.v + LOGTINY;
                            if (c >= bv) { bv = c; ba = lowq.a; }
                        }
                        if (hi_idx < n) {   // higher indices: must be strictly better
                            const double c = highq.v + LOGTINY;
                            if (c > bv) { bv = c; ba = highq.a; }
                        }
                        best[dv][sv] = bv;
                        arg[dv][sv] = ba;
                    }
                }
            }
            if (live) {
                // voiced sources (indices < n) win ties against unvoiced sources
                const bool u0 = best[0][1] > best[0][0], u1 = best[1][1] > best[1][0];
                const double m0 = u0 ? best[0][1] : best[0][0], m1 = u1 ? best[1][1] : best[1][0];
                const int k0 = u0 ? n + arg[0][1] : arg[0][0], k1 = u1 ? n + arg[1][1] : arg[1][0];
                vnew0 = lp_v + m0;
                vnew1 = lp_u + m1;
                unsigned short* row = bp_out + static_cast<long long>(t) * (2 * n);
                row[b] = static_cast<unsigned short>(k0);
                row[n + b] = static_cast<unsigned short>(k1);
            }
        }

        // publish V[t] and its within-warp leftmost prefix / suffix maxima
        if (live) {
            s.V[nxt][0][VT_HALO + b] = vnew0;
            s.V[nxt][1][VT_HALO + b] = vnew1;
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const VA x{live ? (v == 0 ? vnew0 : vnew1) : NEG_INF, b};
            VA pf = x, sf = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const VA up = shfl_va(pf, max(lane - o, 0));
                if (lane >= o && up.v >= pf.v) pf = up;
                const VA dn = shfl_va(sf, min(lane + o, 31));
                if (lane + o < 32 && dn.v > sf.v) sf = dn;
            }
            s.pw_val[nxt][v][b] = pf.v;
            s.pw_arg[nxt][v][b] = static_cast<short>(pf.a);
            s.sw_val[nxt][v][b] = sf.v;
            s.sw_arg[nxt][v][b] = static_cast<short>(sf.a);
            if (lane == 31) {
                s.seg_val[nxt][v][warp] = pf.v;
                s.seg_arg[nxt][v][warp] = static_cast<short>(pf.a);
            }
        }
        if (b < ncnt) s.obs_lp[nxt][nbin] = log(nprob + DBL_MIN);
        __syncthreads();
    }
    if (live && T > 0) {
        double* fv = p.final_value + static_cast<long long>(clip) * (2 * n);
        fv[b] = vnew0;
        fv[n + b] = vnew1;
    }
}

// back-trace: one thread per clip (latency bound; all clips walk in parallel)
__global__ void __launch_bounds__(64)
viterbi_backtrace_kernel(const aegis_viterbi_params p) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= p.n_clips) return;
    const int n = p.n_pitch_bins, T = p.n_frames;
    if (T <= 0) return;
    const double* fv = p.final_value + static_cast<long long>(clip) * (2 * n);
    int st = 0;
    double best = fv[0];
    for (int j = 1; j < 2 * n; ++j) {
        const double v = fv[j];
        if (v > best) { best = v; st = j; }
    }
    const long long base = static_cast<long long>(clip) * T;
    const unsigned short* bp = p.backptr + base * (2 * n);
    for (int t = T - 1; t >= 0; --t) {
        p.states[base + t] = static_cast<unsigned short>(st);
        const bool voiced = st < n;
        p.voiced_flag[base + t] = voiced ? 1 : 0;
        p.f0[base + t] = voiced ? __ldg(p.freqs + st) : p.fill_value;
        if (t > 0) st = bp[static_cast<long long>(t) * (2 * n) + st];
    }
}

}  // namespace aegis

extern "C" int aegis_viterbi(const aegis_viterbi_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "aegis_viterbi: null params");
    AEGIS_REQUIRE(p->n_pitch_bins >= 2 && p->n_pitch_bins <= VT_MAX_BINS, "aegis_viterbi: n_pitch_bins=%d unsupported (<= %d)", p->n_pitch_bins, VT_MAX_BINS);
    AEGIS_REQUIRE(p->half_width >= 1 && p->half_width <= VT_HALO, "aegis_viterbi: half_width=%d unsupported (<= %d)", p->half_width, VT_HALO);
    AEGIS_REQUIRE(p->n_variants >= 1 && p->n_variants <= 255 && p->n_interior_variants >= 1, "aegis_viterbi: bad variant counts");
    AEGIS_REQUIRE(p->max_cand >= 1 && p->max_cand <= p->n_pitch_bins, "aegis_viterbi: max_cand must be 1..n_pitch_bins");
    AEGIS_REQUIRE(p->cand_bin && p->cand_prob && p->cand_count && p->voiced_prob && p->lt_variants && p->row_variant && p->freqs,
                  "aegis_viterbi: inputs missing");
    AEGIS_REQUIRE(p->backptr && p->final_value && p->states && p->f0 && p->voiced_flag, "aegis_viterbi: outputs missing");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(viterbi_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(VitSmem)));
    if (e != cudaSuccess) {
        set_error("aegis_viterbi: cannot reserve %zu B shared memory: %s", sizeof(VitSmem), cudaGetErrorString(e));
        return 2;
    }
    const int block = ((p->n_pitch_bins + 31) / 32) * 32;
    viterbi_forward_kernel<<<p->n_clips, block, sizeof(VitSmem), st>>>(*p);
    if (int rc = check_launch("aegis_viterbi(forward)")) return rc;
    viterbi_backtrace_kernel<<<(p->n_clips + 63) / 64, 64, 0, st>>>(*p);
    return check_launch("aegis_viterbi(backtrace)");
}
