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
//   * out-of-band sources all carry log(tiny): only the block's leftmost global maximum can matter
//     (see the kernel comment), one block-wide argmax per frame;
//   * in-band sources are pruned in chunks of 8 with exact upper bounds, so a destination typically
//     evaluates a few dozen of its 202 in-band candidates.
// One CTA per clip, one thread per pitch bin (it owns the voiced and the unvoiced state of that
// bin, so every V[b'] load feeds four candidates); sequential over frames with one barrier per
// frame; V, scans and observations are ping-ponged in shared memory.  Back-pointers stream to
// global (uint16, coalesced); a second kernel walks them, one thread per clip.
#include <cfloat>
#include <cmath>
#include "common.cuh"

namespace aegis {

constexpr int VT_MAX_BINS = 512;
constexpr int VT_HALO = 64;            // >= half_width + 8
constexpr int VT_MAX_W = 2 * VT_HALO + 1;
constexpr int VT_SMEM_VARIANTS = 6;
constexpr int VT_MAX_WARPS = VT_MAX_BINS / 32;
constexpr int VT_CHUNK = 8;            // sources are pruned in aligned chunks of 8 bins
constexpr int VT_CHUNK_PAD = 8;        // chunk indices -8 .. (512/8 + 8)
constexpr int VT_N_CHUNKS = VT_MAX_BINS / VT_CHUNK + 2 * VT_CHUNK_PAD;
#ifndef VT_SMALL_MINB
#define VT_SMALL_MINB 2     // 448-thread CTAs per SM for the 441-bin configuration (3 spills and is slower)
#endif

struct VitSmem {
    double V[2][2][VT_MAX_BINS + 2 * VT_HALO];     // [ping][voicing][halo | bins | halo], halo = -inf
    double M[2][2][VT_N_CHUNKS];                   // max of V over each aligned chunk of 8 bins (-inf outside)
    double obs_lp[2][VT_MAX_BINS];                 // log(obs + tiny) of the voiced states
    double lt[VT_SMEM_VARIANTS][2][VT_MAX_W];      // interior transition variants [variant][same|switch][offset]
    double ubr[2][VT_MAX_W + VT_CHUNK];            // [same|switch][q]: max over the shared-memory (interior) row
                                                   // variants and over offsets q-7..q; other rows add their own dub
    double M4[2][2][VT_N_CHUNKS / 4];              // max of M over each aligned group of 4 chunks (= one warp of bins)
    double ubr4[2][VT_MAX_W + 5 * VT_CHUNK];       // [same|switch][q]: max of ubr over q, q-8, q-16, q-24
    double seg_val[2][2][VT_MAX_WARPS];            // per-warp leftmost max of V
    short seg_arg[2][2][VT_MAX_WARPS];
    unsigned char rowvar[VT_MAX_BINS + 2 * VT_HALO];
    double lp_u[2];                                // log-observation of the unvoiced states of frame t (same for all bins)
};

struct VA {
    double v;
    int a;
};

// leftmost-max combine under a butterfly exchange: every lane ends with (max value, lowest index attaining it)
__device__ __forceinline__ VA butterfly_leftmost(VA x, int mask) {
    const double ov = __shfl_xor_sync(0xffffffffu, x.v, mask);
    const int oa = __shfl_xor_sync(0xffffffffu, x.a, mask);
    if (ov > x.v || (ov == x.v && oa < x.a)) { x.v = ov; x.a = oa; }
    return x;
}

// HW: half width of the transition band (compile time so the loops unroll).  MAXT/MINB: the 441-bin
// E2..C6 configuration runs 448-thread CTAs two per SM; wider pitch ranges use one 512-thread CTA.
//
// Exactness of the pruning (everything is float64, first-index argmax as in numpy):
//  * Out-of-band sources of a voicing block all cost log(tiny).  Let g be the block's LEFTMOST global
//    maximum.  If g lies inside the band its in-band candidate (>= V[g] - 14) beats every out-of-band
//    candidate (<= V[g] - 708); otherwise g itself is the best out-of-band source of its side and the
//    other side can at best tie with a higher index (or is strictly smaller).  So one block-wide
//    leftmost argmax per frame replaces the prefix/suffix scans.
//  * In-band sources are visited in aligned chunks of 8.  A chunk is skipped when
//    chunk_max + max(lt over its offsets, over all row variants) is < L, the value of a real candidate
//    (the destination's own bin), or <= the running best of lower-index sources: floating-point addition
//    is monotone, so no source in the chunk can reach the maximum or win a tie.  Chunks that survive are
//    evaluated exactly, in ascending source order with a strict >, voiced sources before unvoiced.
template <int HW, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
viterbi_forward_kernel(const aegis_viterbi_params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VitSmem& s = *reinterpret_cast<VitSmem*>(smem_raw);
    const int clip = blockIdx.x;
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    constexpr int hw = HW, W = 2 * HW + 1;
    constexpr int NCH = (W + 2 * (VT_CHUNK - 1) + VT_CHUNK - 1) / VT_CHUNK;  // chunks that can touch a band
    static_assert(HW + VT_CHUNK <= VT_HALO, "halo too small");
    static_assert(VT_CHUNK_PAD % 4 == 0 && VT_N_CHUNKS % 4 == 0, "chunk groups must align with the padding");
    const int n = p.n_pitch_bins, T = p.n_frames;
    const int n_warps = (n + 31) >> 5;
    const int nsv = min(p.n_interior_variants, VT_SMEM_VARIANTS);
    const bool interior_in_smem = p.n_interior_variants <= VT_SMEM_VARIANTS;  // fast path may index s.lt by variant
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;
    const bool live = b < n;

    // ---- one-time shared set-up
    for (int i = b; i < 2 * 2 * (VT_MAX_BINS + 2 * VT_HALO); i += blockDim.x) (&s.V[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * 2 * VT_N_CHUNKS; i += blockDim.x) (&s.M[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * 2 * (VT_N_CHUNKS / 4); i += blockDim.x) (&s.M4[0][0][0])[i] = NEG_INF;
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
    // Chunk upper bounds.  base[sel][o] = max over the interior variants held in shared memory (they differ in
    // the last ulp); ubr[sel][q] = max_{j<8} base[sel][q-j].  A source row with any other variant (the truncated
    // edge rows are up to log 2 larger) carries its own excess dub = max_{sel,o}(lt_row - base) + margin, which is
    // added to V before the chunk maxima are taken: max_chunk(V + dub) + ubr bounds every candidate of the chunk.
    for (int i = b; i < 2 * (W + VT_CHUNK); i += blockDim.x) {
        const int sel = i / (W + VT_CHUNK), q = i - sel * (W + VT_CHUNK);
        double m = NEG_INF;
        for (int j = 0; j < VT_CHUNK; ++j) {
            const int o = q - j;
            if (o < 0 || o >= W) continue;
            for (int var = 0; var < nsv; ++var) m = fmax(m, s.lt[var][sel][o]);
        }
        s.ubr[sel][q] = m;
    }
    __syncthreads();
    // bound of a group of 4 chunks whose FIRST chunk has offset q: the chunks sit at offsets q, q-8, q-16, q-24
    for (int i = b; i < 2 * (W + 5 * VT_CHUNK); i += blockDim.x) {
        const int sel = i / (W + 5 * VT_CHUNK), q = i - sel * (W + 5 * VT_CHUNK);
        double m = NEG_INF;
        for (int j = 0; j < 4; ++j) {
            const int qq = q - VT_CHUNK * j;
            if (qq >= 0 && qq < W + VT_CHUNK) m = fmax(m, s.ubr[sel][qq]);
        }
        s.ubr4[sel][q] = m;
    }
    double dub = 0.0;
    if (live) {
        const int var = __ldg(p.row_variant + b);
        if (var >= nsv) {
            double ex = 0.0;
            for (int sel = 0; sel < 2; ++sel)
                for (int o = 0; o < W; ++o) {
                    const double v = __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o);
                    double base = NEG_INF;
                    for (int iv = 0; iv < nsv; ++iv) base = fmax(base, s.lt[iv][sel][o]);
                    if (v > NEG_INF) ex = fmax(ex, v - base);
                }
            dub = ex + 1e-6;  // margin >> any rounding of the bound arithmetic (|V| < 1e10)
        }
    }
    __syncthreads();

    const long long f0idx = static_cast<long long>(clip) * T;
    const unsigned short* __restrict__ cbin = p.cand_bin + f0idx * p.max_cand;
    const double* __restrict__ cprob = p.cand_prob + f0idx * p.max_cand;
    const int* __restrict__ ccnt = p.cand_count + f0idx;
    const double* __restrict__ vprob = p.voiced_prob + f0idx;
    unsigned short* __restrict__ bp_out = p.backptr + f0idx * (2 * n);

    if (b == 0 && T > 0) s.lp_u[0] = log((1.0 - __ldg(vprob)) / static_cast<double>(n) + DBL_MIN);
    {   // scatter frame 0 observations
        const int cnt = min(__ldg(ccnt), p.max_cand);
        if (b < cnt) s.obs_lp[0][cbin[b]] = log(cprob[b] + DBL_MIN);
    }
    __syncthreads();

    // table entry for (source row variant, same|switch, offset)
    auto lt_at = [&](int var, int sel, int o) -> double {
        return (var < nsv) ? s.lt[var][sel][o] : __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o);
    };

    double vnew0 = NEG_INF, vnew1 = NEG_INF;   // this bin's voiced / unvoiced value
    int prev0 = b, prev1 = n + b;               // winning sources of the previous frame (temporal coherence)
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, nxt = cur ^ 1;
        int ncnt = 0, nbin = 0;   // prefetch the next frame's sparse observation
        double nprob = 0.0;
        if (t + 1 < T) {
            ncnt = min(__ldg(ccnt + t + 1), p.max_cand);
            if (b < ncnt) {
                nbin = __ldg(cbin + static_cast<long long>(t + 1) * p.max_cand + b);
                nprob = __ldg(cprob + static_cast<long long>(t + 1) * p.max_cand + b);
            }
        }
        // one warp evaluates the (bin independent) unvoiced log-observation of the NEXT frame; everybody reads it after
        // the frame barrier (13 of the 14 warps used to spend ~4 % of their instructions on the same double log)
        if (warp == 0 && t + 1 < T) {
            const double vpn = __ldg(vprob + t + 1);
            const double v = log((1.0 - vpn) / static_cast<double>(n) + DBL_MIN);
            if (lane == 0) s.lp_u[nxt] = v;
        }
        const double lp_u = s.lp_u[cur];
        double lp_v = LOGTINY;
        if (live) {
            lp_v = s.obs_lp[cur][b];
            s.obs_lp[cur][b] = LOGTINY;  // reset for frame t+2
        }

        if (t == 0) {
            vnew0 = lp_v + LOGTINY;               // log(p_init = 0 + tiny)
            vnew1 = lp_u + p.log_init_unvoiced;   // log(1/n + tiny)
        } else {
            // leftmost global maximum of each voicing block (second level of the argmax, per warp)
            VA gmax[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                VA x{NEG_INF, 0x7fffffff};
                if (lane < n_warps) x = VA{s.seg_val[cur][v][lane], s.seg_arg[cur][v][lane]};
#pragma unroll
                for (int m = 1; m < 32; m <<= 1) x = butterfly_leftmost(x, m);
                gmax[v] = x;
            }
            if (live) {
                const double* Vc0 = &s.V[cur][0][VT_HALO];
                const double* Vc1 = &s.V[cur][1][VT_HALO];
                const unsigned char* rvc = &s.rowvar[VT_HALO];
                // lower bounds: the candidates from this destination's own bin (offset hw)
                double L0, L1;
                {
                    const int var = rvc[b];
                    const double ls = lt_at(var, 0, hw), lx = lt_at(var, 1, hw);
                    const double x0 = Vc0[b], x1 = Vc1[b];
                    L0 = fmax(x0 + ls, x1 + lx);
                    L1 = fmax(x0 + lx, x1 + ls);
                }
                // ... and from the source that won in the previous frame (same voicing block, same bin): decoded
                // paths move slowly, so this real candidate is usually (near) optimal and prunes almost every chunk
                {
                    const int sv0 = prev0 >= n, bs0 = prev0 - sv0 * n, o0 = b - bs0 + hw;
                    if (o0 >= 0 && o0 < W) L0 = fmax(L0, (sv0 ? Vc1 : Vc0)[bs0] + lt_at(rvc[bs0], sv0, o0));
                    const int sv1 = prev1 >= n, bs1 = prev1 - sv1 * n, o1 = b - bs1 + hw;
                    if (o1 >= 0 && o1 < W) L1 = fmax(L1, (sv1 ? Vc1 : Vc0)[bs1] + lt_at(rvc[bs1], 1 - sv1, o1));
                }
                double best0 = NEG_INF, best1 = NEG_INF;
                int arg0 = 0, arg1 = 0;
                const int cfirst = ((b - hw + 8 * VT_CHUNK_PAD) >> 3) - VT_CHUNK_PAD;  // floor((b - hw) / 8)
                const int c_int_lo = (hw + VT_CHUNK - 1) / VT_CHUNK, c_int_hi = (n - hw - VT_CHUNK) / VT_CHUNK;  // chunks of untruncated rows
#pragma unroll
                for (int sv = 0; sv < 2; ++sv) {          // source block: 0 voiced, 1 unvoiced
                    const double* Vc = sv == 0 ? Vc0 : Vc1;
                    const int sel0 = sv, sel1 = 1 - sv;   // table for destination voiced / unvoiced
                    const int kbase = sv * n;
                    const bool g_low = gmax[sv].a < b - hw, g_high = gmax[sv].a > b + hw;
                    const double oob = gmax[sv].v + LOGTINY;
                    if (g_low) {   // lower indices than the band
                        if (oob > best0) { best0 = oob; arg0 = kbase + gmax[sv].a; }
                        if (oob > best1) { best1 = oob; arg1 = kbase + gmax[sv].a; }
                    }
                    // Groups of 4 chunks (32 bins, one warp of sources) are tested first with their own exact bound: a typical
                    // destination rejects 10 of its 13 chunks per block with 4 group tests.
                    const int cend = cfirst + NCH;
                    int c = cfirst;
#pragma unroll 1
                    while (c < cend) {
                        const int gs = (c + VT_CHUNK_PAD) >> 2;                    // group index (shifted by the padding)
                        const int c0 = (gs << 2) - VT_CHUNK_PAD;                   // its first chunk
                        const int gnext = min(c0 + 4, cend);
                        {
                            const int q4 = b + hw - VT_CHUNK * c0;                 // offset of the group's first chunk
                            if (q4 < 0) break;                                     // the whole group lies above the band
                            const double m4 = s.M4[cur][sv][gs];
                            const double g0 = m4 + s.ubr4[sel0][q4], g1 = m4 + s.ubr4[sel1][q4];
                            if (!((g0 >= L0 && g0 > best0) || (g1 >= L1 && g1 > best1))) {
                                c = gnext;
                                continue;
                            }
                        }
#pragma unroll 1
                        for (; c < gnext; ++c) {
                        const int ohi = b + hw - VT_CHUNK * c;   // offset of the chunk's first source
                        if (ohi < 0) { c = cend; break; }
                        const double m = s.M[cur][sv][c + VT_CHUNK_PAD];
                        const int kind = (c >= c_int_lo && c <= c_int_hi) ? 0 : 1;  // all 8 source rows untruncated?
                        const double bd0 = m + s.ubr[sel0][ohi], bd1 = m + s.ubr[sel1][ohi];
                        const bool need0 = (bd0 >= L0) && (bd0 > best0);
                        const bool need1 = (bd1 >= L1) && (bd1 > best1);
                        if (need0 || need1) {
                            const int b0 = VT_CHUNK * c;
                            if (kind == 0 && interior_in_smem && ohi >= VT_CHUNK - 1 && ohi < W) {
                                // all 8 sources are untruncated rows inside the band: shared-memory tables only,
                                // vector loads (V and the row variants of an aligned chunk are 16 B / 8 B aligned)
                                double x[VT_CHUNK];
#pragma unroll
                                for (int j = 0; j < VT_CHUNK; j += 2) {
                                    const double2 xx = *reinterpret_cast<const double2*>(&Vc[b0 + j]);
                                    x[j] = xx.x;
                                    x[j + 1] = xx.y;
                                }
                                const unsigned long long vars = *reinterpret_cast<const unsigned long long*>(&rvc[b0]);
#pragma unroll
                                for (int j = 0; j < VT_CHUNK; ++j) {
                                    const int var = static_cast<int>((vars >> (8 * j)) & 0xff);
                                    const double* row = &s.lt[var][0][ohi - j];
                                    if (need0) {
                                        const double cc = x[j] + row[sel0 * VT_MAX_W];
                                        if (cc > best0) { best0 = cc; arg0 = kbase + b0 + j; }
                                    }
                                    if (need1) {
                                        const double cc = x[j] + row[sel1 * VT_MAX_W];
                                        if (cc > best1) { best1 = cc; arg1 = kbase + b0 + j; }
                                    }
                                }
                            } else if (ohi >= VT_CHUNK - 1 && ohi < W) {
                                // all 8 sources inside the band, some of them truncated edge rows (their tables live in
                                // global memory, L1 resident): same unrolled form, the row pointer is selected per source.
                                // Sources outside [0, n) carry V = -inf (halo) and never win.  (Keeping the edge rows'
                                // central window in shared memory as well was tried: 108 ms instead of 79 ms per 1024
                                // clips -- the extra 55 KB per CTA costs more L1 than it saves.)
                                double x[VT_CHUNK];
#pragma unroll
                                for (int j = 0; j < VT_CHUNK; j += 2) {
                                    const double2 xx = *reinterpret_cast<const double2*>(&Vc[b0 + j]);
                                    x[j] = xx.x;
                                    x[j + 1] = xx.y;
                                }
                                const unsigned long long vars = *reinterpret_cast<const unsigned long long*>(&rvc[b0]);
#pragma unroll
                                for (int j = 0; j < VT_CHUNK; ++j) {
                                    const int var = static_cast<int>((vars >> (8 * j)) & 0xff);
                                    const bool in_smem = var < nsv;
                                    const double* row = in_smem ? &s.lt[var][0][ohi - j] : p.lt_variants + static_cast<long long>(var) * 2 * W + (ohi - j);
                                    const int pitch = in_smem ? VT_MAX_W : W;
                                    if (need0) {
                                        const double cc = x[j] + row[sel0 * pitch];
                                        if (cc > best0) { best0 = cc; arg0 = kbase + b0 + j; }
                                    }
                                    if (need1) {
                                        const double cc = x[j] + row[sel1 * pitch];
                                        if (cc > best1) { best1 = cc; arg1 = kbase + b0 + j; }
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < VT_CHUNK; ++j) {
                                    const int o = ohi - j;
                                    if (o >= 0 && o < W) {
                                        const int var = rvc[b0 + j];
                                        const double x = Vc[b0 + j];
                                        if (need0) {
                                            const double cc = x + lt_at(var, sel0, o);
                                            if (cc > best0) { best0 = cc; arg0 = kbase + b0 + j; }
                                        }
                                        if (need1) {
                                            const double cc = x + lt_at(var, sel1, o);
                                            if (cc > best1) { best1 = cc; arg1 = kbase + b0 + j; }
                                        }
                                    }
                                }
                            }
                        }
                        }
                    }
                    if (g_high) {  // higher indices than the band
                        if (oob > best0) { best0 = oob; arg0 = kbase + gmax[sv].a; }
                        if (oob > best1) { best1 = oob; arg1 = kbase + gmax[sv].a; }
                    }
                }
                vnew0 = lp_v + best0;
                vnew1 = lp_u + best1;
                unsigned short* row = bp_out + static_cast<long long>(t) * (2 * n);
                row[b] = static_cast<unsigned short>(arg0);
                row[n + b] = static_cast<unsigned short>(arg1);
                prev0 = arg0;
                prev1 = arg1;
            }
        }

        // publish V[t], its chunk maxima and the per-warp leftmost maxima
        if (live) {
            s.V[nxt][0][VT_HALO + b] = vnew0;
            s.V[nxt][1][VT_HALO + b] = vnew1;
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            VA x{live ? (v == 0 ? vnew0 : vnew1) : NEG_INF, b};
            {   // chunk maxima of V + dub (plain max: no index needed)
                double cm = x.v + dub;
                cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
                cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
                cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 4));
                if ((lane & 7) == 0) s.M[nxt][v][(b >> 3) + VT_CHUNK_PAD] = cm;
                cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
                cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
                if (lane == 0) s.M4[nxt][v][(b >> 5) + VT_CHUNK_PAD / 4] = cm;
            }
            x = butterfly_leftmost(x, 1);
            x = butterfly_leftmost(x, 2);
            x = butterfly_leftmost(x, 4);
            x = butterfly_leftmost(x, 8);
            x = butterfly_leftmost(x, 16);
            if (lane == 0) {
                s.seg_val[nxt][v][warp] = x.v;
                s.seg_arg[nxt][v][warp] = static_cast<short>(x.a);
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
    AEGIS_REQUIRE(p->half_width >= 1 && p->half_width + VT_CHUNK <= VT_HALO, "aegis_viterbi: half_width=%d unsupported (<= %d)", p->half_width, VT_HALO);
    AEGIS_REQUIRE(p->n_variants >= 1 && p->n_variants <= 255 && p->n_interior_variants >= 1, "aegis_viterbi: bad variant counts");
    AEGIS_REQUIRE(p->max_cand >= 1 && p->max_cand <= p->n_pitch_bins, "aegis_viterbi: max_cand must be 1..n_pitch_bins");
    AEGIS_REQUIRE(p->cand_bin && p->cand_prob && p->cand_count && p->voiced_prob && p->lt_variants && p->row_variant && p->freqs,
                  "aegis_viterbi: inputs missing");
    AEGIS_REQUIRE(p->backptr && p->final_value && p->states && p->f0 && p->voiced_flag, "aegis_viterbi: outputs missing");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int block = ((p->n_pitch_bins + 31) / 32) * 32;
    const bool small = block <= 448;
    void (*kern)(const aegis_viterbi_params) = nullptr;
    switch (p->half_width) {   // pYIN's band: 50 bins at 22.05 kHz / hop 512, 25 at 44.1 kHz
        case 50: kern = small ? viterbi_forward_kernel<50, 448, VT_SMALL_MINB> : viterbi_forward_kernel<50, 512, 1>; break;
        case 25: kern = small ? viterbi_forward_kernel<25, 448, VT_SMALL_MINB> : viterbi_forward_kernel<25, 512, 1>; break;
        case 12: kern = small ? viterbi_forward_kernel<12, 448, VT_SMALL_MINB> : viterbi_forward_kernel<12, 512, 1>; break;
        default: set_error("aegis_viterbi: half_width=%d has no compiled kernel (12, 25, 50)", p->half_width); return 1;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(VitSmem)));
    if (e != cudaSuccess) {
        set_error("aegis_viterbi: cannot reserve %zu B shared memory: %s", sizeof(VitSmem), cudaGetErrorString(e));
        return 2;
    }
    kern<<<p->n_clips, block, sizeof(VitSmem), st>>>(*p);
    if (int rc = check_launch("aegis_viterbi(forward)")) return rc;
    viterbi_backtrace_kernel<<<(p->n_clips + 63) / 64, 64, 0, st>>>(*p);
    return check_launch("aegis_viterbi(backtrace)");
}
