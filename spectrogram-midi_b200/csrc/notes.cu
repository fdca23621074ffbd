// K7: frame -> note-event state machine of the v1 logic filter.
//
// Replaces get_midi_events / detect_articulations (aegis_engine_core/midi_logic.py:6-148, called from
// aegis_engine.py:88-96).  The reference walks the frames in Python; its state machine is equivalent to: events =
// maximal runs of frames with the same MIDI note among the frames that are voiced, above the noise gate, pitched and
// not rake noise (:63-105); drop events shorter than min_note_duration (:108); merge neighbours of equal note whose
// gap is <= sustain_frames when the first has no technique (:111-123); mark hammer-on / pull-off pairs (:126-146).
//
// Kernel 1 (frame parallel): rms -> dB relative to the clip maximum (librosa.amplitude_to_db(ref=np.max), float32
// arithmetic as in numpy; log10 is evaluated in double and rounded once, i.e. correctly rounded float32) and the
// per-frame MIDI note (-1 = inactive) and the frame's continuous MIDI pitch 12 log2(f / 440) + 69 (the double-precision
// log2 is by far the longest step of the walk: it is taken here, frame parallel, not in the sequential kernel).
// Kernel 2: one warp per clip walks its frames once (rows staged in shared memory), keeps least-squares sums of the
// running event, finishes events (slope, vibrato range), and applies the duration filter, the merge and the hammer-on /
// pull-off rule in streaming form: an event is written once the next surviving event is known.  A clip is a sequential
// chain of ~T steps of a few dozen instructions, all clips in parallel.  (Round 1 and early round 2 ran one THREAD per
// clip: 32 unrelated state machines per warp serialised on each other's branches, 2.2 ms for 1024 clips.)
#include <cfloat>
#include <cmath>
#include "common.cuh"
#include "notes_common.cuh"

namespace aegis {

constexpr int NT_THREADS = 256;

// librosa.hz_to_midi, the expression both kernels used before the log2 was hoisted (same contraction: one definition)
__device__ __forceinline__ double midi_pitch(double f) { return 12.0 * (log2(f) - log2(440.0)) + 69.0; }

// per frame: rms dB and MIDI note (or -1)
__global__ void __launch_bounds__(NT_THREADS)
notes_frames_kernel(const aegis_notes_params p, const float* __restrict__ rms_max, float* __restrict__ rms_db, short* __restrict__ note,
                    double* __restrict__ pitch) {
    const int clip = blockIdx.y;
    const int t = blockIdx.x * NT_THREADS + threadIdx.x;
    if (t >= p.n_frames) return;
    const long long i = static_cast<long long>(clip) * p.n_frames + t;
    const float e = rms_db_f32(p.rms[static_cast<long long>(clip) * p.rms_clip_stride + t], rms_max[clip]);
    rms_db[i] = e;
    const double f = p.f0[i];
    const bool active = p.voiced_flag[i] != 0 && !(e < p.noise_gate_db) && f > 0.0 && p.rake_mask[i] == 0;
    int n = -1;
    if (active) {
        if (p.note_lut != nullptr && p.pitch_index != nullptr) {
            const int idx = p.pitch_index[i];
            n = idx < p.n_lut ? p.note_lut[idx] : -1;
        } else {
            n = static_cast<int>(rint(12.0 * (log2(f) - log2(440.0)) + 69.0));  // round half to even, as Python's round()
        }
    }
    note[i] = static_cast<short>(n);
    pitch[i] = n >= 0 ? midi_pitch(f) : 0.0;
}

struct Ev {
    int note, start, end, velocity;
    float energy;
    int track, technique;
    double confidence, slope;
};

__device__ __forceinline__ void write_event(aegis_note_event* dst, const Ev& e) {
    dst->note = e.note;
    dst->start = e.start;
    dst->end = e.end;
    dst->velocity = e.velocity;
    dst->rms_energy = e.energy;
    dst->track = static_cast<uint8_t>(e.track);
    dst->technique = static_cast<uint8_t>(e.technique);
    dst->_pad[0] = dst->_pad[1] = 0;
    dst->confidence = e.confidence;
    dst->slope = e.slope;
}

// One warp per clip.  The walk is a sequential chain (the least-squares sums are accumulated in frame order, as the
// oracle does), so every lane runs it redundantly on the same data -- control flow stays warp-uniform, no lane waits for
// another clip's branch -- and the lanes split the one data-parallel piece, the residual range of a finished run (min /
// max are order independent).  Lane 0 writes.  STAGED: the clip's note / pitch rows are first copied to shared memory
// with coalesced loads, so the chain runs on shared-memory latency; clips too long for that walk the global rows.
constexpr int NE_STAGE_MAX_FRAMES = 20000;   // 10 B per frame of dynamic shared memory

__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <bool STAGED>
__global__ void __launch_bounds__(32)
notes_events_kernel(const aegis_notes_params p, const float* __restrict__ rms_db, const short* __restrict__ note,
                    const double* __restrict__ pitch) {
    extern __shared__ __align__(16) unsigned char ne_smem[];
    const int clip = blockIdx.x;
    const int lane = threadIdx.x;
    const int T = p.n_frames;
    const long long base = static_cast<long long>(clip) * T;
    const short* nt = note + base;
    const double* yp = pitch + base;
    if (STAGED) {
        double* yp_s = reinterpret_cast<double*>(ne_smem);
        short* nt_s = reinterpret_cast<short*>(yp_s + T);
        for (int t = lane; t < T; t += 32) {
            yp_s[t] = yp[t];
            nt_s[t] = nt[t];
        }
        __syncwarp();
        nt = nt_s;
        yp = yp_s;
    }
    aegis_note_event* out = p.events + static_cast<long long>(clip) * p.max_events;
    const double ms_per_frame = (static_cast<double>(p.hop) / p.sr) * 1000;
    int n_out = 0;
    bool have_pending = false, have_last = false;
    Ev pending{}, last{};   // pending: survived the duration filter, may still absorb merges; last: previous event written

    auto emit = [&](Ev e) {   // hammer-on / pull-off against the previously written event (:126-146), then write
        if (have_last) {
            const double gap_ms = (e.start - last.end) * ms_per_frame;
            if (gap_ms < 30) {
                const int dp = e.note - last.note;
                const double v_ratio = static_cast<double>(e.velocity) / static_cast<double>(max(last.velocity, 1));
                const float denom = fmaxf(last.energy, -80.0f);
                const float e_ratio = __fdiv_rn(e.energy, denom);
                const bool weak = v_ratio < 0.7 || e_ratio < 0.8f;
                if (dp > 0 && dp <= 2 && weak) { e.technique = 4; e.slope = 0.0; }
                else if (dp >= -2 && dp < 0 && weak) { e.technique = 5; e.slope = 0.0; }
            }
        }
        if (lane == 0 && n_out < p.max_events) write_event(out + n_out, e);
        ++n_out;
        // the rule reads note / velocity / energy / end of the predecessor: none of them is changed by the rule itself
        last = e;
        have_last = true;
    };
    auto finish = [&](Ev e, double sy, double sxy) {   // articulation of a completed run (:6-30), filter, merge
        const int n = e.end - e.start + 1;
        e.technique = 0;
        e.slope = 0.0;
        if (n >= 3) {
            // least squares of the MIDI pitch against the frame index 0..n-1
            const double xm = 0.5 * (n - 1), sxx = static_cast<double>(n) * (static_cast<double>(n) * n - 1.0) / 12.0;
            const double ym = sy / n;
            const double slope = (sxy - xm * sy) / sxx;
            const double icpt = ym - slope * xm;
            double rmin = DBL_MAX, rmax = -DBL_MAX;
            for (int t = e.start + lane; t <= e.end; t += 32) {
                const double r = yp[t] - (slope * (t - e.start) + icpt);
                rmin = fmin(rmin, r);
                rmax = fmax(rmax, r);
            }
            rmin = warp_min_f64(rmin);
            rmax = warp_max_f64(rmax);
            if (rmax - rmin > 0.3) { e.technique = 1; e.slope = slope; }
            else if (slope > 0.05) { e.technique = 2; e.slope = slope; }
            else if (fabs(slope) > 0.02) { e.technique = 3; e.slope = slope; }
        }
        if (e.end - e.start < p.min_note_frames) return;          // :108
        if (have_pending) {
            if (e.note == pending.note && (e.start - pending.end) <= p.sustain_frames && pending.technique == 0) {
                pending.end = e.end;                               // :117-118
                return;
            }
            emit(pending);
        }
        pending = e;
        have_pending = true;
    };

    Ev cur{};
    bool open = false;
    double sy = 0.0, sxy = 0.0;
    for (int t = 0; t < T; ++t) {
        const int n = nt[t];
        if (open && n != cur.note) {
            finish(cur, sy, sxy);
            open = false;
        }
        if (n >= 0) {
            const double y = yp[t];
            if (!open) {
                const float energy = rms_db[base + t];
                const double conf = p.voiced_prob[base + t];
                cur.note = n;
                cur.start = t;
                cur.energy = energy;
                cur.confidence = conf;
                cur.velocity = velocity_from_db(energy);
                cur.track = conf >= p.confidence_threshold ? 1 : 0;
                sy = 0.0;
                sxy = 0.0;
                open = true;
            }
            cur.end = t;
            sy += y;
            sxy += y * (t - cur.start);
        }
    }
    if (open) finish(cur, sy, sxy);
    if (have_pending) emit(pending);
    if (lane == 0) p.n_events[clip] = n_out;
}

}  // namespace aegis

extern "C" int aegis_note_events(const aegis_notes_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "aegis_note_events: null params");
    AEGIS_REQUIRE(p->n_clips >= 0 && p->n_frames >= 0 && p->max_events >= 0, "aegis_note_events: negative size");
    AEGIS_REQUIRE(p->rake_mask && p->f0 && p->voiced_flag && p->voiced_prob && p->rms, "aegis_note_events: inputs missing");
    AEGIS_REQUIRE(p->events && p->n_events, "aegis_note_events: outputs missing");
    AEGIS_REQUIRE(p->rms_clip_stride >= p->n_frames && p->hop > 0 && p->sr > 0, "aegis_note_events: bad rms stride / hop / sr");
    AEGIS_REQUIRE((p->note_lut == nullptr) == (p->pitch_index == nullptr), "aegis_note_events: note_lut and pitch_index go together");
    if (p->n_clips == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // scratch: rms_max [n_clips] f32, rms_db [n_clips*T] f32, note [n_clips*T] i16 -- carved from the events buffer? No:
    // the library never allocates, so the scratch lives behind the event records (see the size query below)
    const long long frames = static_cast<long long>(p->n_clips) * p->n_frames;
    unsigned char* scratch = reinterpret_cast<unsigned char*>(p->events + static_cast<long long>(p->n_clips) * p->max_events);
    float* rms_max = reinterpret_cast<float*>(scratch);
    float* rms_db = rms_max + ((p->n_clips + 3) / 4) * 4;
    short* note = reinterpret_cast<short*>(rms_db + frames);
    double* pitch = reinterpret_cast<double*>(scratch + (((reinterpret_cast<unsigned char*>(note + frames) - scratch) + 15) / 16) * 16);
    if (p->n_frames > 0) {
        rms_max_kernel<NT_THREADS><<<p->n_clips, NT_THREADS, 0, st>>>(p->rms, p->rms_clip_stride, p->n_frames, rms_max);
        if (int rc = check_launch("aegis_note_events(rms max)")) return rc;
        notes_frames_kernel<<<dim3((p->n_frames + NT_THREADS - 1) / NT_THREADS, p->n_clips), NT_THREADS, 0, st>>>(*p, rms_max, rms_db, note, pitch);
        if (int rc = check_launch("aegis_note_events(frames)")) return rc;
    }
    if (p->n_frames <= NE_STAGE_MAX_FRAMES) {
        const size_t smem = static_cast<size_t>(p->n_frames) * 10 + 16;
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(notes_events_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            AEGIS_REQUIRE(e == cudaSuccess, "aegis_note_events: %s", cudaGetErrorString(e));
        }
        notes_events_kernel<true><<<p->n_clips, 32, smem, st>>>(*p, rms_db, note, pitch);
    } else {
        notes_events_kernel<false><<<p->n_clips, 32, 0, st>>>(*p, rms_db, note, pitch);
    }
    return check_launch("aegis_note_events(events)");
}

// bytes the caller must provide behind `events`: the records plus the per-frame scratch
extern "C" long long aegis_note_events_bytes(int n_clips, int n_frames, int max_events) {
    const long long frames = static_cast<long long>(n_clips) * n_frames;
    return static_cast<long long>(n_clips) * max_events * static_cast<long long>(sizeof(aegis_note_event)) +
           ((n_clips + 3) / 4) * 4 * 4LL + frames * 4 + frames * 2 + 32 + frames * 8 + 16;
}
