/* aegis_b200.h -- C ABI of the B200-native Aegis audio-analysis hot path (libaegis_b200.so).
 *
 * The reference (avabag01-ai/spectrogram-midi, "Aegis Engine") has no FFI: its hot path is a set
 * of Python calls into librosa/numpy.  Each entry point below names the reference call site(s) it
 * replaces (paths relative to the reference checkout).  The Python host layer
 * (spectrogram-midi_b200/) binds these with ctypes and mirrors the reference's call signatures;
 * INTEGRATION.md shows the binding a maintainer would add on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its comment says "host"; the library never
 *    allocates, frees or synchronises: outputs and workspaces are caller-provided, work is
 *    enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*).
 *  - return value 0 = enqueued; non-zero = error, text via aegis_last_error() (thread local).
 *    Nothing is thrown across the ABI.
 *  - n_fft == frame_length == 2048 everywhere (the reference never uses another size,
 *    aegis_engine.py:17); 4 <= hop <= 512 and hop % 4 == 0.
 *  - frame t of a clip covers samples [t*hop - pad, t*hop - pad + 2048) with zeros outside
 *    [0, n_samples): pad = 1024 is librosa's center=True / pad_mode='constant'.
 */
#ifndef AEGIS_B200_H
#define AEGIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEGIS_N_FFT 2048
#define AEGIS_N_BINS 1025
#define AEGIS_ABI_VERSION 2

int aegis_abi_version(void);
const char* aegis_last_error(void);
/* Number of SMs of the current device (host query helper for grid sizing in benches). */
int aegis_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  fused frame-gather + Hann window + FP32 FFT -> |X|, mel power, frame RMS
 * replaces: librosa.feature.melspectrogram -> librosa.stft   (aegis_engine.py:25,
 *           aegis_engine_financial.py:46-50) and librosa.feature.rms (aegis_engine.py:70,
 *           aegis_engine_financial.py:154).
 * Any of mag / mel / rms may be NULL (skipped).  mel_max (if mel != NULL) receives the per-clip
 * maximum of the mel power via atomicMax and must be zero-filled by the caller beforehand.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* y;            /* [n_clips][clip_stride] audio, float32 */
    int64_t clip_stride;       /* samples between clip starts */
    int64_t n_samples;         /* valid samples per clip */
    int32_t n_clips;
    int32_t hop;
    int32_t pad;               /* virtual zero samples before sample 0 (1024 = center) */
    int32_t n_frames;          /* T */
    const float* window;       /* [2048] analysis window */
    const float* twiddle;      /* [2048][2] (cos, -sin)(2 pi k / 2048) */
    float* mag;                /* [n_clips][1025][mag_row_stride]  |X| or NULL */
    int64_t mag_clip_stride;
    int32_t mag_row_stride;    /* >= T; == T gives librosa's dense [1025, T] */
    int32_t n_mels;
    float* mel;                /* [n_clips][n_mels][mel_row_stride] mel power or NULL */
    int64_t mel_clip_stride;
    int32_t mel_row_stride;
    int32_t _reserved0;
    const int32_t* mel_seg_start; /* [n_mels + 2]: FFT bins [s[j], s[j+1]) lie between mel band edges j and j+1 */
    const float* mel_rise_fall;   /* [1025][2]: weight of bin k in band seg(k) (rising side), in band seg(k)-1 (falling) */
    float* mel_max;            /* [n_clips] or NULL */
    float* rms;                /* [n_clips][rms_clip_stride] or NULL */
    int64_t rms_clip_stride;
} aegis_stft_params;

int aegis_stft_fused(const aegis_stft_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4a mel power -> dB (ref = clip max, top_db 80), rake-noise mask, spectral-flux onset envelope
 * replaces: librosa.power_to_db(S, ref=np.max) (aegis_engine.py:26),
 *           detect_rake_patterns (aegis_engine_core/vision.py:3-38),
 *           librosa.onset.onset_strength (no reference call site; BASELINE north_star).
 * s_db / rake_mask / onset_env may each be NULL.  env_minmax[2*clip+{0,1}] receives min / max of
 * the envelope (caller pre-fills with +inf / 0).  With input_is_db = 1 the kernel applies the rake
 * test to a caller-supplied dB image (the signature of detect_rake_patterns(S_dB, ...)).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* mel;          /* [n_clips][n_mels][mel_row_stride] mel power */
    int64_t mel_clip_stride;
    int32_t mel_row_stride;
    int32_t n_mels;
    int32_t n_clips;
    int32_t n_frames;
    const float* mel_max;      /* [n_clips] max mel power of each clip */
    float* s_db;               /* [n_clips][n_mels][sdb_row_stride] or NULL */
    int64_t sdb_clip_stride;
    int32_t sdb_row_stride;
    int32_t rake_min_frames;   /* int(10 / ms_per_frame) */
    int32_t rake_max_frames;   /* int(30 / ms_per_frame) */
    int32_t onset_pad;         /* lag + n_fft/(2*hop) = 3 for the defaults */
    double rake_ratio;         /* broadband_threshold_ratio */
    uint8_t* rake_mask;        /* [n_clips][n_frames] or NULL */
    float* onset_env;          /* [n_clips][n_frames] or NULL */
    float* env_minmax;         /* [n_clips][2] or NULL */
    const float* ref_power;    /* [n_clips] reference power for the dB output; NULL = mel_max (ref=np.max) */
    int32_t input_is_db;       /* 1: `mel` already holds dB values (rake mask only; mel_max unused) */
    int32_t _reserved0;
} aegis_melpost_params;

int aegis_mel_post(const aegis_melpost_params* p, void* stream);

/* K4b greedy peak picking on the normalised onset envelope (librosa.util.peak_pick semantics).
 * peaks[clip][t] = 1 at onset frames; n_peaks[clip] = count. */
typedef struct {
    const float* onset_env;    /* [n_clips][n_frames] */
    const float* env_minmax;   /* [n_clips][2] */
    int32_t n_clips;
    int32_t n_frames;
    int32_t pre_max, post_max, pre_avg, post_avg, wait;
    int32_t normalize;
    double delta;
    uint8_t* cand;             /* workspace [n_clips][n_frames] */
    uint8_t* peaks;            /* [n_clips][n_frames] */
    int32_t* n_peaks;          /* [n_clips] */
} aegis_peaks_params;

int aegis_onset_peaks(const aegis_peaks_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  batched YIN: FFT difference function -> CMND -> parabolic interpolation -> trough /
 *     beta-threshold / Boltzmann-prior candidate probabilities -> sparse pitch-bin observations
 * replaces: librosa.pyin stages 1-9 (call sites aegis_engine.py:63,67,190,216;
 *           aegis_engine_core/worker.py:9-15; aegis_engine_financial.py:63-69).
 * Per frame: cand_count[f] candidates (bin ascending, unique), voiced_prob[f] (float64).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* y;
    int64_t clip_stride;
    int64_t n_samples;
    int32_t n_clips;
    int32_t hop;
    int32_t pad;
    int32_t n_frames;
    const float* twiddle;      /* [2048][2] */
    double sr;
    double fmin;
    int32_t min_period;
    int32_t max_period;
    int32_t n_pitch_bins;
    int32_t bins_per_semitone;
    int32_t n_thresholds;      /* 100 */
    int32_t max_cand;          /* capacity per frame of cand_bin / cand_prob */
    const double* thresholds;  /* [n_thresholds] upper edges */
    const double* beta_probs;  /* [n_thresholds] */
    const double* beta_cumsum; /* [n_thresholds + 1] */
    const double* boltz_fact;  /* [max_troughs + 1] */
    const double* boltz_exp;   /* [max_troughs + 1] */
    double no_trough_prob;
    uint16_t* cand_bin;        /* [n_clips * n_frames][max_cand] */
    double* cand_prob;         /* [n_clips * n_frames][max_cand] (librosa keeps the CMND in float64) */
    int32_t* cand_count;       /* [n_clips * n_frames] */
    double* voiced_prob;       /* [n_clips * n_frames] */
    int32_t* overflow;         /* [1]: set non-zero if any frame had more than max_cand candidates */
    double* cmnd_out;          /* optional [n_clips * n_frames][max_period - min_period + 1] CMND curves, or NULL */
    float* block_sums;         /* optional workspace of aegis_yin_workspace_bytes(...) bytes (hop 512 only): with it the
                                  autocorrelation block sums and the per-frame stage run as two kernels (faster: each at
                                  its own register budget); NULL = one fused kernel, bit-identical results */
} aegis_yin_params;

int aegis_yin_candidates(const aegis_yin_params* p, void* stream);
long long aegis_yin_workspace_bytes(int n_clips, int n_frames, int max_period);

/* ---------------------------------------------------------------------------------------------
 * K3  pitch/voicing HMM Viterbi decode (2*n_pitch_bins states, banded transitions, exact
 *     out-of-band log(tiny) competitors, float64, first-index tie-breaking)
 * replaces: librosa.sequence.viterbi inside librosa.pyin (same call sites as K2).
 * One chain per clip.  backptr is workspace: [n_clips][n_frames][2*n_pitch_bins] uint16.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_clips;
    int32_t n_frames;
    int32_t n_pitch_bins;
    int32_t half_width;
    int32_t n_variants;
    int32_t n_interior_variants;
    int32_t max_cand;
    int32_t _reserved0;
    const uint16_t* cand_bin;
    const double* cand_prob;
    const int32_t* cand_count;
    const double* voiced_prob;
    const double* lt_variants; /* [n_variants][2][2*half_width+1] log transition bands */
    const int32_t* row_variant;/* [n_pitch_bins] */
    const double* freqs;       /* [n_pitch_bins] */
    double log_tiny;
    double log_init_unvoiced;
    double fill_value;         /* f0 for unvoiced frames (NaN for librosa's default) */
    uint16_t* backptr;         /* workspace */
    double* final_value;       /* workspace [n_clips][2*n_pitch_bins] */
    uint16_t* states;          /* [n_clips][n_frames] */
    double* f0;                /* [n_clips][n_frames] */
    uint8_t* voiced_flag;      /* [n_clips][n_frames] */
} aegis_viterbi_params;

int aegis_viterbi(const aegis_viterbi_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  batched float64 trend filters on f0 series (NaN = unvoiced)
 * replaces: FinancialNoiseFilters.{savitzky_golay,kalman_filter,holt_winters} and
 *           multi_filter_consensus (aegis_engine_core_v2/financial_filters.py:25-141,256-298);
 *           FinancialPitchAnalyzer.{simple_moving_average,exponential_moving_average,
 *           bollinger_bands,macd} (aegis_engine_core_v2/financial_analysis.py:45-146,203-226).
 * All series are [n_series][n] float64, row stride = n.  Any output may be NULL.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const double* x;
    int32_t n_series;
    int32_t n;
    const double* savgol_coeffs; /* [savgol_window] correlation weights */
    int32_t savgol_window;
    int32_t sma_window;
    int32_t ema_span;
    int32_t boll_window;
    int32_t macd_fast, macd_slow, macd_signal;
    int32_t consensus_mask;    /* filters that vote in the consensus: bit 0 savgol, 1 kalman, 2 holt; 0 = all three
                                  (multi_filter_consensus(data, filters=[...]), financial_filters.py:256-298) */
    double kalman_q, kalman_r;
    double holt_alpha, holt_beta;
    double boll_num_std;
    double* compact;           /* workspace [n_series][n] */
    double* scratch;           /* workspace [n_series][n] */
    double* savgol;
    double* kalman;
    double* holt;
    double* consensus;
    double* consensus_conf;
    double* sma;
    double* ema;
    double* boll_ma;
    double* boll_upper;
    double* boll_lower;
    double* macd_line;
    double* macd_sig;
    double* macd_hist;
} aegis_trend_params;

int aegis_trend_filters(const aegis_trend_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  electric-guitar filters on the perception outputs (v2 engine)
 * replaces: apply_guitar_filters (aegis_engine_core_v2/guitar_specific.py:240-277; caller
 *           aegis_engine_financial.py:132-147) = filter_subharmonic_noise (:24-61),
 *           detect_rake_enhanced (:112-151), detect_palm_mute (:63-110), classify_distortion_level
 *           (:209-233).
 * Any output may be NULL (skipped).  s_db is only needed for rake_out / mute_out / distortion,
 * f0 / voiced only for f0_out / voiced_out.  distortion needs the workspace dist_work of
 * 2 * n_clips * aegis_guitar_blocks(n_frames) doubles.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* s_db;         /* [n_clips][n_mels][sdb_row_stride] mel spectrogram in dB */
    int64_t sdb_clip_stride;
    int32_t sdb_row_stride;
    int32_t n_mels;
    int32_t n_clips;
    int32_t n_frames;
    int32_t mute_max_frames;   /* int(50 / ms_per_frame), <= 30 */
    int32_t rake_frames;       /* int(30 / ms_per_frame), <= 30 */
    double fmin_hz;            /* 82.4 in the reference */
    const double* f0;          /* [n_clips][n_frames], NaN = unvoiced */
    const uint8_t* voiced;     /* [n_clips][n_frames] */
    const uint8_t* rake_in;    /* [n_clips][n_frames] basic rake mask (NULL = all zero) */
    double* f0_out;            /* [n_clips][n_frames] */
    uint8_t* voiced_out;       /* [n_clips][n_frames] */
    uint8_t* rake_out;         /* [n_clips][n_frames] enhanced rake mask */
    uint8_t* mute_out;         /* [n_clips][n_frames] palm-mute mask */
    int32_t* distortion;       /* [n_clips]: 0 clean, 1 light, 2 heavy */
    double* dist_work;         /* workspace, see above */
} aegis_guitar_params;

int aegis_guitar_filters(const aegis_guitar_params* p, void* stream);
/* column blocks per clip of the K6 launch (sizes dist_work) */
int aegis_guitar_blocks(int n_frames);

/* ---------------------------------------------------------------------------------------------
 * K7  frame -> note-event state machine (v1 logic filter)
 * replaces: get_midi_events + detect_articulations (aegis_engine_core/midi_logic.py:6-148; caller
 *           aegis_engine.py:88-96 extract_events) with the raw-f0 branch the reference always takes
 *           (:43-49: librosa.util.softmask has no `margin` keyword).
 * One record per note event, in time order; n_events[clip] is the number found (if it exceeds
 * max_events the list is truncated).  The MIDI note of a frame is note_lut[pitch_index[frame]] when
 * both are given (the host computes round(hz_to_midi(f)) with numpy for every distinct f0 value, so
 * exact quarter-tone ties round as in the reference), else rint(12 log2(f/440) + 69) on the device.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t note;              /* MIDI note number */
    int32_t start;             /* first frame */
    int32_t end;               /* last frame (inclusive) */
    int32_t velocity;          /* int(clip((rms_dB + 80) * 1.5, 0, 127)) at the start frame */
    float rms_energy;          /* rms_dB at the start frame */
    uint8_t track;             /* 1 = main (confidence >= threshold), 0 = safe */
    uint8_t technique;         /* 0 none, 1 vibrato, 2 bend, 3 slide, 4 hammer_on, 5 pull_off */
    uint8_t _pad[2];
    double confidence;         /* voiced probability at the start frame */
    double slope;              /* semitones per frame (0 when no technique / hammer_on / pull_off) */
} aegis_note_event;

typedef struct {
    const uint8_t* rake_mask;  /* [n_clips][n_frames] */
    const double* f0;          /* [n_clips][n_frames] Hz; <= 0 or NaN = no pitch */
    const uint8_t* voiced_flag;/* [n_clips][n_frames] */
    const double* voiced_prob; /* [n_clips][n_frames] */
    const float* rms;          /* [n_clips][rms_clip_stride] */
    int64_t rms_clip_stride;
    const uint16_t* pitch_index; /* [n_clips][n_frames] index into note_lut, or NULL */
    const int16_t* note_lut;   /* [n_lut] MIDI note of every distinct f0 value, or NULL */
    int32_t n_lut;
    int32_t n_clips;
    int32_t n_frames;
    int32_t hop;
    double sr;
    double confidence_threshold;
    float noise_gate_db;       /* -40 in the reference */
    int32_t min_note_frames;   /* int(min_note_duration_ms / 1000 * sr / hop) */
    int32_t sustain_frames;    /* int(sustain_ms / 1000 * sr / hop) */
    int32_t max_events;        /* capacity of events per clip */
    aegis_note_event* events;  /* [n_clips][max_events], buffer of aegis_note_events_bytes(...) bytes */
    int32_t* n_events;         /* [n_clips] */
} aegis_notes_params;

int aegis_note_events(const aegis_notes_params* p, void* stream);
/* bytes to provide at `events`: n_clips * max_events records followed by the kernel's per-frame scratch */
long long aegis_note_events_bytes(int n_clips, int n_frames, int max_events);

/* ---------------------------------------------------------------------------------------------
 * K8  v2 ("financial") logic filter: frames -> note events, with the RSI ghost-note filter and the key /
 *     out-of-scale / chord-context pass
 * replaces: get_midi_events_financial with use_financial=True (aegis_engine_core_v2/midi_logic_financial.py:117-388,
 *           called from aegis_engine_financial.py:160-171), adaptive_confidence_threshold (:77-114),
 *           FinancialPitchAnalyzer.{detect_articulation_bollinger,detect_slides_macd,rsi,filter_ghost_notes_rsi}
 *           (aegis_engine_core_v2/financial_analysis.py:148-196,228-271,277-364) and HarmonicAnalyzer.{detect_key,
 *           filter_out_of_scale_notes,analyze_chord_progression,adaptive_filter_by_context}
 *           (aegis_engine_core_v2/harmonic_analysis.py:46-283).
 * Three calls: aegis_fin_prepare writes f0_clean (f0 where voiced, else NaN), its MIDI-number series and the rms dB
 * scratch; the caller runs aegis_trend_filters on f0_clean (consensus, Bollinger bands window 10 / 2 sigma) and on
 * the semitone series (MACD 5/20/9); aegis_fin_events turns the frames into events.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t note;              /* MIDI note number */
    int32_t start;             /* first frame */
    int32_t end;               /* last frame (inclusive) */
    int32_t velocity;          /* 0..127 */
    uint8_t track;             /* 1 = main (confidence >= threshold), 0 = safe */
    uint8_t technique;         /* = financial_artic: 0 None, 1 normal, 2 bend, 3 vibrato, 4 noise */
    uint8_t slide;             /* financial_slide at the start frame: 0 None, 1 normal, 2 slide_up, 3 slide_down */
    int8_t harmonic_valid;     /* -1 key absent (nothing was out of scale), 1 True */
    uint8_t _pad[4];
    double confidence;         /* 0.5 voiced_prob + 0.5 Bollinger confidence at the start frame, after the context pass */
} aegis_fin_event;

typedef struct {
    const uint8_t* rake_mask;  /* [n_clips][n_frames] */
    const double* f0;          /* [n_clips][n_frames] Hz, NaN = no pitch */
    const uint8_t* voiced_flag;/* [n_clips][n_frames] */
    const double* voiced_prob; /* [n_clips][n_frames] */
    const float* rms;          /* [n_clips][rms_clip_stride] */
    int64_t rms_clip_stride;
    int32_t n_clips;
    int32_t n_frames;
    int32_t hop;
    int32_t _reserved0;
    double sr;
    double* f0_clean;          /* [n_clips][n_frames] written by aegis_fin_prepare */
    double* semitones;         /* [n_clips][n_frames] written by aegis_fin_prepare */
    const double* trend;       /* consensus of aegis_trend_filters(f0_clean) */
    const double* boll_upper;  /* bands of aegis_trend_filters(f0_clean), window 10, 2 sigma */
    const double* boll_lower;
    const double* macd_line;   /* aegis_trend_filters(semitones), 5 / 20 / 9 */
    const double* macd_hist;
    double confidence_threshold; /* NaN = adaptive: clip(mean - std of the positive confidences, 0.3, 0.8) */
    double slide_threshold;    /* 0.3 in the reference */
    double rsi_threshold;      /* 70 in the reference */
    float noise_gate_db;       /* -40 in the reference */
    int32_t min_note_frames;   /* int(min_note_duration_ms / 1000 * sr / hop) */
    int32_t sustain_frames;    /* int(sustain_ms / 1000 * sr / hop) */
    int32_t use_harmonic_filter;
    int32_t harmonic_tolerance;/* semitones a note may lie off the detected scale (reference default 1) */
    int32_t max_events;        /* capacity per clip; n_frames / (min_note_frames + 1) + 1 always suffices */
    aegis_fin_event* events;   /* [n_clips][max_events] */
    int32_t* n_events;         /* [n_clips]; a value > max_events means the clip overflowed and its records are invalid */
    double* threshold_out;     /* [n_clips] threshold in force, or NULL */
    int32_t* key_out;          /* [n_clips] root | mode << 8 (mode 0 major, 1 minor, 2 blues) when notes were removed, else -1; or NULL */
    double* key_confidence_out;/* [n_clips] or NULL */
    void* scratch;             /* aegis_fin_scratch_bytes(n_clips, n_frames) bytes; shared by the two calls */
} aegis_fin_params;

int aegis_fin_prepare(const aegis_fin_params* p, void* stream);
int aegis_fin_events(const aegis_fin_params* p, void* stream);
long long aegis_fin_scratch_bytes(int n_clips, int n_frames);

/* ---------------------------------------------------------------------------------------------
 * K9  audio ingest: PCM -> float32, channel mix-down, polyphase rate conversion
 * replaces: librosa.load(path, sr=engine rate, res_type='polyphase') after the file read: soundfile's
 *           int16 -> float32 / 32768, librosa.to_mono, librosa.resample(res_type='polyphase')
 *           = scipy.signal.resample_poly(y, up, down).  NOTE: the reference's own call sites (aegis_engine.py:24,
 *           aegis_engine_financial.py:45) pass no res_type, so librosa converts with its default 'soxr_hq' (libsoxr):
 *           that resampler is NOT what this entry point computes (libsoxr is absent from the image and its
 *           arithmetic cannot be restated); the Python layer warns (ResampleDivergenceWarning) whenever it has to
 *           substitute.  Parity with the reference holds for files already at the engine rate, or when both sides
 *           use res_type='polyphase'.  `taps` is scipy.signal.firwin(2 * 10 * max(up, down) + 1,
 *           1 / max(up, down), window=('kaiser', 5.0)) in float32 times `up`; n_pre_pad = down - half_len % down,
 *           n_pre_remove = (half_len + n_pre_pad) / down, n_out = ceil(n_in * up / down)  (scipy/signal/_signaltools.py).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const void* x;             /* [n_clips][in_clip_stride] float32 or int16, channels interleaved */
    int64_t in_clip_stride;    /* elements */
    int64_t n_in;              /* sample frames per clip */
    int32_t in_format;         /* 0 float32, 1 int16 (scaled by 1 / 32768) */
    int32_t n_channels;        /* 1..8, averaged */
    int32_t n_clips;
    int32_t up, down;          /* reduced ratio: target_sr / gcd, orig_sr / gcd */
    int32_t n_taps;
    const float* taps;         /* [n_taps] */
    int32_t n_pre_pad;
    int32_t n_pre_remove;
    float* out;                /* [n_clips][out_clip_stride] */
    int64_t out_clip_stride;
    int64_t n_out;
} aegis_resample_params;

int aegis_resample_poly(const aegis_resample_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Event serialisation (HOST code, HOST pointers): Standard MIDI File bytes and tablature positions
 * replaces: the MIDI writer of AegisEngine.extract_events (aegis_engine.py:98-179, mido.MidiFile.save), the MIDI
 *           writer of AegisFinancialEngine.audio_to_midi_financial (aegis_engine_financial.py:185-246) and
 *           generate_tabs (aegis_engine_core/tabs.py:1-40).
 * The writers return the file size in bytes (and fill `out` when `capacity` suffices), or -1 on error.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t hop;
    int32_t midi_program;      /* v1: program change sent on both tracks (reference default 27) */
    double sr;
    double vibrato_rate;       /* v1: Hz, reference default 5.0 */
    double vibrato_depth;      /* v1: fraction of the pitch-wheel range, reference default 0.3 */
} aegis_smf_options;

long long aegis_smf_write_v1(const aegis_note_event* events, int32_t n_events, const aegis_smf_options* opt,
                             uint8_t* out, long long capacity);
long long aegis_smf_write_v2(const aegis_fin_event* events, int32_t n_events, const aegis_smf_options* opt,
                             uint8_t* out, long long capacity);
/* string_out[i] in 1..6 (0: no string can play the note, the reference skips it), fret_out[i] in 0..24 */
int aegis_tabs(const int32_t* notes, int32_t n, int32_t* string_out, int32_t* fret_out);
/* export_musicxml (aegis_engine_core/tabs.py:42-112) as bytes: techniques 0 none, 1 vibrato, 2 bend, 3 slide (may be
 * NULL).  Returns the document size (fills `out` when `capacity` suffices), or -1 on error. */
long long aegis_musicxml_write(const int32_t* notes, const int32_t* strings, const int32_t* frets,
                               const uint8_t* techniques, int32_t n, uint8_t* out, long long capacity);

/* ---------------------------------------------------------------------------------------------
 * Corpus synthesis on device (Karplus-Strong plucks + noise rakes, generate_test_signal.py:5-53)
 * One event per note/rake; events of a clip do not overlap.  out must be zero-filled.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    float* out;                /* [n_clips][clip_stride] */
    int64_t clip_stride;
    int32_t n_events;
    int32_t _reserved0;
    const int32_t* ev_clip;    /* [n_events] */
    const int32_t* ev_start;   /* [n_events] first sample */
    const int32_t* ev_len;     /* [n_events] samples */
    const int32_t* ev_period;  /* [n_events] KS delay length; 0 = noise rake */
    const float* ev_amp;       /* [n_events] */
    const uint32_t* ev_seed;   /* [n_events] */
    float decay;
    float _reserved1;
} aegis_synth_params;

int aegis_synth_ks(const aegis_synth_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AEGIS_B200_H */
