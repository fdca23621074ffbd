// Event serialisation after the path (host code, no kernels): note events -> Standard MIDI File bytes, and the
// string / fret assignment of the tablature.  HOST pointers in, HOST bytes out.
//
// Replaces, for batches where a Python loop per clip is the bottleneck:
//   * the MIDI writer of AegisEngine.extract_events (aegis_engine.py:98-179): two tracks (main / safe), program
//     change, note on / off with hammer-on / pull-off velocity scaling, 15-point bend curves and 10..20-point vibrato
//     curves as pitch-wheel messages, stable sort by tick, per-track delta times, saved with mido's MidiFile.save;
//   * the MIDI writer of AegisFinancialEngine.audio_to_midi_financial (aegis_engine_financial.py:185-246): two named
//     tracks, note on / off only, ticks from milliseconds at 120 BPM;
//   * generate_tabs (aegis_engine_core/tabs.py:1-40).
//
// The byte layout is what mido 1.x writes (mido is NOT in this image, so this part is restated from the SMF 1.0
// specification and mido's documented behaviour and is unpinned against mido itself -- tests parse the bytes back with
// an independent reader): 'MThd' 6, format 1, n tracks, 480 ticks per beat; per track 'MTrk' length, messages as
// variable-length delta + status + data with running status between consecutive channel messages of equal status
// (meta events reset it), and an end-of-track meta event appended with delta 0.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include "common.cuh"

namespace aegis {
namespace {

struct Msg {
    long long tick;
    int kind;    // 0 note_on, 1 note_off, 2 pitchwheel
    int track;   // 1 main, 0 safe
    int a, b;    // note, velocity | pitch, -
};

struct TrackBytes {
    std::vector<uint8_t> data;
    int running = -1;
    void varlen(unsigned long long v) {
        uint8_t tmp[10];
        int n = 0;
        tmp[n++] = static_cast<uint8_t>(v & 0x7f);
        while (v >>= 7) tmp[n++] = static_cast<uint8_t>((v & 0x7f) | 0x80);
        while (n) data.push_back(tmp[--n]);
    }
    void channel(unsigned long long delta, int status, int d1, int d2 = -1) {
        varlen(delta);
        if (status != running) data.push_back(static_cast<uint8_t>(status));
        data.push_back(static_cast<uint8_t>(d1));
        if (d2 >= 0) data.push_back(static_cast<uint8_t>(d2));
        running = status;
    }
    void meta(unsigned long long delta, int type, const char* text, size_t len) {
        varlen(delta);
        data.push_back(0xff);
        data.push_back(static_cast<uint8_t>(type));
        varlen(len);
        data.insert(data.end(), text, text + len);
        running = -1;
    }
};

void put_be(std::vector<uint8_t>& v, unsigned long long x, int bytes) {
    for (int i = bytes - 1; i >= 0; --i) v.push_back(static_cast<uint8_t>((x >> (8 * i)) & 0xff));
}

long long finish_file(TrackBytes* tracks, int n_tracks, uint8_t* out, long long capacity) {
    std::vector<uint8_t> file;
    file.insert(file.end(), {'M', 'T', 'h', 'd'});
    put_be(file, 6, 4);
    put_be(file, 1, 2);
    put_be(file, static_cast<unsigned>(n_tracks), 2);
    put_be(file, 480, 2);
    for (int t = 0; t < n_tracks; ++t) {
        tracks[t].meta(0, 0x2f, "", 0);   // end of track
        file.insert(file.end(), {'M', 'T', 'r', 'k'});
        put_be(file, tracks[t].data.size(), 4);
        file.insert(file.end(), tracks[t].data.begin(), tracks[t].data.end());
    }
    const long long need = static_cast<long long>(file.size());
    if (out != nullptr && capacity >= need) std::memcpy(out, file.data(), file.size());
    return need;
}

bool note_ok(int note, int velocity) { return note >= 0 && note <= 127 && velocity >= 0 && velocity <= 127; }

}  // namespace
}  // namespace aegis

extern "C" long long aegis_smf_write_v1(const aegis_note_event* events, int32_t n_events, const aegis_smf_options* opt,
                                        uint8_t* out, long long capacity) {
    using namespace aegis;
    if (n_events < 0 || (n_events > 0 && events == nullptr) || opt == nullptr || opt->hop <= 0 || !(opt->sr > 0) ||
        opt->midi_program < 0 || opt->midi_program > 127) {
        set_error("aegis_smf_write_v1: bad arguments");
        return -1;
    }
    const double secs_per_frame = static_cast<double>(opt->hop) / opt->sr;
    // mido.second2tick(1.0, ticks_per_beat=480, tempo=500000).  mido is unpinned in the reference's requirements.txt:
    // mido >= 1.3 returns the rounded int 960, mido 1.2.x the float 1.0 / (500000 * 1e-6 / 480), which evaluates to
    // exactly 960.0 in binary64 (500000 * 1e-6 == 0.5) -- the same constant either way, so int(start * spf * tps) cannot
    // differ between the two.  (The file bytes themselves remain unpinned against mido: DESIGN.md section 3.)
    const double ticks_per_sec = 960.0;
    std::vector<Msg> msgs;
    msgs.reserve(static_cast<size_t>(n_events) * 4);
    for (int i = 0; i < n_events; ++i) {
        const aegis_note_event& e = events[i];
        const long long st = static_cast<long long>(e.start * secs_per_frame * ticks_per_sec);
        const long long et = static_cast<long long>(e.end * secs_per_frame * ticks_per_sec);
        int velocity = e.velocity;
        if (e.technique == 4) velocity = static_cast<int>(velocity * 0.6);        // hammer_on (:119-120)
        else if (e.technique == 5) velocity = static_cast<int>(velocity * 0.5);   // pull_off (:121-122)
        if (!note_ok(e.note, velocity)) {
            set_error("aegis_smf_write_v1: event %d has note %d / velocity %d outside 0..127", i, e.note, velocity);
            return -1;
        }
        const int tr = e.track ? 1 : 0;
        msgs.push_back({st, 0, tr, e.note, velocity});
        msgs.push_back({et, 1, tr, e.note, 0});
        const long long duration_ticks = et - st;
        if (e.technique == 2) {            // bend (:128-147): accelerating curve towards min(2, 10 |slope|) semitones
            const double bend_semitones = std::min(2.0, std::fabs(e.slope) * 10);
            const int direction = e.slope > 0 ? 1 : -1;
            const int max_bend = static_cast<int>(direction * (bend_semitones / 2.0) * 8191);
            for (int k = 0; k < 15; ++k) {
                const double progress = static_cast<double>(k) / 15;
                const double curve = 1 - std::pow(1 - progress, 2.0);
                msgs.push_back({st + static_cast<long long>(progress * duration_ticks), 2, tr, static_cast<int>(max_bend * curve), 0});
            }
            msgs.push_back({et, 2, tr, 0, 0});
        } else if (e.technique == 1) {     // vibrato (:150-162)
            const double duration_secs = duration_ticks / ticks_per_sec;
            const int n_points = std::max(10, std::min(20, static_cast<int>(duration_secs * opt->vibrato_rate * 4)));
            for (int k = 0; k < n_points; ++k) {
                const double frac = static_cast<double>(k) / n_points;
                const double phase = frac * duration_secs * opt->vibrato_rate * 2 * 3.141592653589793;
                const int bend = static_cast<int>(std::sin(phase) * 8191 * opt->vibrato_depth);
                if (bend < -8192 || bend > 8191) {
                    set_error("aegis_smf_write_v1: vibrato depth %g drives the pitch wheel out of range", opt->vibrato_depth);
                    return -1;
                }
                msgs.push_back({st + static_cast<long long>(frac * duration_ticks), 2, tr, bend, 0});
            }
            msgs.push_back({et, 2, tr, 0, 0});
        }
    }
    std::stable_sort(msgs.begin(), msgs.end(), [](const Msg& x, const Msg& y) { return x.tick < y.tick; });   // list.sort is stable
    TrackBytes tracks[2];   // [0] main, [1] safe, in the order the reference appends them
    long long last[2] = {0, 0};
    for (int t = 0; t < 2; ++t) tracks[t].channel(0, 0xc0, opt->midi_program);
    for (const Msg& m : msgs) {
        const int t = m.track ? 0 : 1;
        const long long delta = m.tick - last[t];
        if (delta < 0) {
            set_error("aegis_smf_write_v1: negative delta time (events end before they start)");
            return -1;
        }
        if (m.kind == 2) {
            const int v = m.a + 8192;
            tracks[t].channel(static_cast<unsigned long long>(delta), 0xe0, v & 0x7f, v >> 7);
        } else {
            tracks[t].channel(static_cast<unsigned long long>(delta), m.kind == 0 ? 0x90 : 0x80, m.a, m.b);
        }
        last[t] = m.tick;
    }
    return finish_file(tracks, 2, out, capacity);
}

extern "C" long long aegis_smf_write_v2(const aegis_fin_event* events, int32_t n_events, const aegis_smf_options* opt,
                                        uint8_t* out, long long capacity) {
    using namespace aegis;
    if (n_events < 0 || (n_events > 0 && events == nullptr) || opt == nullptr || opt->hop <= 0 || !(opt->sr > 0)) {
        set_error("aegis_smf_write_v2: bad arguments");
        return -1;
    }
    TrackBytes tracks[2];
    static const char kMain[] = "Aegis Financial - Main", kSafe[] = "Aegis Financial - Safe";
    tracks[0].meta(0, 0x03, kMain, sizeof(kMain) - 1);
    tracks[1].meta(0, 0x03, kSafe, sizeof(kSafe) - 1);
    const double ms_per_tick = 500.0 / 480;                                         // 120 BPM (:206)
    const double ms_per_frame = (static_cast<double>(opt->hop) / opt->sr) * 1000;   // :216
    long long last[2] = {0, 0};
    for (int i = 0; i < n_events; ++i) {
        const aegis_fin_event& e = events[i];
        const int t = e.track ? 0 : 1;
        const double start_ms = e.start * ms_per_frame;
        const double duration_ms = (e.end - e.start) * ms_per_frame;
        const long long start_ticks = static_cast<long long>(start_ms / ms_per_tick);
        const long long duration_ticks = static_cast<long long>(duration_ms / ms_per_tick);
        const long long delta = start_ticks - last[t];
        if (!note_ok(e.note, e.velocity) || delta < 0 || duration_ticks < 0) {
            set_error("aegis_smf_write_v2: event %d cannot be written (note %d, velocity %d, delta %lld)", i, e.note, e.velocity, delta);
            return -1;
        }
        tracks[t].channel(static_cast<unsigned long long>(delta), 0x90, e.note, e.velocity);
        tracks[t].channel(static_cast<unsigned long long>(duration_ticks), 0x80, e.note, 0);
        last[t] = start_ticks + duration_ticks;
    }
    return finish_file(tracks, 2, out, capacity);
}

// generate_tabs (aegis_engine_core/tabs.py:1-40): per note the (string, fret) closest to a leaking "centre of gravity"
// of the fretting hand; string_out[i] = 0 marks a note no string can play (the reference skips it).
extern "C" int aegis_tabs(const int32_t* notes, int32_t n, int32_t* string_out, int32_t* fret_out) {
    using namespace aegis;
    AEGIS_REQUIRE(n >= 0 && (n == 0 || (notes && string_out && fret_out)), "aegis_tabs: bad arguments");
    static const int kOpen[6] = {64, 59, 55, 50, 45, 40};   // standard tuning, string 1 = high E
    double centre = 5;
    for (int i = 0; i < n; ++i) {
        int best_s = 0, best_f = 0;
        double best = 0;
        for (int s = 0; s < 6; ++s) {
            const int fret = notes[i] - kOpen[s];
            if (fret < 0 || fret > 24) continue;
            const double score = std::fabs(fret - centre) * 1.5 + (s + 1) * 0.2;
            if (best_s == 0 || score < best) { best = score; best_s = s + 1; best_f = fret; }   // min(): first minimum wins
        }
        string_out[i] = best_s;
        fret_out[i] = best_f;
        if (best_s) centre = (centre * 0.7) + (best_f * 0.3);
    }
    return 0;
}

// export_musicxml (aegis_engine_core/tabs.py:42-112): one 4/4 measure of quarter notes with string / fret technical
// marks and bend / slide / vibrato symbols, byte for byte what xml.etree.ElementTree writes for the reference's tree
// (declaration with single quotes, no whitespace between elements, attributes in insertion order, ` />` for empty
// elements).  techniques: 0 none, 1 vibrato, 2 bend, 3 slide (the v1 codes; anything else adds no symbol).
extern "C" long long aegis_musicxml_write(const int32_t* notes, const int32_t* strings, const int32_t* frets,
                                          const uint8_t* techniques, int32_t n, uint8_t* out, long long capacity) {
    using namespace aegis;
    if (n < 0 || (n > 0 && (!notes || !strings || !frets))) {
        set_error("aegis_musicxml_write: bad arguments");
        return -1;
    }
    static const char* const kStep[12] = {"C", "C", "D", "D", "E", "F", "F", "G", "G", "A", "A", "B"};
    static const bool kSharp[12] = {false, true, false, true, false, false, true, false, true, false, true, false};
    std::string x;
    x.reserve(512 + static_cast<size_t>(n) * 256);
    x += "<?xml version='1.0' encoding='UTF-8'?>\n"
         "<score-partwise version=\"3.1\"><part-list><score-part id=\"P1\"><part-name>Aegis Guitar</part-name></score-part>"
         "</part-list><part id=\"P1\"><measure number=\"1\"><attributes><divisions>1</divisions><key><fifths>0</fifths></key>"
         "<time><beats>4</beats><beat-type>4</beat-type></time><clef><sign>G</sign><line>2</line></clef>"
         "<staff-details><staff-lines>6</staff-lines></staff-details></attributes>";
    for (int i = 0; i < n; ++i) {
        const int pitch = notes[i];
        const int pc = ((pitch % 12) + 12) % 12;
        // Python's floor division: (pitch // 12) - 1
        const int octave = (pitch >= 0 ? pitch / 12 : -((-pitch + 11) / 12)) - 1;
        x += "<note><pitch><step>";
        x += kStep[pc];
        x += "</step>";
        if (kSharp[pc]) x += "<alter>1</alter>";
        x += "<octave>" + std::to_string(octave) + "</octave></pitch><duration>1</duration><type>quarter</type><notations><technical><string>" +
             std::to_string(strings[i]) + "</string><fret>" + std::to_string(frets[i]) + "</fret>";
        const int tech = techniques ? techniques[i] : 0;
        if (tech == 2) x += "<bend><bend-alter>2</bend-alter></bend></technical>";
        else if (tech == 3) x += "</technical><slur type=\"start\" number=\"1\" />";
        else if (tech == 1) x += "<hammer-on type=\"start\" /></technical><ornaments><wavy-line type=\"start\" number=\"1\" /></ornaments>";
        else x += "</technical>";
        x += "</notations></note>";
    }
    x += "</measure></part></score-partwise>";
    const long long need = static_cast<long long>(x.size());
    if (out != nullptr && capacity >= need) std::memcpy(out, x.data(), x.size());
    return need;
}

