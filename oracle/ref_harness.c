/*
 * ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Builds the UNMODIFIED reference translation unit (/root/reference/ctts.c,
 * found through -I at compile time; never copied into this repo) into
 * oracle/_ref/libctts_ref.so and exposes (a) the reference's own public
 * `ctts_synthesize` and (b) thin wrappers around its `static` signal
 * functions so that per-stage golden vectors can be generated and the CPU
 * restatement in ctts_oracle.c can be pinned against the real thing.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#define main ctts_reference_main
#include "ctts.c"
#undef main

#include <errno.h>

static CTTS* g_engine = NULL;

/* Open the engine the way `ctts synth` does (ctts.c:3983-3990), but with
 * explicit paths instead of cwd-relative ones. */
int ref_open(const char* db_path, const char* config_path, const char* norm_csv_path) {
    if (g_engine) {
        ctts_free(g_engine);
        g_engine = NULL;
    }
    g_engine = ctts_init(db_path);
    if (!g_engine) return -1;
    if (config_path) ctts_load_config(&g_engine->config, config_path);
    /* ctts_synthesize loads "normalization.csv" from cwd only if no rules
     * were loaded before (ctts.c:344); load ours first. */
    if (norm_csv_path) ctts_load_normalization(norm_csv_path);
    else norm_rules_loaded = 1;
    /* same for duration rules (loaded, never applied: ctts.c:3636) */
    duration_rules_loaded = 1;
    return 0;
}

void ref_close(void) {
    if (g_engine) ctts_free(g_engine);
    g_engine = NULL;
}

int ref_get_config(CTTSConfig* out) {
    if (!g_engine) return -1;
    *out = g_engine->config;
    return 0;
}

int ref_set_config(const CTTSConfig* in) {
    if (!g_engine) return -1;
    g_engine->config = *in;
    return 0;
}

size_t ref_rule_count(void) { return norm_rule_count; }

/* The reference's public API, verbatim. */
int ref_synth(const char* text, float speed, int16_t** out, size_t* n) {
    if (!g_engine) return -1;
    return ctts_synthesize(g_engine, text, out, n, speed);
}

void ref_free(void* p) { free(p); }

/* ctts_synthesize with config.print_units=1 and stderr captured: the unit
 * texts it prints at ctts.c:3795 are the cheapest bit-exact check of a unit
 * plan.  Returns a malloc'ed string "  [u1]   [u2] ...\n". */
char* ref_unit_trace(const char* text) {
    if (!g_engine) return NULL;
    char* cap = NULL;
    size_t cap_len = 0;
    FILE* saved = stderr;
    FILE* mem = open_memstream(&cap, &cap_len);
    if (!mem) return NULL;
    int saved_flag = g_engine->config.print_units;
    g_engine->config.print_units = 1;
    stderr = mem;
    int16_t* s = NULL;
    size_t n = 0;
    int err = ctts_synthesize(g_engine, text, &s, &n, 1.0f);
    stderr = saved;
    fclose(mem);
    g_engine->config.print_units = saved_flag;
    if (err == CTTS_OK) free(s);
    return cap;
}

/* Front-end text pipeline of ctts.c:3643-3655. */
char* ref_normalized_text(const char* text) {
    char* a = expand_numbers(text);
    if (!a) return NULL;
    char* b = ctts_apply_normalization(a);
    free(a);
    if (!b) return NULL;
    char* c = ctts_normalize(b);
    free(b);
    return c;
}

/* ---- per-stage wrappers around the reference's statics ---- */

void ref_fade_luts(float* out3x1024) {
    init_fade_luts();
    memcpy(out3x1024, fade_out_lut, sizeof(fade_out_lut));
    memcpy(out3x1024 + FADE_LUT_SIZE, fade_in_lut, sizeof(fade_in_lut));
    memcpy(out3x1024 + 2 * FADE_LUT_SIZE, sine_fade_lut, sizeof(sine_fade_lut));
}

void ref_hann256(float* out) {
    init_hanning_window();
    memcpy(out, hanning_window, sizeof(hanning_window));
}

void ref_hann512(float* out) {
    for (size_t i = 0; i < 512; i++) out[i] = hanning(i, 512);
}

void ref_normalize_rms(int16_t* s, size_t n, float target) { normalize_rms(s, n, target); }
float ref_calculate_rms(const int16_t* s, size_t n) { return calculate_rms(s, n); }
void ref_remove_dc_offset(int16_t* s, size_t n) { remove_dc_offset(s, n); }
float ref_estimate_pitch(const int16_t* s, size_t n) { return estimate_pitch(s, n); }
void ref_apply_pitch_shift(int16_t* s, size_t n, float f) { apply_pitch_shift(s, n, f); }

void ref_smooth_pitch_boundary(int16_t* prev, size_t prev_n, int16_t* next, size_t next_n,
                               size_t boundary) {
    smooth_pitch_boundary(prev, prev_n, next, next_n, boundary);
}

void ref_match_boundary_energy(int16_t* prev, size_t prev_n, int16_t* next, size_t next_n,
                               size_t xf) {
    match_boundary_energy(prev, prev_n, next, next_n, xf);
}

size_t ref_remove_silence_regions(int16_t* s, size_t n, float thr, size_t min_sil) {
    return remove_silence_regions(s, n, thr, min_sil);
}

void ref_apply_smooth_pitch_contour(int16_t* s, size_t n, float f0, float f1) {
    apply_smooth_pitch_contour(s, n, f0, f1);
}

void ref_apply_fade_in(int16_t* s, size_t n, size_t f) { apply_fade_in(s, n, f); }
void ref_apply_fade_out(int16_t* s, size_t n, size_t f) { apply_fade_out(s, n, f); }

/* apply_phrase_intonation for (phrase_type, word_index, total_words). */
void ref_apply_phrase_intonation(int16_t* s, size_t n, int phrase_type, int word_index,
                                 int total_words, float max_pitch_change) {
    PhraseIntonation in = get_phrase_intonation_limited((PhraseType)phrase_type, max_pitch_change);
    apply_phrase_intonation(s, n, &in, word_index, total_words, max_pitch_change);
}

/* buffer_append_crossfade on a caller-provided buffer (capacity must be
 * >= *count + n).  Uses the opened engine's config. */
int ref_append_crossfade(int16_t* buf, size_t* count, size_t cap, const int16_t* unit, size_t n,
                         float xf_ms, int after_boundary) {
    if (!g_engine) return -1;
    SampleBuffer b;
    b.data = malloc((cap + n + 8192) * sizeof(int16_t));
    if (!b.data) return -2;
    memcpy(b.data, buf, *count * sizeof(int16_t));
    b.count = *count;
    b.capacity = cap + n + 8192;
    int err = buffer_append_crossfade(&b, unit, n, xf_ms, &g_engine->config, after_boundary);
    if (err == CTTS_OK && b.count <= cap) {
        memcpy(buf, b.data, b.count * sizeof(int16_t));
        *count = b.count;
    } else if (err == CTTS_OK) {
        err = -3;
    }
    free(b.data);
    return err;
}

int ref_time_stretch(const int16_t* in, size_t n, int16_t** out, size_t* n_out, float speed) {
    return time_stretch(in, n, out, n_out, speed);
}

int ref_wsola_offset(const int16_t* in, size_t n, const int16_t* prev_frame, size_t nominal) {
    return find_best_match_wsola(in, n, prev_frame, 384, nominal, 512, 128);
}

float ref_cross_correlation(const int16_t* a, const int16_t* b, size_t n) {
    return cross_correlation(a, b, n);
}

int ref_find_unit(const char* text, size_t len) {
    if (!g_engine) return -2;
    return find_unit(g_engine, text, len);
}

int ref_build_database(const char* dataset_dir, const char* out_db) {
    char ld[1024], li[1024], sd[1024], si[1024];
    snprintf(ld, sizeof ld, "%s/letters/wavs", dataset_dir);
    snprintf(li, sizeof li, "%s/letters/letters.txt", dataset_dir);
    snprintf(sd, sizeof sd, "%s/syllables/wavs", dataset_dir);
    snprintf(si, sizeof si, "%s/syllables/sillabes.txt", dataset_dir);
    return ctts_build_database(ld, li, sd, si, out_db);
}

int ref_write_wav(const char* path, const int16_t* s, size_t n) {
    return ctts_write_wav(path, s, n, CTTS_SAMPLE_RATE);
}
