/*
 * ctts_oracle.c -- TEST INFRASTRUCTURE, not product code (see ctts_oracle.h).
 *
 * Plain-C restatement of the sample-touching half of the reference's
 * `ctts_synthesize` (ctts.c:3689-3921), executed from a plan.  Integer and
 * float arithmetic follows the reference expression by expression (same
 * operand types, same order, no FMA: build with -ffp-contract=off) because the
 * bar is bit-exact PCM.  Each function cites the reference lines it follows.
 */
#include "ctts_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_D 3.14159265358979323846 /* double, as ctts.c:45 */
#define LUT_N 1024                  /* ctts.c:52 */
#define PITCH_FRAME 256             /* ctts.c:2194 */
#define WS_FRAME 512                /* ctts.c:3506 */
#define WS_HOP 128
#define WS_OVERLAP 384
#define WS_SHIFT 128

typedef struct {
    uint32_t magic, version, unit_count, sample_rate, bits_per_sample, index_offset,
        strings_offset, audio_offset, total_samples, max_unit_chars, hash_table_size,
        hash_table_offset;
    uint8_t reserved[16];
} db_header;

typedef struct {
    uint32_t hash, string_offset;
    uint16_t string_len, char_count;
    uint32_t audio_offset, sample_count, flags, next_hash, reserved;
} db_entry;

struct ctts_oracle {
    const uint8_t* db;
    db_header hdr;
    const db_entry* index;
    const uint8_t* pcm_bytes; /* may be 2-byte misaligned (odd audio_offset) */
    float fade_out_lut[LUT_N], fade_in_lut[LUT_N], sine_lut[LUT_N];
    float hann256[PITCH_FRAME], hann512[WS_FRAME];
};

/* float -> int16 the way x86-64 gcc does it: cvttss2si then truncate */
static inline int16_t f2s(float v) { return (int16_t)(int32_t)v; }

static inline float clamp16f(float v) {
    if (v > 32767.0f) v = 32767.0f;
    if (v < -32768.0f) v = -32768.0f;
    return v;
}

/* ------------------------------------------------------------------ tables */

int ctts_oracle_open(ctts_oracle** out, const void* voice_db, size_t db_size) {
    if (!out || !voice_db || db_size < sizeof(db_header)) return -1;
    ctts_oracle* o = calloc(1, sizeof *o);
    if (!o) return -6;
    o->db = voice_db;
    memcpy(&o->hdr, voice_db, sizeof o->hdr);
    if (o->hdr.magic != 0x53545443u) { free(o); return -5; }
    if (o->hdr.version != 1u) { free(o); return -8; }
    if ((uint64_t)o->hdr.audio_offset + 2ull * o->hdr.total_samples > db_size) { free(o); return -5; }
    o->index = (const db_entry*)(o->db + o->hdr.index_offset);
    o->pcm_bytes = o->db + o->hdr.audio_offset;
    for (int i = 0; i < LUT_N; i++) { /* init_fade_luts, ctts.c:60-73 */
        float t = (float)i / (float)(LUT_N - 1);
        o->fade_out_lut[i] = 0.5f * (1.0f + cosf(PI_D * t));
        o->fade_in_lut[i] = 0.5f * (1.0f - cosf(PI_D * t));
        o->sine_lut[i] = sinf(t * PI_D * 0.5f);
    }
    for (int i = 0; i < PITCH_FRAME; i++) /* init_hanning_window, ctts.c:2198 */
        o->hann256[i] = 0.5f * (1.0f - cosf(2.0f * PI_D * i / PITCH_FRAME));
    for (size_t i = 0; i < WS_FRAME; i++) /* hanning(), ctts.c:1624 */
        o->hann512[i] = 0.5f * (1.0f - cosf(2.0f * (float)PI_D * (float)i / (float)WS_FRAME));
    *out = o;
    return 0;
}

void ctts_oracle_close(ctts_oracle* o) { free(o); }
void ctts_oracle_free(void* p) { free(p); }

void ctts_oracle_tables(const ctts_oracle* o, float* luts, float* h256, float* h512) {
    if (luts) {
        memcpy(luts, o->fade_out_lut, sizeof o->fade_out_lut);
        memcpy(luts + LUT_N, o->fade_in_lut, sizeof o->fade_in_lut);
        memcpy(luts + 2 * LUT_N, o->sine_lut, sizeof o->sine_lut);
    }
    if (h256) memcpy(h256, o->hann256, sizeof o->hann256);
    if (h512) memcpy(h512, o->hann512, sizeof o->hann512);
}

/* fast_fade_out / fast_fade_in / fast_sine_fade, ctts.c:76-101 */
static inline float lut_lerp(const float* lut, float t) {
    float x = t * (LUT_N - 1);
    int k = (int)x;
    if (k >= LUT_N - 1) return lut[LUT_N - 1];
    if (k < 0) return lut[0];
    float fr = x - k;
    return lut[k] * (1.0f - fr) + lut[k + 1] * fr;
}

/* ------------------------------------------------------------ unit stages */

float ctts_oracle_rms(const int16_t* s, size_t n) { /* calculate_rms, ctts.c:1697 */
    if (n == 0) return 0.0f;
    double acc = 0.0;
    for (size_t i = 0; i < n; i++) {
        double v = (double)s[i];
        acc += v * v;
    }
    return (float)sqrt(acc / n);
}

void ctts_oracle_normalize_rms(int16_t* s, size_t n, float target) { /* ctts.c:1709 */
    if (n == 0 || target <= 0) return;
    float rms = ctts_oracle_rms(s, n);
    if (rms < 1.0f) return;
    float g = target / rms;
    if (g > 3.0f) g = 3.0f;
    if (g < 0.1f) g = 0.1f;
    for (size_t i = 0; i < n; i++) s[i] = f2s(clamp16f(s[i] * g));
}

void ctts_oracle_remove_dc(int16_t* s, size_t n) { /* remove_dc_offset, ctts.c:1568 */
    if (n == 0) return;
    int64_t sum = 0;
    for (size_t i = 0; i < n; i++) sum += s[i];
    int16_t dc = (int16_t)(sum / (int64_t)n);
    for (size_t i = 0; i < n; i++) {
        int32_t v = s[i] - dc;
        if (v > 32767) v = 32767;
        if (v < -32768) v = -32768;
        s[i] = (int16_t)v;
    }
}

float ctts_oracle_estimate_pitch(const int16_t* s, size_t n) { /* ctts.c:1899 */
    if (n < 200) return 0.0f;
    size_t lo = CTTS_PLAN_SAMPLE_RATE / 400, hi = CTTS_PLAN_SAMPLE_RATE / 80;
    if (hi > n / 2) hi = n / 2;
    size_t len = CTTS_PLAN_SAMPLE_RATE / 100;
    if (len > n - hi) len = n - hi;
    float best = 0.0f;
    size_t best_lag = 0;
    for (size_t lag = lo; lag <= hi; lag++) {
        float c = 0.0f, e1 = 0.0f, e2 = 0.0f;
        for (size_t i = 0; i < len; i++) {
            float a = s[i], b = s[i + lag];
            c += a * b;
            e1 += a * a;
            e2 += b * b;
        }
        float nrm = sqrtf(e1 * e2);
        if (nrm > 0) c /= nrm;
        if (c > best) {
            best = c;
            best_lag = lag;
        }
    }
    if (best > 0.3f && best_lag > 0) return (float)CTTS_PLAN_SAMPLE_RATE / best_lag;
    return 0.0f;
}

/* apply_pitch_shift, ctts.c:1946: linear-interpolation resample, same length */
static void resample_head(int16_t* s, size_t n, float factor) {
    if (factor < 0.9f || factor > 1.1f || n < 100) return;
    size_t m = (size_t)(n / factor);
    size_t keep = m < n ? m : n;
    int16_t* tmp = calloc(keep ? keep : 1, sizeof *tmp);
    if (!tmp) return;
    for (size_t i = 0; i < keep; i++) {
        float x = i * factor;
        size_t k = (size_t)x;
        float fr = x - k;
        if (k + 1 < n) tmp[i] = f2s(s[k] * (1.0f - fr) + s[k + 1] * fr);
        else if (k < n) tmp[i] = s[k];
    }
    memcpy(s, tmp, keep * sizeof *tmp);
    if (keep < n) memset(s + keep, 0, (n - keep) * sizeof *s);
    free(tmp);
}

/* smooth_pitch_boundary, ctts.c:1979.  Returns 1 if the >15% branch ran. */
int ctts_oracle_smooth_pitch(const int16_t* buf, size_t count, int16_t* unit, size_t n, size_t xf) {
    if (xf == 0 || count < 200 || n < 200) return 0;
    size_t reg = xf * 2;
    if (reg > count / 2) reg = count / 2;
    if (reg > n / 2) reg = n / 2;
    float pp = ctts_oracle_estimate_pitch(buf + count - reg, reg);
    float np = ctts_oracle_estimate_pitch(unit, reg);
    if (!(pp > 0 && np > 0)) return 0;
    float ratio = np / pp;
    if (!(ratio > 1.15f || ratio < 0.85f)) return 0;
    float target = (ratio > 1.0f) ? 1.0f + (ratio - 1.0f) * 0.5f : 1.0f - (1.0f - ratio) * 0.5f;
    float shift = target / ratio;
    size_t len = xf;
    if (len > n / 4) len = n / 4;
    int16_t* head = malloc((len ? len : 1) * sizeof *head);
    if (!head) return 1;
    memcpy(head, unit, len * sizeof *head);
    resample_head(head, len, shift);
    for (size_t i = 0; i < len; i++) {
        float t = (float)i / len;
        unit[i] = f2s(head[i] * (1.0f - t) + unit[i] * t);
    }
    free(head);
    return 1;
}

/* match_boundary_energy, ctts.c:1730 */
void ctts_oracle_match_energy(const int16_t* buf, size_t count, int16_t* unit, size_t n, size_t xf) {
    if (xf == 0 || count == 0 || n == 0) return;
    size_t len = xf;
    if (len > count) len = count;
    if (len > n) len = n;
    float pr = ctts_oracle_rms(buf + count - len, len);
    float nr = ctts_oracle_rms(unit, len);
    if (pr < 1.0f || nr < 1.0f) return;
    float ratio = pr / nr;
    if (ratio > 2.0f) ratio = 2.0f;
    if (ratio < 0.5f) ratio = 0.5f;
    for (size_t i = 0; i < len; i++) {
        float t = (float)i / (float)len;
        float g = ratio * (1.0f - t) + 1.0f * t;
        unit[i] = f2s(clamp16f(unit[i] * g));
    }
}

void ctts_oracle_fade_in(const ctts_oracle* o, int16_t* s, size_t n, size_t f) { /* ctts.c:3015 */
    if (f == 0 || n == 0) return;
    if (f > n) f = n;
    float inv = 1.0f / (float)f;
    for (size_t i = 0; i < f; i++) s[i] = f2s(s[i] * lut_lerp(o->sine_lut, (float)i * inv));
}

void ctts_oracle_fade_out(const ctts_oracle* o, int16_t* s, size_t n, size_t f) { /* ctts.c:3028 */
    if (f == 0 || n == 0) return;
    if (f > n) f = n;
    int16_t* tail = s + (n - f);
    float inv = 1.0f / (float)f;
    for (size_t i = 0; i < f; i++)
        tail[i] = f2s(tail[i] * lut_lerp(o->sine_lut, (float)(f - i) * inv));
}

/* buffer_append_crossfade, ctts.c:3279 (unit is already a private copy) */
size_t ctts_oracle_append(const ctts_oracle* o, int16_t* buf, size_t count, int16_t* unit, size_t n,
                          size_t xf, size_t fade_in, int remove_dc, int after_boundary) {
    if (n == 0) return count;
    if (remove_dc) ctts_oracle_remove_dc(unit, n);
    if (count == 0 || after_boundary) {
        ctts_oracle_fade_in(o, unit, n, fade_in);
        memcpy(buf + count, unit, n * sizeof *unit);
        return count + n;
    }
    if (xf == 0) {
        memcpy(buf + count, unit, n * sizeof *unit);
        return count + n;
    }
    size_t a = xf;
    if (a > count) a = count;
    if (a > n) a = n;
    if (a > 0) {
        int16_t* tail = buf + (count - a);
        float inv = 1.0f / (float)a;
        for (size_t i = 0; i < a; i++) {
            float t = (float)i * inv;
            float pg = lut_lerp(o->fade_out_lut, t);
            float ng = lut_lerp(o->fade_in_lut, t);
            int32_t p = tail[i], q = unit[i];
            int32_t mix = (int32_t)(p * pg + q * ng);
            if (mix > 32767) mix = 32767;
            else if (mix < -32768) mix = -32768;
            tail[i] = (int16_t)mix;
        }
    }
    if (n > a) {
        memcpy(buf + count, unit + a, (n - a) * sizeof *unit);
        count += n - a;
    }
    return count;
}

/* ------------------------------------------------------------ word stages */

/* abs() as the reference computes it on int16 (ctts.c:1641): -32768 stays -32768 */
static inline int16_t abs16(int16_t v) { return (int16_t)(v > 0 ? v : -v); }

/* remove_silence_regions, ctts.c:1634 */
size_t ctts_oracle_trim(int16_t* s, size_t n, float thr, size_t min_sil) {
    if (n == 0) return 0;
    int16_t peak = 0;
    for (size_t i = 0; i < n; i++)
        if (abs16(s[i]) > peak) peak = abs16(s[i]);
    if (peak == 0) return n;
    int16_t limit = f2s(peak * thr);
    size_t keep = min_sil / 4;
    if (keep < 10) keep = 10;
    size_t w = 0, r = 0;
    while (r < n) {
        if (abs16(s[r]) > limit) {
            s[w++] = s[r++];
            continue;
        }
        size_t start = r;
        while (r < n && abs16(s[r]) <= limit) r++;
        size_t run = r - start;
        size_t take = run >= min_sil ? keep : run;
        for (size_t i = 0; i < take && start + i < n; i++) s[w++] = s[start + i];
    }
    return w;
}

/* apply_smooth_pitch_contour, ctts.c:2206 */
uint32_t ctts_oracle_contour(const ctts_oracle* o, int16_t* s, size_t n, float f0, float f1,
                             size_t* taint_lo, size_t* taint_hi) {
    if (taint_lo) *taint_lo = 0;
    if (taint_hi) *taint_hi = 0;
    if (n < 100 || fabsf(f0 - f1) < 0.01f) return 0;
    int16_t* orig = malloc(n * sizeof *orig);
    float* norm = calloc(n, sizeof *norm);
    uint8_t* taint = calloc(n, 1);
    if (!orig || !norm || !taint) {
        free(orig);
        free(norm);
        free(taint);
        return 0;
    }
    memcpy(orig, s, n * sizeof *orig);
    memset(s, 0, n * sizeof *s);
    uint32_t oob = 0;
    float inv = 1.0f / (float)(n - PITCH_FRAME);
    for (size_t pos = 0; pos + PITCH_FRAME <= n; pos += PITCH_FRAME / 2) {
        float t = (float)pos * inv;
        float st = t * t * (3.0f - 2.0f * t);
        float pf = f0 + (f1 - f0) * st;
        for (size_t i = 0; i < PITCH_FRAME; i++) {
            float w = o->hann256[i];
            float x = i * pf;
            size_t k = (size_t)x;
            float fr = x - k;
            float v;
            if (k + 1 < PITCH_FRAME) {
                v = orig[pos + k] * (1.0f - fr) + orig[pos + k + 1] * fr;
            } else if (pos + k < n) {
                v = orig[pos + k];
            } else { /* the reference reads past its heap copy here */
                v = 0.0f;
                oob++;
                taint[pos + i] = 1;
            }
            s[pos + i] = (int16_t)(s[pos + i] + f2s(v * w));
            norm[pos + i] += w;
        }
    }
    size_t lo = n, hi = 0;
    for (size_t i = 0; i < n; i++) {
        if (norm[i] > 0.01f) {
            s[i] = f2s(clamp16f(s[i] / norm[i]));
            if (taint[i]) {
                if (i < lo) lo = i;
                hi = i + 1;
            }
        } else {
            s[i] = orig[i];
        }
    }
    if (hi > 0) {
        if (taint_lo) *taint_lo = lo;
        if (taint_hi) *taint_hi = hi;
    }
    free(orig);
    free(norm);
    free(taint);
    return oob;
}

static void add_span(ctts_oracle_stats* st, size_t base, size_t lo, size_t hi) {
    if (!st || hi <= lo) return;
    if (st->ub_spans < CTTS_ORACLE_MAX_UB_SPANS) {
        st->ub_span[st->ub_spans][0] = base + lo;
        st->ub_span[st->ub_spans][1] = base + hi;
    }
    st->ub_spans++;
}

/* device half of apply_phrase_intonation, ctts.c:2740, :2774-2790, :2839-2865 */
static void word_prosody(const ctts_oracle* o, int16_t* s, size_t n, size_t abs_base,
                         const ctts_plan_op* op, ctts_oracle_stats* st) {
    if (!(op->flags & CTTS_WE_INTON) || n < 100) return;
    size_t lo, hi;
    int done = 0;
    if (op->flags & CTTS_WE_CIRCUMFLEX) {
        size_t rise = (size_t)(n * 0.6f);
        if (rise > 100 && n - rise > 100) {
            ctts_oracle_contour(o, s, rise, op->f0, op->f2, &lo, &hi);
            add_span(st, abs_base, lo, hi);
            ctts_oracle_contour(o, s + rise, n - rise, op->f2, op->f1, &lo, &hi);
            add_span(st, abs_base + rise, lo, hi);
            if (st) st->contour_calls += 2;
            done = 1;
        }
    }
    if (!done) {
        ctts_oracle_contour(o, s, n, op->f0, op->f1, &lo, &hi);
        add_span(st, abs_base, lo, hi);
        if (st) st->contour_calls += 1;
    }
    if (op->flags & CTTS_WE_ENERGY) {
        for (size_t i = 0; i < n; i++) {
            float t = (float)i / (float)(n - 1);
            float e = op->e0 + (op->e1 - op->e0) * t;
            s[i] = f2s(clamp16f(s[i] * e));
        }
    }
}

/* ------------------------------------------------------------------ WSOLA */

/* cross_correlation, ctts.c:3390 (len is always 384 here, a multiple of 4) */
static float xcorr(const int16_t* a, const int16_t* b, size_t len) {
    if (len == 0) return 0.0f;
    float sp = 0.0f, sa = 0.0f, sb = 0.0f;
    size_t i = 0, len4 = len & ~(size_t)3;
    for (; i < len4; i += 4) {
        float a0 = a[i], a1 = a[i + 1], a2 = a[i + 2], a3 = a[i + 3];
        float b0 = b[i], b1 = b[i + 1], b2 = b[i + 2], b3 = b[i + 3];
        sp += a0 * b0 + a1 * b1 + a2 * b2 + a3 * b3;
        sa += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        sb += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
    }
    for (; i < len; i++) {
        float x = a[i], y = b[i];
        sp += x * y;
        sa += x * x;
        sb += y * y;
    }
    float den = sqrtf(sa * sb);
    if (den < 1.0f) return 0.0f;
    return sp / den;
}

/* find_best_match_wsola, ctts.c:3436: coarse step 4, then +-3 around it */
int ctts_oracle_wsola_offset(const int16_t* in, size_t n, const int16_t* prev_frame, size_t nominal) {
    const int16_t* target = prev_frame + WS_FRAME - WS_OVERLAP;
    float best = -2.0f;
    int best_off = 0;
    for (int off = -WS_SHIFT; off <= WS_SHIFT; off += 4) {
        int pos = (int)nominal + off;
        if (pos < 0 || (size_t)pos + WS_FRAME > n) continue;
        float c = xcorr(in + pos, target, WS_OVERLAP);
        if (c > best) {
            best = c;
            best_off = off;
        }
    }
    int lo = best_off - 3, hi = best_off + 3;
    if (lo < -WS_SHIFT) lo = -WS_SHIFT;
    if (hi > WS_SHIFT) hi = WS_SHIFT;
    for (int off = lo; off <= hi; off++) {
        if (off == best_off) continue;
        int pos = (int)nominal + off;
        if (pos < 0 || (size_t)pos + WS_FRAME > n) continue;
        float c = xcorr(in + pos, target, WS_OVERLAP);
        if (c > best) {
            best = c;
            best_off = off;
        }
    }
    return best_off;
}

/* time_stretch, ctts.c:3490 */
int ctts_oracle_time_stretch(const ctts_oracle* o, const int16_t* in, size_t n, float speed,
                             int16_t** out, size_t* n_out, uint32_t* frames_out) {
    if (speed < 0.5f) speed = 0.5f;
    if (speed > 2.0f) speed = 2.0f;
    if (frames_out) *frames_out = 0;
    if (fabsf(speed - 1.0f) < 0.01f) {
        *out = malloc((n ? n : 1) * sizeof **out);
        if (!*out) return -6;
        memcpy(*out, in, n * sizeof **out);
        *n_out = n;
        return 0;
    }
    size_t hop = (size_t)((size_t)WS_HOP / speed);
    if (hop < 1) hop = 1;
    size_t frames = n > WS_FRAME ? (n - WS_FRAME) / WS_HOP + 1 : 1;
    size_t cap = frames * hop + WS_FRAME + 1024;
    int16_t* y = calloc(cap, sizeof *y);
    float* norm = calloc(cap, sizeof *norm);
    if (!y || !norm) {
        free(y);
        free(norm);
        return -6;
    }
    int16_t prev[WS_FRAME];
    int have_prev = 0;
    size_t nominal = 0, syn = 0, used = 0;
    uint32_t nf = 0;
    while (nominal + WS_FRAME <= n && syn + WS_FRAME <= cap) {
        int off = have_prev ? ctts_oracle_wsola_offset(in, n, prev, nominal) : 0;
        size_t pos = nominal + off;
        if (pos + WS_FRAME > n) pos = n - WS_FRAME;
        for (size_t i = 0; i < WS_FRAME; i++) {
            float v = in[pos + i] * o->hann512[i];
            y[syn + i] = (int16_t)(y[syn + i] + f2s(v));
            norm[syn + i] += o->hann512[i];
            prev[i] = in[pos + i];
        }
        have_prev = 1;
        if (syn + WS_FRAME > used) used = syn + WS_FRAME;
        nominal += WS_HOP;
        syn += hop;
        nf++;
    }
    for (size_t i = 0; i < used; i++)
        if (norm[i] > 0.01f) y[i] = f2s(clamp16f(y[i] / norm[i]));
    free(norm);
    while (used > 0 && y[used - 1] == 0) used--;
    *out = y;
    *n_out = used;
    if (frames_out) *frames_out = nf;
    return 0;
}

/* --------------------------------------------------------------- executor */

typedef struct {
    int16_t* v;
    size_t n, cap;
} pcmbuf;

static int reserve(pcmbuf* b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 1 << 16;
    while (nc < b->n + extra) nc *= 2;
    int16_t* nv = realloc(b->v, nc * sizeof *nv);
    if (!nv) return -6;
    b->v = nv;
    b->cap = nc;
    return 0;
}

int ctts_oracle_synth(const ctts_oracle* o, const ctts_assembly_params* prm,
                      const ctts_plan_op* ops, uint32_t n_ops, float speed, int16_t** out,
                      size_t* n_out, int16_t** pre, size_t* n_pre, ctts_oracle_stats* st) {
    if (!o || !prm || (!ops && n_ops) || !out || !n_out) return -1;
    if (st) memset(st, 0, sizeof *st);
    pcmbuf b = {NULL, 0, 0};
    if (reserve(&b, 1)) return -6;
    size_t word_start = 0;
    int err = 0;

    for (uint32_t k = 0; k < n_ops && !err; k++) {
        const ctts_plan_op* op = &ops[k];
        switch (op->kind) {
            case CTTS_OP_UNIT: { /* ctts.c:3785-3846 */
                if (op->a >= o->hdr.unit_count) { err = -1; break; }
                const db_entry* e = &o->index[op->a];
                size_t n = e->sample_count;
                int16_t* unit = malloc((n ? n : 1) * sizeof *unit);
                if (!unit) { err = -6; break; }
                memcpy(unit, o->pcm_bytes + 2ull * e->audio_offset, n * sizeof *unit);
                ctts_oracle_normalize_rms(unit, n, prm->target_rms);
                int boundary = (op->flags & CTTS_UNIT_AFTER_BOUNDARY) != 0;
                if (!boundary && b.n > 0) {
                    int shifted = ctts_oracle_smooth_pitch(b.v, b.n, unit, n, op->b);
                    ctts_oracle_match_energy(b.v, b.n, unit, n, op->b);
                    if (st) {
                        st->joins++;
                        st->pitch_shifts += (uint32_t)shifted;
                    }
                }
                if ((err = reserve(&b, n)) == 0)
                    b.n = ctts_oracle_append(o, b.v, b.n, unit, n, op->b, prm->fade_in_samples,
                                             (int)prm->remove_dc_offset, boundary);
                free(unit);
                if (st) st->units++;
                break;
            }
            case CTTS_OP_SILENCE: /* buffer_append_silence, ctts.c:3361 */
                if ((err = reserve(&b, op->a)) == 0) {
                    memset(b.v + b.n, 0, (size_t)op->a * sizeof *b.v);
                    b.n += op->a;
                }
                break;
            case CTTS_OP_FADE_OUT: /* ctts.c:3716-3719, :3752-3755, :3371 */
                if (b.n > 0) ctts_oracle_fade_out(o, b.v, b.n, op->a);
                break;
            case CTTS_OP_WORD_END: /* ctts.c:3693-3713 */
                if ((op->flags & CTTS_WE_TRIM) && b.n > word_start) {
                    size_t len = b.n - word_start;
                    if (len > prm->min_silence_samples) {
                        size_t kept = ctts_oracle_trim(b.v + word_start, len, prm->silence_threshold,
                                                       prm->min_silence_samples);
                        if (st) st->trimmed += len - kept;
                        b.n = word_start + kept;
                    }
                }
                if (b.n > word_start)
                    word_prosody(o, b.v + word_start, b.n - word_start, word_start, op, st);
                break;
            case CTTS_OP_MARK:
                word_start = b.n;
                break;
            default:
                err = -1;
        }
    }
    if (err) {
        free(b.v);
        return err;
    }
    if (st) st->pre_count = b.n;
    if (pre && n_pre) {
        *pre = malloc((b.n ? b.n : 1) * sizeof **pre);
        if (*pre) memcpy(*pre, b.v, b.n * sizeof **pre);
        *n_pre = b.n;
    }
    if (speed != 1.0f) { /* ctts.c:3907 */
        uint32_t frames = 0;
        err = ctts_oracle_time_stretch(o, b.v, b.n, speed, out, n_out, &frames);
        free(b.v);
        if (st) st->wsola_frames = frames;
    } else {
        *out = b.v;
        *n_out = b.n;
    }
    if (st && !err) st->out_count = *n_out;
    return err;
}
