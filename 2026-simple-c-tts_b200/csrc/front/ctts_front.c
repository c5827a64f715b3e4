/*
 * ctts_front.c -- host front end: text -> plan ops (see include/ctts_front.h).
 *
 * Behavioural restatement of the text side of the reference's
 * `ctts_synthesize` (ctts.c:3623-3924).  No sample is touched here.  Each
 * function cites the reference lines whose decisions it reproduces; unit ids,
 * crossfade sample counts, pause lengths and op order must be bit-identical to
 * what the reference would do for the same text, voice.db and config.
 *
 * Unlike the reference (global rule tables, ctts.c:34-36) all state lives in
 * the handle, so several handles can plan concurrently on different threads.
 */
#define _GNU_SOURCE
#include "ctts_front.h"

#include <locale.h>
#include <math.h>
#include <pthread.h>
#include <regex.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DB_MAGIC 0x53545443u /* ctts.h:22 */
#define DB_VERSION 1u
#define NO_UNIT 0xFFFFFFFFu
#define MAX_RULES 256      /* ctts.c:25 */
#define MAX_REPLACE 256    /* ctts.c:26 */
#define MAX_CANDIDATES 64  /* ctts.c:1435 */

/* on-disk records, ctts.h:84-111 */
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

typedef struct {
    regex_t re;
    char* pat;   /* the ERE that was compiled (kept so that worker threads can compile their own copy) */
    char replace[MAX_REPLACE];
} norm_rule;

typedef enum { PH_VOWEL, PH_PLOSIVE, PH_FRICATIVE, PH_NASAL, PH_LIQUID, PH_OTHER } phoneme;
typedef enum { PT_DECL, PT_INTERROG, PT_EXCLAM, PT_CONT, PT_LISTING } phrase_type;

typedef struct {
    phrase_type type;
    float start, end, peak, peak_pos, energy;
} contour;

struct ctts_front {
    const uint8_t* db;
    size_t db_size;
    db_header hdr;
    const db_entry* index;
    const uint32_t* table;
    const char* strings;
    ctts_front_config cfg;
    norm_rule* rules;
    uint32_t n_rules;
    locale_t c_locale;
};

/* ------------------------------------------------------------------ config */

void ctts_front_config_defaults(ctts_front_config* c) {
    c->crossfade_ms = 20.0f;
    c->crossfade_vowel_ms = 45.0f;
    c->crossfade_s_ending_ms = 30.0f;
    c->crossfade_r_ending_ms = 30.0f;
    c->vowel_to_consonant_factor = 0.5f;
    c->word_pause_ms = 120.0f;
    c->unknown_silence_ms = 30.0f;
    c->fade_in_ms = 3.0f;
    c->fade_out_ms = 3.0f;
    c->remove_word_silence = 1;
    c->silence_threshold = 0.02f;
    c->min_silence_ms = 15.0f;
    c->remove_dc_offset = 1;
    c->normalize_level = 0.0f;
    c->compression = 0.0f;
    c->default_speed = 1.0f;
    c->min_speed = 0.5f;
    c->max_speed = 2.0f;
    c->max_pitch_change = 0.10f;
    c->print_units = 0;
    c->print_timing = 0;
}

static int truthy(const char* v) { return strcmp(v, "true") == 0 || strcmp(v, "1") == 0; }

/* One line of the flat scanner, ctts.c:1215-1292: first ':' splits key and
 * value, both trimmed of blanks; section headers fall through because no key
 * matches; inline comments are NOT stripped (strtof stops at them). */
static void config_line(ctts_front_config* c, const char* line) {
    while (*line == ' ' || *line == '\t') line++;
    if (*line == '#' || *line == '\0' || *line == '\n') return;
    const char* colon = strchr(line, ':');
    if (!colon) return;

    char key[64], val[64];
    size_t klen = (size_t)(colon - line);
    if (klen >= sizeof key) klen = sizeof key - 1;
    memcpy(key, line, klen);
    key[klen] = '\0';
    while (klen > 1 && (key[klen - 1] == ' ' || key[klen - 1] == '\t')) key[--klen] = '\0';

    const char* v = colon + 1;
    while (*v == ' ' || *v == '\t') v++;
    size_t vlen = strlen(v);
    if (vlen >= sizeof val) vlen = sizeof val - 1;
    memcpy(val, v, vlen);
    val[vlen] = '\0';
    while (vlen > 1 && strchr(" \t\n\r", val[vlen - 1])) val[--vlen] = '\0';

    static const struct { const char* name; size_t off; int is_bool; } keys[] = {
#define F(n) {#n, offsetof(ctts_front_config, n), 0}
#define B(n) {#n, offsetof(ctts_front_config, n), 1}
        F(crossfade_ms), F(crossfade_vowel_ms), F(crossfade_s_ending_ms),
        F(crossfade_r_ending_ms), F(vowel_to_consonant_factor), F(word_pause_ms),
        F(unknown_silence_ms), F(fade_in_ms), F(fade_out_ms), B(remove_word_silence),
        F(silence_threshold), F(min_silence_ms), B(remove_dc_offset), F(normalize_level),
        F(compression), F(default_speed), F(min_speed), F(max_speed), F(max_pitch_change),
        B(print_units), B(print_timing),
#undef F
#undef B
    };
    for (size_t i = 0; i < sizeof keys / sizeof keys[0]; i++) {
        if (strcmp(key, keys[i].name) != 0) continue;
        if (keys[i].is_bool) *(int*)((char*)c + keys[i].off) = truthy(val);
        else *(float*)((char*)c + keys[i].off) = strtof(val, NULL);
        return;
    }
}

int ctts_front_config_load(ctts_front_config* c, const char* path) {
    ctts_front_config_defaults(c);
    FILE* f = path ? fopen(path, "r") : NULL;
    if (!f) return CTTS_FRONT_OK;
    char line[256]; /* ctts.c:1304: longer lines are scanned in 255-byte pieces */
    while (fgets(line, sizeof line, f)) config_line(c, line);
    fclose(f);
    return CTTS_FRONT_OK;
}

/* ------------------------------------------------------------------- UTF-8 */

static int u8_len(const char* s) { /* utf8_char_len, ctts.c:211 */
    unsigned char c = (unsigned char)*s;
    if (c < 0x80) return 1;
    if ((c & 0xE0) == 0xC0) return 2;
    if ((c & 0xF0) == 0xE0) return 3;
    if ((c & 0xF8) == 0xF0) return 4;
    return 1;
}

/* Step over one character without ever passing the terminating NUL (the
 * reference trusts its input to be well-formed UTF-8). */
static const char* u8_step(const char* s) {
    int n = u8_len(s);
    while (n-- > 0 && *s) s++;
    return s;
}

static uint32_t u8_decode(const char** sp) { /* ctts_utf8_next, ctts.c:183 */
    const unsigned char* s = (const unsigned char*)*sp;
    uint32_t cp;
    if (*s < 0x80) {
        cp = *s++;
    } else if ((*s & 0xE0) == 0xC0) {
        cp = (uint32_t)(*s++ & 0x1F) << 6;
        if ((*s & 0xC0) == 0x80) cp |= *s++ & 0x3F;
    } else if ((*s & 0xF0) == 0xE0) {
        cp = (uint32_t)(*s++ & 0x0F) << 12;
        if ((*s & 0xC0) == 0x80) cp |= (uint32_t)(*s++ & 0x3F) << 6;
        if ((*s & 0xC0) == 0x80) cp |= *s++ & 0x3F;
    } else if ((*s & 0xF8) == 0xF0) {
        cp = (uint32_t)(*s++ & 0x07) << 18;
        if ((*s & 0xC0) == 0x80) cp |= (uint32_t)(*s++ & 0x3F) << 12;
        if ((*s & 0xC0) == 0x80) cp |= (uint32_t)(*s++ & 0x3F) << 6;
        if ((*s & 0xC0) == 0x80) cp |= *s++ & 0x3F;
    } else {
        cp = '?';
        s++;
    }
    *sp = (const char*)s;
    return cp;
}

static int u8_encode(uint32_t cp, char* out) { /* ctts.c:249 */
    if (cp < 0x80) {
        out[0] = (char)cp;
        return 1;
    }
    if (cp < 0x800) {
        out[0] = (char)(0xC0 | (cp >> 6));
        out[1] = (char)(0x80 | (cp & 0x3F));
        return 2;
    }
    if (cp < 0x10000) {
        out[0] = (char)(0xE0 | (cp >> 12));
        out[1] = (char)(0x80 | ((cp >> 6) & 0x3F));
        out[2] = (char)(0x80 | (cp & 0x3F));
        return 3;
    }
    out[0] = (char)(0xF0 | (cp >> 18));
    out[1] = (char)(0x80 | ((cp >> 12) & 0x3F));
    out[2] = (char)(0x80 | ((cp >> 6) & 0x3F));
    out[3] = (char)(0x80 | (cp & 0x3F));
    return 4;
}

static uint32_t last_codepoint(const char* text, size_t len) {
    const char* p = text;
    const char* last = text;
    while (p < text + len) {
        last = p;
        p += u8_len(p);
    }
    return u8_decode(&last);
}

static int is_vowel(uint32_t cp) { /* ctts.c:3042-3064 */
    switch (cp) {
        case 'a': case 'e': case 'i': case 'o': case 'u':
        case 'A': case 'E': case 'I': case 'O': case 'U':
        case 0xE1: case 0xC1: case 0xE0: case 0xC0: case 0xE2: case 0xC2: case 0xE3: case 0xC3:
        case 0xE9: case 0xC9: case 0xEA: case 0xCA: case 0xED: case 0xCD:
        case 0xF3: case 0xD3: case 0xF4: case 0xD4: case 0xF5: case 0xD5:
        case 0xFA: case 0xDA: case 0xFC: case 0xDC:
            return 1;
        default:
            return 0;
    }
}

static char ascii_lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }

/* ------------------------------------------------------- text normalisation */

static const char* const k_units[] = {"", "um", "dois", "três", "quatro", "cinco", "seis", "sete",
                                      "oito", "nove", "dez", "onze", "doze", "treze", "quatorze",
                                      "quinze", "dezesseis", "dezessete", "dezoito", "dezenove"};
static const char* const k_tens[] = {"", "", "vinte", "trinta", "quarenta", "cinquenta",
                                     "sessenta", "setenta", "oitenta", "noventa"};
static const char* const k_hundreds[] = {"", "cento", "duzentos", "trezentos", "quatrocentos",
                                         "quinhentos", "seiscentos", "setecentos", "oitocentos",
                                         "novecentos"};

typedef struct {
    char* p;
    size_t len, cap;
} strbuf;

static void sb_put(strbuf* b, const char* s) {
    size_t n = strlen(s);
    if (b->len + n > b->cap) n = b->cap - b->len;
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = '\0';
}

/* 0..999 in words, ctts.c:541-575 (temp[64] there: a chunk never exceeds it) */
static void words_0_999(int n, strbuf* b) {
    if (n == 0) { sb_put(b, "zero"); return; }
    if (n == 100) { sb_put(b, "cem"); return; }
    int h = n / 100, rest = n % 100;
    if (h > 0) sb_put(b, k_hundreds[h]);
    if (rest > 0) {
        if (h > 0) sb_put(b, " e ");
        if (rest < 20) {
            sb_put(b, k_units[rest]);
        } else {
            sb_put(b, k_tens[rest / 10]);
            if (rest % 10 > 0) {
                sb_put(b, " e ");
                sb_put(b, k_units[rest % 10]);
            }
        }
    }
}

/* full_number_to_words_pt, ctts.c:578-639.  The reference indexes
 * hundreds_pt[] out of bounds for values >= 10^12; we require < 10^12 and
 * otherwise spell the digits' value modulo 10^12. */
static void number_words(long long n, strbuf* b) {
    if (n == 0) { sb_put(b, "zero"); return; }
    n %= 1000000000000LL;
    if (n >= 1000000000LL) {
        int q = (int)(n / 1000000000LL);
        words_0_999(q, b);
        sb_put(b, q == 1 ? " bilhão" : " bilhões");
        n %= 1000000000LL;
        if (n > 0) sb_put(b, " e ");
    }
    if (n >= 1000000LL) {
        int q = (int)(n / 1000000LL);
        words_0_999(q, b);
        sb_put(b, q == 1 ? " milhão" : " milhões");
        n %= 1000000LL;
        if (n > 0) sb_put(b, " e ");
    }
    if (n >= 1000) {
        int q = (int)(n / 1000);
        if (q != 1) {
            words_0_999(q, b);
            sb_put(b, " mil");
        } else {
            sb_put(b, "mil");
        }
        n %= 1000;
        if (n > 0) sb_put(b, n < 100 ? " e " : " ");
    }
    if (n > 0) words_0_999((int)n, b);
}

/* expand_numbers, ctts.c:642-681 */
static char* expand_numbers(const char* text) {
    size_t cap = strlen(text) * 20 + 1024;
    char* res = malloc(cap);
    if (!res) return NULL;
    size_t room = cap - 1, w = 0;
    const char* s = text;
    while (*s && room > 0) {
        if (*s >= '0' && *s <= '9') {
            unsigned long long v = 0;
            while (*s >= '0' && *s <= '9') v = v * 10 + (unsigned)(*s++ - '0');
            char words[256]; /* ctts.c:664 */
            strbuf b = {words, 0, sizeof words - 1};
            words[0] = '\0';
            number_words((long long)(v % 1000000000000ULL), &b);
            size_t n = b.len > room ? room : b.len;
            memcpy(res + w, words, n);
            w += n;
            room -= n;
        } else {
            res[w++] = *s++;
            room--;
        }
    }
    res[w] = '\0';
    return res;
}

/* `\b` -> `[[:<:]]` before an alphanumeric / '[' / '(' else `[[:>:]]`
 * (convert_word_boundaries, ctts.c:294-340).  glibc's regcomp rejects both,
 * so such rules are dropped on Linux exactly as in the reference. */
static char* bsd_word_boundaries(const char* pat) {
    size_t n = 0;
    for (const char* p = pat; (p = strstr(p, "\\b")) != NULL; p += 2) n++;
    char* out = malloc(strlen(pat) + n * 5 + 1);
    if (!out) return NULL;
    char* d = out;
    for (const char* s = pat; *s;) {
        if (s[0] == '\\' && s[1] == 'b') {
            char nx = s[2];
            int starts = (nx >= 'a' && nx <= 'z') || (nx >= 'A' && nx <= 'Z') ||
                         (nx >= '0' && nx <= '9') || nx == '[' || nx == '(';
            memcpy(d, starts ? "[[:<:]]" : "[[:>:]]", 7);
            d += 7;
            s += 2;
        } else {
            *d++ = *s++;
        }
    }
    *d = '\0';
    return out;
}

/* ctts_load_normalization, ctts.c:343-408 */
static int load_rules(ctts_front* f, const char* path) {
    f->rules = NULL;
    f->n_rules = 0;
    FILE* fp = path ? fopen(path, "r") : NULL;
    if (!fp) return CTTS_FRONT_OK;
    f->rules = calloc(MAX_RULES, sizeof(norm_rule));
    if (!f->rules) {
        fclose(fp);
        return CTTS_FRONT_ERR_OUT_OF_MEMORY;
    }
    locale_t prev = uselocale(f->c_locale);
    char line[512];
    while (fgets(line, sizeof line, fp) && f->n_rules < MAX_RULES) {
        size_t len = strlen(line);
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = '\0';
        if (len == 0 || line[0] == '#') continue;
        char* comma = strchr(line, ',');
        if (!comma) continue;
        *comma = '\0';
        char* pat = bsd_word_boundaries(line);
        if (!pat) continue;
        norm_rule* r = &f->rules[f->n_rules];
        int bad = regcomp(&r->re, pat, REG_EXTENDED);
        if (bad) {
            free(pat);
            continue;
        }
        r->pat = pat;
        strncpy(r->replace, comma + 1, MAX_REPLACE - 1);
        r->replace[MAX_REPLACE - 1] = '\0';
        f->n_rules++;
    }
    uselocale(prev);
    fclose(fp);
    return CTTS_FRONT_OK;
}

/* apply_replacement, ctts.c:411-436 */
static size_t put_replacement(char* dst, size_t room, const char* rep, const char* src,
                              const regmatch_t* m, size_t nm) {
    size_t w = 0;
    while (*rep && w < room) {
        if (rep[0] == '\\' && rep[1] >= '0' && rep[1] <= '9') {
            size_t g = (size_t)(rep[1] - '0');
            if (g < nm && m[g].rm_so >= 0) {
                size_t gl = (size_t)(m[g].rm_eo - m[g].rm_so);
                if (gl > room - w) gl = room - w;
                memcpy(dst + w, src + m[g].rm_so, gl);
                w += gl;
            }
            rep += 2;
        } else {
            dst[w++] = *rep++;
        }
    }
    return w;
}

/* ctts_apply_normalization, ctts.c:439-505: every rule is applied left to
 * right over the whole string; each regexec restarts on the remaining suffix
 * (no REG_NOTBOL); a zero-length match drops one input byte (ctts.c:485). */
static char* apply_rules(ctts_front* f, const char* text) {
    if (f->n_rules == 0) return strdup(text);
    size_t cap = strlen(text) * 4 + 1024;
    char* cur = malloc(cap);
    char* nxt = malloc(cap);
    if (!cur || !nxt) {
        free(cur);
        free(nxt);
        return strdup(text);
    }
    strcpy(cur, text);
    locale_t prev = uselocale(f->c_locale);
    for (uint32_t i = 0; i < f->n_rules; i++) {
        regmatch_t m[10];
        char* s = cur;
        char* d = nxt;
        size_t room = cap - 1;
        while (*s && room > 0) {
            if (regexec(&f->rules[i].re, s, 10, m, 0) != 0 || m[0].rm_so < 0) {
                size_t rest = strlen(s);
                if (rest > room) rest = room;
                memcpy(d, s, rest);
                d += rest;
                break;
            }
            size_t before = (size_t)m[0].rm_so;
            if (before > room) before = room;
            memcpy(d, s, before);
            d += before;
            room -= before;
            size_t w = put_replacement(d, room, f->rules[i].replace, s, m, 10);
            d += w;
            room -= w;
            s += m[0].rm_eo;
            if (m[0].rm_eo == 0) s++;
        }
        *d = '\0';
        char* t = cur;
        cur = nxt;
        nxt = t;
    }
    uselocale(prev);
    free(nxt);
    return cur;
}

/* ctts_normalize, ctts.c:271 with unicode_tolower :238 */
static char* lowercase(const char* text) {
    char* res = malloc(strlen(text) * 4 + 1);
    if (!res) return NULL;
    char* d = res;
    for (const char* s = text; *s;) {
        uint32_t cp = u8_decode(&s);
        if (cp >= 'A' && cp <= 'Z') cp += 32;
        else if (cp == 0xC9) cp = 0xE9;
        else if (cp == 0xD3) cp = 0xF3;
        else if (cp == 0xD4) cp = 0xF4;
        else if (cp == 0xC7) cp = 0xE7;
        d += u8_encode(cp, d);
    }
    *d = '\0';
    return res;
}

char* ctts_front_normalize_text(ctts_front* f, const char* text) {
    char* a = expand_numbers(text);
    if (!a) return NULL;
    char* b = apply_rules(f, a);
    free(a);
    if (!b) return NULL;
    char* c = lowercase(b);
    free(b);
    return c;
}

void ctts_front_free(void* p) { free(p); }

/* ----------------------------------------------------------------- prosody */

static float clamp_pitch(float p, float lim) { /* ctts.c:2589 */
    float lo = 1.0f - lim, hi = 1.0f + lim;
    if (p < lo) return lo;
    if (p > hi) return hi;
    return p;
}

/* get_phrase_intonation + scale_intonation_to_limit, ctts.c:2611-2728 */
static contour phrase_contour(phrase_type t, float lim) {
    contour c;
    c.type = t;
    switch (t) {
        case PT_INTERROG:
            c.start = 0.98f; c.end = 1.08f; c.peak = 1.18f; c.peak_pos = 0.75f; c.energy = 1.05f;
            break;
        case PT_EXCLAM:
            c.start = 1.18f; c.end = 0.88f; c.peak = 1.22f; c.peak_pos = 0.15f; c.energy = 1.25f;
            break;
        case PT_CONT:
            c.start = 1.0f; c.end = 1.12f; c.peak = 1.08f; c.peak_pos = 0.7f; c.energy = 0.95f;
            break;
        case PT_LISTING:
            c.start = 1.0f; c.end = 1.06f; c.peak = 1.12f; c.peak_pos = 0.55f; c.energy = 1.0f;
            break;
        default:
            c.start = 1.04f; c.end = 0.88f; c.peak = 1.04f; c.peak_pos = 0.08f; c.energy = 1.0f;
            break;
    }
    if (lim > 0.0f) {
        float dev = fabsf(c.start - 1.0f);
        float d2 = fabsf(c.end - 1.0f), d3 = fabsf(c.peak - 1.0f);
        if (d2 > dev) dev = d2;
        if (d3 > dev) dev = d3;
        if (dev > lim) {
            float k = lim / dev;
            c.start = 1.0f + (c.start - 1.0f) * k;
            c.end = 1.0f + (c.end - 1.0f) * k;
            c.peak = 1.0f + (c.peak - 1.0f) * k;
        }
    }
    return c;
}

/* analyze_prosody, ctts.c:2883-2933: words and final punctuation of the
 * ORIGINAL text (before number expansion). */
static contour analyze_text(const char* text, float lim, int* word_count) {
    int words = 0, inside = 0;
    size_t len = strlen(text);
    for (size_t i = 0; i < len; i++) {
        if (text[i] == ' ' || text[i] == '\t' || text[i] == '\n') inside = 0;
        else if (!inside) {
            inside = 1;
            words++;
        }
    }
    phrase_type t = PT_DECL;
    for (size_t i = len; i > 0; i--) {
        char c = text[i - 1];
        if (c == '?') { t = PT_INTERROG; break; }
        if (c == '!') { t = PT_EXCLAM; break; }
        if (c == ',' || c == ';') { t = PT_CONT; break; }
        if (c != ' ' && c != '\t' && c != '\n') break;
    }
    *word_count = words;
    return phrase_contour(t, lim);
}

static float smoothstep(float t) { return t * t * (3.0f - 2.0f * t); }

/* Scalar half of apply_phrase_intonation, ctts.c:2740-2855: everything that
 * depends only on (word_index, total_words, contour, max_pitch_change).  The
 * sample count tests (count < 100, the 60/40 circumflex split) stay with the
 * executor. */
static void word_end_op(const contour* in, int wi, int total, float lim, int trim,
                        ctts_plan_op* op) {
    memset(op, 0, sizeof *op);
    op->kind = CTTS_OP_WORD_END;
    op->flags = trim ? CTTS_WE_TRIM : 0;
    if (total == 0) return;
    op->flags |= CTTS_WE_INTON;

    float pos = (float)wi / (float)(total > 1 ? total - 1 : 1);
    int final_word = (wi == total - 1);
    int penultimate = (wi == total - 2) && (total > 1);
    float pf;
    if (pos <= in->peak_pos) {
        float t = smoothstep(pos / in->peak_pos);
        pf = in->start + (in->peak - in->start) * t;
    } else {
        float t = smoothstep((pos - in->peak_pos) / (1.0f - in->peak_pos));
        pf = in->peak + (in->end - in->peak) * t;
    }
    pf = clamp_pitch(pf, lim);

    float ws, we;
    if (in->type == PT_INTERROG && (final_word || penultimate)) {
        if (final_word) {
            ws = clamp_pitch(pf * 0.95f, lim);
            we = clamp_pitch(in->end, lim);
            op->flags |= CTTS_WE_CIRCUMFLEX;
            op->f2 = clamp_pitch(in->peak, lim);
        } else {
            ws = clamp_pitch(pf * 0.98f, lim);
            we = clamp_pitch(pf * 1.05f, lim);
        }
    } else if (in->type == PT_EXCLAM) {
        if (wi == 0) {
            ws = clamp_pitch(in->peak, lim);
            we = clamp_pitch(pf, lim);
        } else if (final_word) {
            ws = clamp_pitch(pf, lim);
            we = clamp_pitch(in->end, lim);
        } else {
            ws = clamp_pitch(pf * 1.02f, lim);
            we = clamp_pitch(pf * 0.98f, lim);
        }
    } else if (in->type == PT_CONT && final_word) {
        ws = clamp_pitch(pf * 0.96f, lim);
        we = clamp_pitch(in->end, lim);
    } else {
        ws = clamp_pitch(pf * 0.98f, lim);
        we = clamp_pitch(pf * 1.02f, lim);
        if (final_word) we = clamp_pitch(in->end, lim);
    }
    op->f0 = ws;
    op->f1 = we;

    if (fabsf(in->energy - 1.0f) > 0.01f) {
        op->flags |= CTTS_WE_ENERGY;
        op->e0 = in->energy;
        op->e1 = in->energy;
        if (in->type == PT_EXCLAM && wi == 0) {
            op->e0 = in->energy * 1.1f;
            op->e1 = in->energy * 0.95f;
        }
    }
}

/* ---------------------------------------------------------- unit selection */

static uint32_t fnv1a(const char* s, size_t n) { /* ctts_hash, ctts.c:224 */
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) {
        h ^= (unsigned char)s[i];
        h *= 16777619u;
    }
    return h;
}

static int lookup(const ctts_front* f, const char* s, size_t n) { /* find_unit, ctts.c:1337 */
    uint32_t h = fnv1a(s, n);
    for (uint32_t i = f->table[h % f->hdr.hash_table_size]; i != NO_UNIT; i = f->index[i].next_hash) {
        if (i >= f->hdr.unit_count) return -1; /* corrupt chain */
        const db_entry* e = &f->index[i];
        if (e->hash == h && e->string_len == n && memcmp(f->strings + e->string_offset, s, n) == 0)
            return (int)i;
    }
    return -1;
}

/* start of the last character of [pos, end) */
static const char* back_one(const char* pos, const char* end) {
    const char* last = pos;
    for (const char* s = pos; s < end;) {
        last = s;
        s += u8_len(s);
    }
    return last;
}

/* find_longest_match, ctts.c:1357-1387.  Returns BYTES (the caller adds it to
 * a character count, as the reference does). */
static size_t longest_match_bytes(const ctts_front* f, const char* pos, size_t max_chars) {
    size_t remaining = strlen(pos); /* bytes, compared against a char count: ctts.c:1358-1360 */
    size_t chars = max_chars < remaining ? max_chars : remaining;
    const char* end = pos;
    for (size_t c = 0; c < chars && *end; c++) end = u8_step(end);
    while (end > pos) {
        if (lookup(f, pos, (size_t)(end - pos)) >= 0) return (size_t)(end - pos);
        end = back_one(pos, end);
    }
    return 0;
}

static int is_digraph(char a, char b) { /* is_pt_digraph, ctts.c:3146 */
    a = ascii_lower(a);
    b = ascii_lower(b);
    return (b == 'h' && (a == 'c' || a == 'l' || a == 'n')) || (b == 'u' && (a == 'q' || a == 'g'));
}

static int is_onset_cluster(char a, char b) { /* is_pt_valid_cluster, ctts.c:3167 */
    a = ascii_lower(a);
    b = ascii_lower(b);
    if (b == 'r') return a && strchr("pbtdcgfv", a) != NULL;
    if (b == 'l') return a && strchr("pbcgf", a) != NULL;
    return 0;
}

static int is_consonant(uint32_t cp) { /* is_pt_consonant, ctts.c:3138 */
    if (cp >= 'A' && cp <= 'Z') cp += 32;
    if (cp == 0xC7) cp = 0xE7;
    return (cp >= 'a' && cp <= 'z' && !is_vowel(cp)) || cp == 0xE7;
}

/* pt_reject_single_consonant, ctts.c:3193 */
static int reject_lone_consonant(const char* pos, size_t chars, int word_start) {
    if (chars != 1) return 0;
    const char* p = pos;
    uint32_t cp = u8_decode(&p);
    if (is_vowel(cp)) return 0;
    if (word_start) return 1;
    if (*p) {
        char a = (cp >= 'A' && cp <= 'Z') ? (char)(cp + 32) : (char)cp;
        if (is_digraph(a, *p)) return 1;
    }
    return 0;
}

/* pt_syllable_score, ctts.c:3220 */
static int syllable_score(const char* text, size_t bytes, size_t chars, int word_start) {
    if (chars == 0) return -1000;
    int score = (int)chars * 10;
    const char* p = text;
    uint32_t first = u8_decode(&p);
    int cons = is_consonant(first);
    if (chars >= 2 && bytes >= 2) {
        if (is_digraph(text[0], text[1])) score += 20;
        if (cons && is_onset_cluster(text[0], text[1])) score += 15;
    }
    if (word_start && cons) {
        if (chars == 1) score -= 100;
        else if (*p) {
            uint32_t second = u8_decode(&p);
            if (is_vowel(second)) score += 25;
        }
    }
    if (is_vowel(last_codepoint(text, bytes))) score += 10;
    return score;
}

typedef struct {
    size_t bytes, chars, next_bytes;
    int unit, score;
} candidate;

/* find_best_match_with_lookahead, ctts.c:1406-1554.  Returns matched bytes
 * (0 = miss) and the unit index. */
static size_t select_unit(const ctts_front* f, const char* pos, int word_start, int* unit) {
    *unit = -1;
    if (!*pos) return 0;
    size_t max_chars = f->hdr.max_unit_chars;
    size_t left = 0;
    for (const char* t = pos; *t; t = u8_step(t)) left++;
    size_t chars = max_chars < left ? max_chars : left;
    const char* end = pos;
    for (size_t c = 0; c < chars && *end; c++) end = u8_step(end);

    candidate cand[MAX_CANDIDATES];
    size_t n = 0;
    for (size_t cc = chars; end > pos && n < MAX_CANDIDATES; cc--) {
        size_t bytes = (size_t)(end - pos);
        int u = lookup(f, pos, bytes);
        if (u >= 0 && !reject_lone_consonant(pos, cc, word_start)) {
            cand[n].bytes = bytes;
            cand[n].chars = cc;
            cand[n].unit = u;
            cand[n].next_bytes = 0;
            cand[n].score = syllable_score(pos, bytes, cc, word_start);
            n++;
        }
        end = back_one(pos, end);
    }
    if (n == 0) return 0;
    if (n > 1) {
        for (size_t i = 0; i < n; i++) {
            const char* nx = pos + cand[i].bytes;
            while (*nx == ' ' || *nx == '\t' || *nx == '\n') nx++;
            if (*nx) cand[i].next_bytes = longest_match_bytes(f, nx, max_chars);
        }
    }
    /* ranking, ctts.c:1509-1550: score, then chars+next, then the four
     * end-of-word tie rules */
    size_t best = 0;
    int best_score = cand[0].score;
    size_t best_total = cand[0].chars + cand[0].next_bytes;
    for (size_t i = 1; i < n; i++) {
        size_t total = cand[i].chars + cand[i].next_bytes;
        if (cand[i].score > best_score) {
            best = i;
            best_score = cand[i].score;
            best_total = total;
        } else if (cand[i].score == best_score) {
            if (total > best_total) {
                best = i;
                best_total = total;
            } else if (total == best_total) {
                int b_end = cand[best].next_bytes == 0, c_end = cand[i].next_bytes == 0;
                if (!b_end && c_end) best = i;
                else if (b_end && c_end) {
                    if (cand[i].chars > cand[best].chars) best = i;
                } else if (!b_end && !c_end) {
                    if (cand[i].next_bytes > cand[best].next_bytes) best = i;
                }
            }
        }
    }
    *unit = cand[best].unit;
    return cand[best].bytes;
}

/* classify_first_phoneme / classify_last_phoneme, ctts.c:1775-1854 */
static phoneme consonant_class(char c) {
    if (c && strchr("ptkbdg", c)) return PH_PLOSIVE;
    if (c && strchr("fvszxj", c)) return PH_FRICATIVE;
    return PH_OTHER;
}

static phoneme first_phoneme(const char* t, size_t len) {
    if (len == 0) return PH_OTHER;
    char c = ascii_lower(t[0]);
    const char* p = t;
    if (is_vowel(u8_decode(&p))) return PH_VOWEL;
    phoneme k = consonant_class(c);
    if (k != PH_OTHER) return k;
    if (len >= 2 && c == 'c' && (t[1] == 'h' || t[1] == 'H')) return PH_FRICATIVE;
    if (c == 'm' || c == 'n') return PH_NASAL;
    if (c == 'l' || c == 'r') return PH_LIQUID;
    return PH_OTHER;
}

static phoneme last_phoneme(const char* t, size_t len) {
    if (len == 0) return PH_OTHER;
    if (is_vowel(last_codepoint(t, len))) return PH_VOWEL;
    char c = ascii_lower(t[len - 1]);
    if (len >= 2 && c == 'h') {
        char c2 = ascii_lower(t[len - 2]);
        if (c2 == 'l') return PH_LIQUID;
        if (c2 == 'n') return PH_NASAL;
        if (c2 == 'c') return PH_FRICATIVE;
    }
    phoneme k = consonant_class(c);
    if (k != PH_OTHER) return k;
    if (c == 'm' || c == 'n') return PH_NASAL;
    if (c == 'l' || c == 'r') return PH_LIQUID;
    return PH_OTHER;
}

/* get_adaptive_crossfade, ctts.c:1857-1892 */
static float join_crossfade_ms(phoneme prev_end, phoneme next_start, const ctts_front_config* c) {
    float base = c->crossfade_ms;
    if (next_start == PH_PLOSIVE) return base * 0.2f;
    if (prev_end == PH_PLOSIVE) return base * 0.3f;
    if (next_start == PH_FRICATIVE || prev_end == PH_FRICATIVE) return base * 0.4f;
    if (prev_end == PH_VOWEL && next_start == PH_VOWEL) return c->crossfade_vowel_ms;
    if (prev_end == PH_VOWEL) return base * c->vowel_to_consonant_factor;
    if (prev_end == PH_NASAL || prev_end == PH_LIQUID || next_start == PH_NASAL ||
        next_start == PH_LIQUID)
        return base * 0.7f;
    return base;
}

/* milliseconds -> samples exactly as the reference converts everywhere:
 * (size_t)(ms * 22050 / 1000.0f) in float arithmetic (e.g. ctts.c:3285) */
static uint32_t ms_to_samples(float ms) {
    return (uint32_t)(size_t)(ms * CTTS_PLAN_SAMPLE_RATE / 1000.0f);
}

/* get_punctuation_pause_ms, ctts.c:690-709 */
static float pause_ms(char c, const ctts_front_config* cfg) {
    switch (c) {
        case ',': return cfg->word_pause_ms * 1.8f;
        case ';': return cfg->word_pause_ms * 2.2f;
        case ':': return cfg->word_pause_ms * 2.0f;
        case '.': return cfg->word_pause_ms * 3.0f;
        case '!': return cfg->word_pause_ms * 3.2f;
        case '?': return cfg->word_pause_ms * 3.0f;
        default: return cfg->word_pause_ms;
    }
}

/* ------------------------------------------------------------- op emission */

typedef struct {
    ctts_plan_op* v;
    uint32_t n, cap;
    int oom;
} opvec;

static ctts_plan_op* push(opvec* o) {
    if (o->n == o->cap) {
        uint32_t nc = o->cap ? o->cap * 2 : 256;
        ctts_plan_op* nv = realloc(o->v, (size_t)nc * sizeof *nv);
        if (!nv) {
            o->oom = 1;
            return NULL;
        }
        o->v = nv;
        o->cap = nc;
    }
    ctts_plan_op* op = &o->v[o->n++];
    memset(op, 0, sizeof *op);
    return op;
}

static void emit_simple(opvec* o, uint16_t kind, uint32_t a) {
    ctts_plan_op* op = push(o);
    if (!op) return;
    op->kind = kind;
    op->a = a;
}

/* The main loop of ctts_synthesize, ctts.c:3689-3904, with each
 * sample-touching statement replaced by an op. */
static int plan_text(ctts_front* f, const char* text, opvec* o, uint32_t* found, uint32_t* missing) {
    const ctts_front_config* cfg = &f->cfg;
    int total_words = 0;
    contour in = analyze_text(text, cfg->max_pitch_change, &total_words);
    char* norm = ctts_front_normalize_text(f, text);
    if (!norm) return CTTS_FRONT_ERR_OUT_OF_MEMORY;

    const uint32_t word_pause = ms_to_samples(cfg->word_pause_ms);
    const uint32_t unknown_sil = ms_to_samples(cfg->unknown_silence_ms);
    const uint32_t fade_out = ms_to_samples(cfg->fade_out_ms);

    int boundary = 1, word_index = 0, prev_unit = -1;
    phoneme prev_end = PH_OTHER;
    *found = *missing = 0;

    for (const char* p = norm; *p;) {
        char c = *p;
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') { /* ctts.c:3691-3732 */
            ctts_plan_op* we = push(o);
            if (we) word_end_op(&in, word_index, total_words, cfg->max_pitch_change,
                                cfg->remove_word_silence, we);
            emit_simple(o, CTTS_OP_FADE_OUT, fade_out);
            emit_simple(o, CTTS_OP_SILENCE, word_pause);
            emit_simple(o, CTTS_OP_MARK, 0);
            word_index++;
            p++;
            boundary = 1;
            prev_unit = -1;
            prev_end = PH_OTHER;
        } else if (c == '-') { /* ctts.c:3736-3741 */
            p++;
        } else if (c == ',' || c == ';' || c == ':' || c == '.' || c == '!' || c == '?') {
            /* ctts.c:3744-3771 */
            uint32_t n = ms_to_samples(pause_ms(c, cfg));
            emit_simple(o, CTTS_OP_FADE_OUT, fade_out);
            if (n > 0) emit_simple(o, CTTS_OP_SILENCE, n);
            if (c == '.' || c == '!' || c == '?') {
                word_index = 0;
                emit_simple(o, CTTS_OP_MARK, 0);
            }
            p++;
            boundary = 1; /* prev_unit deliberately kept, as in the reference */
        } else if (c == '(' || c == ')' || c == '[' || c == ']' || c == '"' || c == '\'' ||
                   c == '`') { /* ctts.c:3774-3778 */
            p++;
        } else {
            int unit;
            size_t bytes = select_unit(f, p, boundary, &unit);
            if (bytes > 0 && unit >= 0) { /* ctts.c:3785-3861 */
                const db_entry* e = &f->index[unit];
                const char* ut = f->strings + e->string_offset;
                float xf_ms = cfg->crossfade_ms;
                if (!boundary && prev_unit >= 0) {
                    const db_entry* pe = &f->index[prev_unit];
                    xf_ms = join_crossfade_ms(prev_end, first_phoneme(ut, e->string_len), cfg);
                    uint32_t tail = last_codepoint(f->strings + pe->string_offset, pe->string_len);
                    if ((tail == 's' || tail == 'S') && pe->string_len > 0) {
                        if (xf_ms > cfg->crossfade_s_ending_ms) xf_ms = cfg->crossfade_s_ending_ms;
                    } else if ((tail == 'r' || tail == 'R') && pe->string_len > 0) {
                        if (xf_ms > cfg->crossfade_r_ending_ms) xf_ms = cfg->crossfade_r_ending_ms;
                    }
                }
                ctts_plan_op* op = push(o);
                if (op) {
                    op->kind = CTTS_OP_UNIT;
                    op->flags = boundary ? CTTS_UNIT_AFTER_BOUNDARY : 0;
                    op->a = (uint32_t)unit;
                    op->b = ms_to_samples(xf_ms);
                }
                prev_unit = unit;
                prev_end = last_phoneme(ut, e->string_len);
                boundary = 0;
                p += bytes;
                (*found)++;
            } else { /* ctts.c:3862-3870 */
                emit_simple(o, CTTS_OP_SILENCE, unknown_sil);
                p = u8_step(p);
                (*missing)++;
                prev_unit = -1;
                prev_end = PH_OTHER;
            }
        }
    }
    /* ctts.c:3877-3904: last region, then buffer_finalize */
    ctts_plan_op* we = push(o);
    if (we) word_end_op(&in, word_index, total_words, cfg->max_pitch_change,
                        cfg->remove_word_silence, we);
    if (fade_out > 0) emit_simple(o, CTTS_OP_FADE_OUT, fade_out);
    free(norm);
    return o->oom ? CTTS_FRONT_ERR_OUT_OF_MEMORY : CTTS_FRONT_OK;
}

/* --------------------------------------------------------------- public API */

int ctts_front_open(ctts_front** out, const void* voice_db, size_t db_size,
                    const ctts_front_config* cfg, const char* normalization_csv) {
    if (!out || !voice_db || db_size < sizeof(db_header)) return CTTS_FRONT_ERR_INVALID_ARG;
    ctts_front* f = calloc(1, sizeof *f);
    if (!f) return CTTS_FRONT_ERR_OUT_OF_MEMORY;
    f->db = voice_db;
    f->db_size = db_size;
    memcpy(&f->hdr, voice_db, sizeof f->hdr);
    if (f->hdr.magic != DB_MAGIC) { free(f); return CTTS_FRONT_ERR_INVALID_FORMAT; }
    if (f->hdr.version != DB_VERSION) { free(f); return CTTS_FRONT_ERR_VERSION; }
    const db_header* h = &f->hdr;
    if ((uint64_t)h->index_offset + (uint64_t)h->unit_count * sizeof(db_entry) > db_size ||
        (uint64_t)h->hash_table_offset + (uint64_t)h->hash_table_size * 4 > db_size ||
        h->strings_offset > db_size || h->hash_table_size == 0) {
        free(f);
        return CTTS_FRONT_ERR_INVALID_FORMAT;
    }
    f->index = (const db_entry*)(f->db + h->index_offset);
    f->table = (const uint32_t*)(f->db + h->hash_table_offset);
    f->strings = (const char*)(f->db + h->strings_offset);
    if (cfg) f->cfg = *cfg;
    else ctts_front_config_defaults(&f->cfg);
    f->c_locale = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    int err = load_rules(f, normalization_csv);
    if (err) {
        ctts_front_close(f);
        return err;
    }
    *out = f;
    return CTTS_FRONT_OK;
}

void ctts_front_close(ctts_front* f) {
    if (!f) return;
    for (uint32_t i = 0; i < f->n_rules; i++) {
        regfree(&f->rules[i].re);
        free(f->rules[i].pat);
    }
    free(f->rules);
    if (f->c_locale) freelocale(f->c_locale);
    free(f);
}

uint32_t ctts_front_rule_count(const ctts_front* f) { return f ? f->n_rules : 0; }
uint32_t ctts_front_unit_count(const ctts_front* f) { return f ? f->hdr.unit_count : 0; }
uint32_t ctts_front_max_unit_samples(const ctts_front* f) {
    uint32_t m = 0;
    for (uint32_t u = 0; f && u < f->hdr.unit_count; u++)
        if (f->index[u].sample_count > m) m = f->index[u].sample_count;
    return m;
}

void ctts_front_params(const ctts_front* f, ctts_assembly_params* out) {
    memset(out, 0, sizeof *out);
    out->fade_in_samples = ms_to_samples(f->cfg.fade_in_ms);
    out->min_silence_samples = ms_to_samples(f->cfg.min_silence_ms);
    out->silence_threshold = f->cfg.silence_threshold;
    out->target_rms = 3000.0f; /* ctts.c:3684 */
    out->remove_dc_offset = f->cfg.remove_dc_offset ? 1u : 0u;
}

int ctts_front_word_end_op(const ctts_front* f, int type_id, int word_index, int total_words,
                           ctts_plan_op* out) {
    if (!f || !out || type_id < 0 || type_id > (int)PT_LISTING) return CTTS_FRONT_ERR_INVALID_ARG;
    contour c = phrase_contour((phrase_type)type_id, f->cfg.max_pitch_change);
    word_end_op(&c, word_index, total_words, f->cfg.max_pitch_change, f->cfg.remove_word_silence, out);
    return CTTS_FRONT_OK;
}

/* One worker of ctts_front_plan_batch: plans texts [u0, u1) into its own op vector.  glibc's
 * regexec serialises on a lock inside the compiled pattern, so every worker compiles its own
 * copy of the rules; everything else in ctts_front is read-only after ctts_front_open. */
typedef struct {
    const ctts_front* f;
    const char* const* texts;
    uint32_t u0, u1;
    opvec o;
    uint32_t* n_ops;  /* per utterance (shared array, disjoint slices) */
    uint32_t* stats;  /* may be NULL */
    int err;
} plan_job;

static void* plan_worker(void* arg) {
    plan_job* j = arg;
    ctts_front local = *j->f;
    norm_rule* rules = NULL;
    if (local.n_rules) {
        rules = calloc(local.n_rules, sizeof *rules);
        if (!rules) {
            j->err = CTTS_FRONT_ERR_OUT_OF_MEMORY;
            return NULL;
        }
        locale_t prev = uselocale(local.c_locale);
        for (uint32_t i = 0; i < local.n_rules; i++) {
            rules[i] = j->f->rules[i];
            if (regcomp(&rules[i].re, j->f->rules[i].pat, REG_EXTENDED) != 0) j->err = CTTS_FRONT_ERR_INVALID_ARG;
        }
        uselocale(prev);
        local.rules = rules;
    }
    for (uint32_t u = j->u0; u < j->u1 && !j->err; u++) {
        uint32_t before = j->o.n, found = 0, missing = 0;
        j->err = j->texts[u] ? plan_text(&local, j->texts[u], &j->o, &found, &missing) : CTTS_FRONT_ERR_INVALID_ARG;
        if (!j->err && j->o.oom) j->err = CTTS_FRONT_ERR_OUT_OF_MEMORY;
        j->n_ops[u] = j->o.n - before;
        if (j->stats) {
            j->stats[2 * u] = found;
            j->stats[2 * u + 1] = missing;
        }
    }
    if (rules) {
        for (uint32_t i = 0; i < local.n_rules; i++) regfree(&rules[i].re);
        free(rules);
    }
    return NULL;
}

static uint32_t plan_threads(uint32_t n) {
    long t = sysconf(_SC_NPROCESSORS_ONLN);
    const char* e = getenv("CTTS_FRONT_THREADS");
    if (e && atoi(e) > 0) t = atoi(e);
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    uint32_t by_work = n / 64; /* below ~64 texts per worker the start-up (rule compilation) dominates */
    if (by_work < 1) by_work = 1;
    return (uint32_t)t < by_work ? (uint32_t)t : by_work;
}

int ctts_front_plan_batch(ctts_front* f, const char* const* texts, const float* speeds,
                          uint32_t n, ctts_batch_plan* out, uint32_t* stats) {
    return ctts_front_plan_batch_threads(f, texts, speeds, n, 0, out, stats);
}

int ctts_front_plan_batch_threads(ctts_front* f, const char* const* texts, const float* speeds,
                                  uint32_t n, uint32_t threads, ctts_batch_plan* out, uint32_t* stats) {
    if (!f || !out || (n && !texts)) return CTTS_FRONT_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    uint32_t* begin = malloc(((size_t)n + 1) * sizeof *begin);
    float* sp = malloc(((size_t)n + 1) * sizeof *sp);
    if (!begin || !sp) {
        free(begin);
        free(sp);
        return CTTS_FRONT_ERR_OUT_OF_MEMORY;
    }
    for (uint32_t u = 0; u < n; u++) sp[u] = speeds ? speeds[u] : 1.0f;

    /* utterances are independent: contiguous slices planned by worker threads, then concatenated */
    const uint32_t T = threads ? (threads > 64 ? 64 : threads) : plan_threads(n);
    plan_job* jobs = calloc(T, sizeof *jobs);
    pthread_t* tids = calloc(T, sizeof *tids);
    if (!jobs || !tids) {
        free(jobs);
        free(tids);
        free(begin);
        free(sp);
        return CTTS_FRONT_ERR_OUT_OF_MEMORY;
    }
    for (uint32_t t = 0; t < T; t++) {
        jobs[t].f = f;
        jobs[t].texts = texts;
        jobs[t].u0 = (uint32_t)((uint64_t)n * t / T);
        jobs[t].u1 = (uint32_t)((uint64_t)n * (t + 1) / T);
        jobs[t].n_ops = begin; /* per-utterance counts first, prefix-summed below */
        jobs[t].stats = stats;
    }
    int err = CTTS_FRONT_OK;
    if (T == 1) {
        plan_worker(&jobs[0]);
    } else {
        uint32_t started = 0;
        for (; started < T; started++)
            if (pthread_create(&tids[started], NULL, plan_worker, &jobs[started]) != 0) break;
        for (uint32_t t = started; t < T; t++) plan_worker(&jobs[t]); /* could not start: run inline */
        for (uint32_t t = 0; t < started; t++) pthread_join(tids[t], NULL);
    }
    uint64_t total = 0;
    for (uint32_t t = 0; t < T; t++) {
        if (jobs[t].err && !err) err = jobs[t].err;
        total += jobs[t].o.n;
    }
    ctts_plan_op* ops = NULL;
    if (!err && total > 0xffffffffull) err = CTTS_FRONT_ERR_INVALID_ARG;
    if (!err) {
        ops = malloc((total ? total : 1) * sizeof *ops);
        if (!ops) err = CTTS_FRONT_ERR_OUT_OF_MEMORY;
    }
    if (!err) {
        uint64_t at = 0;
        for (uint32_t t = 0; t < T; t++) {
            if (jobs[t].o.n) memcpy(ops + at, jobs[t].o.v, (size_t)jobs[t].o.n * sizeof *ops);
            at += jobs[t].o.n;
        }
        uint32_t run = 0;
        for (uint32_t u = 0; u < n; u++) {
            uint32_t c = begin[u];
            begin[u] = run;
            run += c;
        }
        begin[n] = run;
    }
    for (uint32_t t = 0; t < T; t++) free(jobs[t].o.v);
    free(jobs);
    free(tids);
    if (err) {
        free(begin);
        free(sp);
        free(ops);
        return err;
    }
    out->n_utts = n;
    out->n_ops = (uint32_t)total;
    out->utt_op_begin = begin;
    out->speed = sp;
    out->ops = ops;
    return CTTS_FRONT_OK;
}

void ctts_front_plan_free(ctts_batch_plan* p) {
    if (!p) return;
    free((void*)p->utt_op_begin);
    free((void*)p->speed);
    free((void*)p->ops);
    memset(p, 0, sizeof *p);
}

int ctts_front_plan_bounds(const ctts_front* f, const ctts_batch_plan* plan, uint64_t* pre,
                           uint64_t* out, uint32_t* region) {
    if (!f || !plan) return CTTS_FRONT_ERR_INVALID_ARG;
    for (uint32_t u = 0; u < plan->n_utts; u++) {
        uint64_t total = 0, cur = 0, longest = 0;
        for (uint32_t i = plan->utt_op_begin[u]; i < plan->utt_op_begin[u + 1]; i++) {
            const ctts_plan_op* op = &plan->ops[i];
            uint64_t add = 0;
            if (op->kind == CTTS_OP_UNIT) {
                if (op->a >= f->hdr.unit_count) return CTTS_FRONT_ERR_INVALID_ARG;
                add = f->index[op->a].sample_count;
            } else if (op->kind == CTTS_OP_SILENCE) {
                add = op->a;
            } else if (op->kind == CTTS_OP_MARK) {
                if (cur > longest) longest = cur;
                cur = 0;
            }
            total += add;
            cur += add;
        }
        if (cur > longest) longest = cur;
        if (pre) pre[u] = total;
        if (region) region[u] = (uint32_t)(longest > 0xFFFFFFFFu ? 0xFFFFFFFFu : longest);
        if (out) {
            float s = plan->speed[u];
            if (s == 1.0f) {
                out[u] = total;
            } else { /* time_stretch, ctts.c:3493-3517 */
                if (s < 0.5f) s = 0.5f;
                if (s > 2.0f) s = 2.0f;
                if (fabsf(s - 1.0f) < 0.01f) {
                    out[u] = total;
                } else {
                    size_t hop = (size_t)(128 / s);
                    if (hop < 1) hop = 1;
                    uint64_t frames = total > 512 ? (total - 512) / 128 + 1 : 1;
                    out[u] = frames * hop + 512;
                }
            }
        }
    }
    return CTTS_FRONT_OK;
}
