/*
 * ctts_front.h -- host front end of the B200 CTTS back end: text -> batch plan.
 *
 * Plain C, no GPU.  It performs exactly the text-side decisions the reference
 * makes inside `ctts_synthesize` (ctts.c:3638-3655 text pipeline, :2883
 * prosody analysis, :1406 unit selection, :1857 adaptive crossfade, :690
 * punctuation pauses) and records them as plan ops (ctts_plan.h) instead of
 * touching samples.  Same voice.db / config.yaml / normalization.csv formats
 * as the reference.
 */
#ifndef CTTS_FRONT_H
#define CTTS_FRONT_H

#include "ctts_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same fields, order, types and defaults as CTTSConfig (ctts.h:44-77,
 * defaults ctts.c:1190-1212) so a reference config can be passed by pointer. */
typedef struct ctts_front_config {
    float crossfade_ms;
    float crossfade_vowel_ms;
    float crossfade_s_ending_ms;
    float crossfade_r_ending_ms;
    float vowel_to_consonant_factor;
    float word_pause_ms;
    float unknown_silence_ms;
    float fade_in_ms;
    float fade_out_ms;
    int remove_word_silence;
    float silence_threshold;
    float min_silence_ms;
    int remove_dc_offset;
    float normalize_level;
    float compression;
    float default_speed;
    float min_speed;
    float max_speed;
    float max_pitch_change;
    int print_units;
    int print_timing;
} ctts_front_config;

typedef struct ctts_front ctts_front;

/* error codes: the reference's (ctts.h:333-341) */
#define CTTS_FRONT_OK 0
#define CTTS_FRONT_ERR_INVALID_ARG (-1)
#define CTTS_FRONT_ERR_INVALID_FORMAT (-5)
#define CTTS_FRONT_ERR_OUT_OF_MEMORY (-6)
#define CTTS_FRONT_ERR_VERSION (-8)

/* ctts_config_defaults, ctts.c:1190 */
void ctts_front_config_defaults(ctts_front_config* cfg);
/* ctts_load_config, ctts.c:1294: flat "key: value" scanner; a missing file
 * leaves the defaults and returns 0. */
int ctts_front_config_load(ctts_front_config* cfg, const char* path);

/* Open a front end over voice.db bytes (borrowed: must outlive the handle;
 * only header, index, hash table and string pool are read -- ctts.c:1144-1158).
 * `normalization_csv` may be NULL (no rules), as when the file is absent
 * (ctts.c:349-354). */
int ctts_front_open(ctts_front** out, const void* voice_db, size_t db_size,
                    const ctts_front_config* cfg, const char* normalization_csv);
void ctts_front_close(ctts_front* f);

/* number of normalisation rules that compiled (7 of the 49 shipped rules on
 * glibc, which rejects the BSD-only [[:<:]] the reference emits) */
uint32_t ctts_front_rule_count(const ctts_front* f);
uint32_t ctts_front_unit_count(const ctts_front* f);
/* longest unit of the voice in samples (CTTSIndexEntry.sample_count, ctts.h:108) */
uint32_t ctts_front_max_unit_samples(const ctts_front* f);

/* batch-constant scalar parameters for the executor */
void ctts_front_params(const ctts_front* f, ctts_assembly_params* out);

/* expand_numbers -> rules -> lowercase (ctts.c:3643-3655); malloc'ed, free with
 * ctts_front_free. */
char* ctts_front_normalize_text(ctts_front* f, const char* text);
void ctts_front_free(void* p);

/* Plan a batch.  `out` receives malloc'ed CSR arrays (release with
 * ctts_front_plan_free).  speeds may be NULL (all 1.0).  `stats`, if not NULL,
 * receives 2*n uint32: units_found, units_missing per utterance
 * (engine->units_found/missing, ctts.c:3861, :3866). */
int ctts_front_plan_batch(ctts_front* f, const char* const* texts, const float* speeds,
                          uint32_t n, ctts_batch_plan* out, uint32_t* stats);
/* The same with an explicit number of worker threads (0: chosen from the batch size and the host's
 * cores; 1: the calling thread alone -- what a caller that runs its own pool of planners wants,
 * ctts_b200_synth_texts).  Re-entrant: the handle is only read. */
int ctts_front_plan_batch_threads(ctts_front* f, const char* const* texts, const float* speeds,
                                  uint32_t n, uint32_t threads, ctts_batch_plan* out, uint32_t* stats);
void ctts_front_plan_free(ctts_batch_plan* plan);

/* The WORD_END op the walk emits for word `word_index` of `total_words` in a
 * phrase of the given type (0 declarative, 1 interrogative, 2 exclamatory,
 * 3 continuation, 4 listing: PhraseType, ctts.c:2526): the scalar half of
 * apply_phrase_intonation (ctts.c:2740-2855).  Exposed for tests and tools. */
int ctts_front_word_end_op(const ctts_front* f, int phrase_type, int word_index, int total_words,
                           ctts_plan_op* out);

/* Host-known upper bounds on sample counts of utterance u of a plan:
 * pre[u]  >= samples in the assembly buffer before time stretching
 *            (sum of unit lengths + silences; crossfades and trimming only shrink it),
 * out[u]  >= samples returned (== pre bound at speed 1.0f, else the WSOLA
 *            bound num_frames*synthesis_hop + 512 from ctts.c:3515-3517),
 * region[u] >= longest stretch between two CTTS_OP_MARKs. */
int ctts_front_plan_bounds(const ctts_front* f, const ctts_batch_plan* plan,
                           uint64_t* pre, uint64_t* out, uint32_t* region);

#ifdef __cplusplus
}
#endif

#endif /* CTTS_FRONT_H */
