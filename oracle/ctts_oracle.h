/*
 * ctts_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C) of the reference's audio-assembly hot path, driven
 * by the same batch plan the GPU back end consumes.  It exists to check the
 * CUDA path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.  Pinned against the compiled reference (oracle/_ref, built
 * from /root/reference/ctts.c) by tests/test_oracle.py (test_oracle_vs_live_reference) and against
 * the committed vectors under tests/golden/.
 */
#ifndef CTTS_ORACLE_H
#define CTTS_ORACLE_H

#include "ctts_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ctts_oracle ctts_oracle;

#define CTTS_ORACLE_MAX_UB_SPANS 8

typedef struct ctts_oracle_stats {
    uint64_t pre_count;       /* samples in the assembly buffer before time stretching */
    uint64_t out_count;       /* samples returned */
    uint64_t trimmed;         /* samples removed by remove_silence_regions */
    uint32_t units;           /* CTTS_OP_UNIT executed */
    uint32_t joins;           /* units that went through smooth/match */
    uint32_t pitch_shifts;    /* joins that took the >15% pitch-jump branch */
    uint32_t contour_calls;   /* apply_smooth_pitch_contour bodies executed */
    uint32_t wsola_frames;
    /* The reference reads past the end of a heap copy in
     * apply_smooth_pitch_contour (ctts.c:2243-2252: idx = (size_t)(i*pf) can
     * reach 280 > frame, and pos+idx can pass `count` in the last frame).  What
     * it reads is heap garbage; the oracle reads 0 there and reports the spans
     * (pre-stretch sample indices, [begin,end)) of output samples whose value
     * depended on such a read, so parity tests can mask them. */
    uint32_t ub_spans;
    uint64_t ub_span[CTTS_ORACLE_MAX_UB_SPANS][2];
} ctts_oracle_stats;

/* voice.db bytes are borrowed (must outlive the handle).  Builds the fade
 * LUTs and Hann windows with the host libm exactly as ctts.c:60-73, :1624,
 * :2198-2204. */
int ctts_oracle_open(ctts_oracle** out, const void* voice_db, size_t db_size);
void ctts_oracle_close(ctts_oracle* o);

/* tables: 3x1024 fade LUTs (fade_out, fade_in, sine), hann256, hann512 */
void ctts_oracle_tables(const ctts_oracle* o, float* luts3x1024, float* hann256, float* hann512);

/* Execute one utterance's ops.  *out is malloc'ed (free with ctts_oracle_free).
 * If `pre` is not NULL it receives a malloc'ed copy of the pre-stretch buffer. */
int ctts_oracle_synth(const ctts_oracle* o, const ctts_assembly_params* prm,
                      const ctts_plan_op* ops, uint32_t n_ops, float speed, int16_t** out,
                      size_t* n_out, int16_t** pre, size_t* n_pre, ctts_oracle_stats* stats);
void ctts_oracle_free(void* p);

/* ---- stage functions (same arithmetic the executor uses), for unit tests ---- */
float ctts_oracle_rms(const int16_t* s, size_t n);
void ctts_oracle_normalize_rms(int16_t* s, size_t n, float target);
void ctts_oracle_remove_dc(int16_t* s, size_t n);
float ctts_oracle_estimate_pitch(const int16_t* s, size_t n);
int ctts_oracle_smooth_pitch(const int16_t* buf, size_t count, int16_t* unit, size_t n, size_t xf);
void ctts_oracle_match_energy(const int16_t* buf, size_t count, int16_t* unit, size_t n, size_t xf);
size_t ctts_oracle_trim(int16_t* s, size_t n, float thr, size_t min_sil);
/* returns number of out-of-bounds reads (reference UB) */
uint32_t ctts_oracle_contour(const ctts_oracle* o, int16_t* s, size_t n, float f0, float f1,
                             size_t* taint_lo, size_t* taint_hi);
void ctts_oracle_fade_in(const ctts_oracle* o, int16_t* s, size_t n, size_t f);
void ctts_oracle_fade_out(const ctts_oracle* o, int16_t* s, size_t n, size_t f);
/* buf must have room for count+n samples; returns the new count */
size_t ctts_oracle_append(const ctts_oracle* o, int16_t* buf, size_t count, int16_t* unit, size_t n,
                          size_t xf, size_t fade_in, int remove_dc, int after_boundary);
int ctts_oracle_wsola_offset(const int16_t* in, size_t n, const int16_t* prev_frame, size_t nominal);
int ctts_oracle_time_stretch(const ctts_oracle* o, const int16_t* in, size_t n, float speed,
                             int16_t** out, size_t* n_out, uint32_t* frames);

#ifdef __cplusplus
}
#endif

#endif /* CTTS_ORACLE_H */
