/*
 * ctts_plan.h -- the batch plan: what crosses the boundary between the CTTS
 * host front end and the GPU audio-assembly back end.
 *
 * The reference has no such boundary: `ctts_synthesize` (ctts.c:3623-3924)
 * interleaves text decisions with sample loops.  The plan is that loop with
 * every sample-touching statement replaced by one fixed-size op.  Everything
 * that depends on sample data (buffer counts, silence trimming, pitch
 * decisions) is evaluated by the executor exactly where the reference
 * evaluates it; everything that depends only on text, the voice.db index and
 * config.yaml is resolved by the front end and stored in the op.
 *
 * Layout: CSR.  `utt_op_begin[u] .. utt_op_begin[u+1]` indexes `ops` for
 * utterance u; `speed[u]` is the (already clamped) speed argument.
 */
#ifndef CTTS_PLAN_H
#define CTTS_PLAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTTS_PLAN_SAMPLE_RATE 22050 /* ctts.h:24 */

/* Op kinds.  One per sample-touching statement of ctts.c:3689-3904. */
enum {
    /* unit hit, ctts.c:3785-3861: gather unit `a` from the PCM pool,
     * normalize_rms(3000) -> [smooth_pitch_boundary, match_boundary_energy
     * when !after_boundary && count>0] -> buffer_append_crossfade with
     * crossfade length `b` samples. */
    CTTS_OP_UNIT = 1,
    /* buffer_append_silence(a samples): word pause ctts.c:3720, punctuation
     * pause :3759, unknown character :3864. */
    CTTS_OP_SILENCE = 2,
    /* `if (buf.count > 0) apply_fade_out(buf.data, buf.count, a)`:
     * ctts.c:3716-3719, :3752-3755, and buffer_finalize :3904. */
    CTTS_OP_FADE_OUT = 3,
    /* end of a whitespace-delimited region, ctts.c:3693-3713 and :3878-3898:
     * remove_silence_regions on [word_start, count) when flags&TRIM, then
     * apply_phrase_intonation with the host-resolved scalar contour. */
    CTTS_OP_WORD_END = 4,
    /* word_start_sample = buf.count  (ctts.c:3723, :3765) */
    CTTS_OP_MARK = 5
};

/* flags of CTTS_OP_UNIT */
#define CTTS_UNIT_AFTER_BOUNDARY 1u /* prev_was_word_boundary, ctts.c:3845 */

/* flags of CTTS_OP_WORD_END */
#define CTTS_WE_TRIM        1u  /* config->remove_word_silence */
#define CTTS_WE_INTON       2u  /* total_words != 0 (ctts.c:2740) */
#define CTTS_WE_CIRCUMFLEX  4u  /* interrogative final word, ctts.c:2775-2790 */
#define CTTS_WE_ENERGY      8u  /* |energy_factor-1| > 0.01, ctts.c:2843 */

/* 32 bytes, no padding. */
typedef struct ctts_plan_op {
    uint16_t kind;   /* CTTS_OP_* */
    uint16_t flags;
    uint32_t a;      /* UNIT: unit index; SILENCE/FADE_OUT: samples */
    uint32_t b;      /* UNIT: crossfade samples = (size_t)(ms*22050/1000.0f) */
    float f0;        /* WORD_END: word_start pitch factor */
    float f1;        /* WORD_END: word_end pitch factor */
    float f2;        /* WORD_END: circumflex peak factor */
    float e0;        /* WORD_END: energy factor at the first sample */
    float e1;        /* WORD_END: energy factor at the last sample */
} ctts_plan_op;

typedef struct ctts_batch_plan {
    uint32_t n_utts;
    uint32_t n_ops;
    const uint32_t* utt_op_begin; /* n_utts + 1 */
    const float* speed;           /* n_utts */
    const ctts_plan_op* ops;      /* n_ops */
} ctts_batch_plan;

/* Scalar parameters of the assembly stage that are constant for a batch;
 * all come from CTTSConfig (ctts.h:44-77) via the conversions at
 * ctts.c:3285-3286, :3666-3667, :3687. */
typedef struct ctts_assembly_params {
    uint32_t fade_in_samples;      /* (size_t)(fade_in_ms*22050/1000.0f) */
    uint32_t min_silence_samples;  /* (size_t)(min_silence_ms*22050/1000.0f) */
    float silence_threshold;       /* config->silence_threshold */
    float target_rms;              /* 3000.0f, ctts.c:3684 */
    uint32_t remove_dc_offset;     /* config->remove_dc_offset */
    uint32_t reserved[3];
} ctts_assembly_params;

#ifdef __cplusplus
}
#endif

#endif /* CTTS_PLAN_H */
