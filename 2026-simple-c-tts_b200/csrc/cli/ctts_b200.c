/*
 * ctts_b200 -- the `ctts synth` command line (ctts.c:3970-4030) over the B200 back end, in the
 * reference's own language: plain C on top of the two C-ABI libraries, no Python, no CUDA headers.
 *
 *   ctts_b200 synth <database.db> "text" <output.wav> [speed]
 *   ctts_b200 synth-batch <database.db> <texts.tsv> <out_dir>     lines: speed<TAB>text
 *
 * Like the reference it reads config.yaml and normalization.csv from the working directory
 * (ctts.c:3990, :3636), clamps the speed to [0.5, 2.0] (ctts.c:3976-3981) and takes default_speed
 * from the config when none is given (ctts.c:3993).  ctts_b200_synthesize_batch() below is the
 * glue INTEGRATION.md describes (texts -> ctts_b200_synth_texts: planner threads feeding a device
 * session -> per-utterance PCM); everything that touches samples runs on the GPU, and without a CUDA
 * device it fails.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "ctts_b200.h"

#define SAMPLE_RATE CTTS_PLAN_SAMPLE_RATE

typedef struct {
    void* db;
    size_t db_size;
    ctts_front_config cfg;
    ctts_front* front;
    ctts_gpu_ctx* gpu;
} engine;

static void engine_close(engine* e) {
    if (e->gpu) ctts_gpu_free(e->gpu);
    if (e->front) ctts_front_close(e->front);
    free(e->db);
    memset(e, 0, sizeof *e);
}

/* ctts_init (ctts.c:1117) + ctts_load_config("config.yaml") + the rule table, for both halves */
static int engine_open(engine* e, const char* db_path) {
    memset(e, 0, sizeof *e);
    FILE* f = fopen(db_path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n <= 0) { fclose(f); return -1; }
    e->db = malloc((size_t)n);
    if (!e->db || fread(e->db, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(e->db); e->db = NULL; return -1; }
    fclose(f);
    e->db_size = (size_t)n;
    ctts_front_config_defaults(&e->cfg);
    ctts_front_config_load(&e->cfg, "config.yaml");
    FILE* nf = fopen("normalization.csv", "rb");
    if (nf) fclose(nf);
    int rc = ctts_front_open(&e->front, e->db, e->db_size, &e->cfg, nf ? "normalization.csv" : NULL);
    if (rc) { engine_close(e); return rc; }
    const char* dev = getenv("CTTS_GPU_DEVICE");
    rc = ctts_gpu_init(&e->gpu, e->db, e->db_size, dev ? atoi(dev) : 0);
    if (rc) {
        fprintf(stderr, "ctts_gpu_init: %d (no CPU fallback)\n", rc);
        engine_close(e);
        return rc;
    }
    return 0;
}

/* ctts_write_wav, ctts.c:809: 44-byte RIFF/WAVE header (PCM, mono, 16 bit) + samples */
static int write_wav(const char* path, const int16_t* samples, size_t count) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    const uint32_t data = (uint32_t)(count * sizeof(int16_t)), riff = 36 + data, fmt = 16, sr = SAMPLE_RATE, br = SAMPLE_RATE * 2;
    const uint16_t pcm = 1, ch = 1, align = 2, bits = 16;
    fwrite("RIFF", 1, 4, f); fwrite(&riff, 4, 1, f); fwrite("WAVE", 1, 4, f);
    fwrite("fmt ", 1, 4, f); fwrite(&fmt, 4, 1, f);
    fwrite(&pcm, 2, 1, f); fwrite(&ch, 2, 1, f); fwrite(&sr, 4, 1, f); fwrite(&br, 4, 1, f);
    fwrite(&align, 2, 1, f); fwrite(&bits, 2, 1, f);
    fwrite("data", 1, 4, f); fwrite(&data, 4, 1, f);
    fwrite(samples, sizeof(int16_t), count, f);
    return fclose(f) == 0 ? 0 : -1;
}

/* N texts -> N PCM spans of one pinned buffer through ctts_b200_synth_texts (the front end's planner threads
 * feed the device piece by piece): *pcm (release with ctts_gpu_host_free), off[u] and cnt[u] (malloc'ed, n
 * entries each); stats may be NULL (2n: found, missing).  on_chunk (may be NULL) is called as soon as a range of
 * utterances is in host memory; it finds the three arrays through *pcm, *off, *cnt, which are set before the
 * device starts. */
static int ctts_b200_synthesize_batch(engine* e, const char* const* texts, const float* speeds, uint32_t n,
                                      int16_t** pcm, uint64_t** off, uint32_t** cnt, uint32_t* stats,
                                      ctts_gpu_chunk_fn on_chunk, void* user) {
    int rc = 0;
    const uint64_t cap = ctts_b200_capacity_hint(e->front, texts, speeds, n);   /* replaces the SampleBuffer growth policy */
    *off = malloc(sizeof **off * ((size_t)n + 1));
    *cnt = calloc(n ? n : 1, sizeof **cnt);
    *pcm = ctts_gpu_host_alloc(sizeof(int16_t) * (cap ? cap : 8));
    if (!*off || !*cnt || !*pcm) rc = CTTS_GPU_ERR_OUT_OF_MEMORY;
    ctts_b200_options opt;
    memset(&opt, 0, sizeof opt);
    opt.on_piece = on_chunk;
    opt.user = user;
    if (!rc) rc = ctts_b200_synth_texts(e->front, e->gpu, texts, speeds, n, *pcm, cap, *off, *cnt, stats, NULL, &opt, NULL);
    if (rc) {
        fprintf(stderr, "synthesis failed: %d %s\n", rc, ctts_gpu_last_error(e->gpu));
        ctts_gpu_host_free(*pcm);
        free(*off);
        free(*cnt);
        *pcm = NULL; *off = NULL; *cnt = NULL;
    }
    return rc;
}

static float clamp_speed(float s) { return s < 0.5f ? 0.5f : s > 2.0f ? 2.0f : s; }

static int cmd_synth(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "Usage: %s synth <database.db> \"text\" <output.wav> [speed]\n", argv[0]);
        return 1;
    }
    float speed = 1.0f;
    if (argc > 5) speed = clamp_speed(strtof(argv[5], NULL));
    engine e;
    if (engine_open(&e, argv[2])) {
        fprintf(stderr, "Failed to load database: %s\n", argv[2]);
        return 1;
    }
    if (argc <= 5 && e.cfg.default_speed != 1.0f) speed = e.cfg.default_speed;
    printf("Loaded database with %u units\n", ctts_front_unit_count(e.front));
    const char* texts[1] = {argv[3]};
    int16_t* pcm;
    uint64_t* off;
    uint32_t *cnt, stats[2] = {0, 0};
    if (ctts_b200_synthesize_batch(&e, texts, &speed, 1, &pcm, &off, &cnt, stats, NULL, NULL)) { engine_close(&e); return 1; }
    printf("Synthesized %u samples (%.2f seconds)\n", cnt[0], (float)cnt[0] / SAMPLE_RATE);
    printf("Units found: %u, missing: %u\n", stats[0], stats[1]);
    int rc = write_wav(argv[4], pcm + off[0], cnt[0]);
    if (rc) fprintf(stderr, "Failed to write WAV: %s\n", argv[4]);
    else printf("Written to %s\n", argv[4]);
    ctts_gpu_host_free(pcm);
    free(off);
    free(cnt);
    engine_close(&e);
    return rc ? 1 : 0;
}

/* synth-batch writes every WAV file as soon as its utterance has arrived (the device is still
 * working on later ones): state shared with the chunk callback */
typedef struct {
    const char* dir;
    uint32_t base;   /* first utterance of the group being synthesised */
    int16_t** pcm;
    uint64_t** off;
    uint32_t** cnt;
    double seconds;
    int rc;
} wav_sink;

static void write_chunk(void* user, uint32_t u0, uint32_t u1) {
    wav_sink* w = user;
    for (uint32_t u = u0; u < u1 && !w->rc; u++) {
        char path[4096];
        snprintf(path, sizeof path, "%s/%06u.wav", w->dir, w->base + u);
        w->rc = write_wav(path, *w->pcm + (*w->off)[u], (*w->cnt)[u]);
        w->seconds += (double)(*w->cnt)[u] / SAMPLE_RATE;
    }
}

static int cmd_synth_batch(int argc, char** argv) {
    if (argc != 5) {
        fprintf(stderr, "Usage: %s synth-batch <database.db> <texts.tsv> <out_dir>\n", argv[0]);
        return 1;
    }
    FILE* f = fopen(argv[3], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
    char** texts = NULL;
    float* speeds = NULL;
    uint32_t n = 0, cap = 0;
    char* line = NULL;
    size_t lcap = 0;
    ssize_t len;
    while ((len = getline(&line, &lcap, f)) > 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        char* tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        if (n == cap) {
            cap = cap ? 2 * cap : 1024;
            texts = realloc(texts, cap * sizeof *texts);
            speeds = realloc(speeds, cap * sizeof *speeds);
            if (!texts || !speeds) { fprintf(stderr, "out of memory\n"); return 1; }
        }
        speeds[n] = clamp_speed(strtof(line, NULL));
        texts[n] = strdup(tab + 1);
        n++;
    }
    free(line);
    fclose(f);
    engine e;
    if (engine_open(&e, argv[2])) {
        fprintf(stderr, "Failed to load database: %s\n", argv[2]);
        return 1;
    }
    if (mkdir(argv[4], 0777) != 0 && errno != EEXIST) { fprintf(stderr, "cannot create %s\n", argv[4]); engine_close(&e); return 1; }
    int16_t* pcm;
    uint64_t* off;
    uint32_t* cnt;
    wav_sink sink = {argv[4], 0, &pcm, &off, &cnt, 0.0, 0};
    /* groups of 1024 utterances: the pinned buffer is sized from the texts alone (ctts_b200_capacity_hint) */
    int rc = 0;
    for (uint32_t g0 = 0; g0 < n && !rc; g0 += 1024) {
        const uint32_t gn = n - g0 < 1024 ? n - g0 : 1024;
        sink.base = g0;
        rc = ctts_b200_synthesize_batch(&e, (const char* const*)texts + g0, speeds + g0, gn, &pcm, &off, &cnt, NULL, write_chunk, &sink);
        if (!rc) {
            rc = sink.rc;
            ctts_gpu_host_free(pcm);
            free(off);
            free(cnt);
        }
    }
    if (rc) fprintf(stderr, "Failed to write WAV files into %s\n", argv[4]);
    else printf("Synthesized %u utterances (%.1f seconds of audio) into %s\n", n, sink.seconds, argv[4]);
    for (uint32_t u = 0; u < n; u++) free(texts[u]);
    free(texts);
    free(speeds);
    engine_close(&e);
    return rc ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "synth") == 0) return cmd_synth(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "synth-batch") == 0) return cmd_synth_batch(argc, argv);
    fprintf(stderr, "Usage: %s synth <database.db> \"text\" <output.wav> [speed]\n       %s synth-batch <database.db> <texts.tsv> <out_dir>\n",
            argv[0], argv[0]);
    return 1;
}
