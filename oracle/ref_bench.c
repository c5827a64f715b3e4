/*
 * ref_bench.c -- TEST / BASELINE INFRASTRUCTURE, not product code.
 *
 * CPU baseline runner: the UNMODIFIED reference (`ctts_synthesize` from
 * /root/reference/ctts.c, included at compile time, built with the
 * reference's own flags) looped in-process over a shard of a text file by N
 * worker PROCESSES (the reference is not thread-safe: global rule / LUT /
 * window tables, ctts.c:34-36, 55-58, 2195-2196).  Process start-up (mmap +
 * regcomp) is outside the timed region.
 *
 * usage: ctts_ref_bench <voice.db> <config.yaml|-> <normalization.csv|-> <texts.tsv> <nproc> [dump.bin]
 *   texts.tsv: one utterance per line, "<speed>\t<text>"
 *   dump.bin (optional, nproc must be 1): for each utterance, uint64 count + int16 PCM
 * prints one JSON line: {"utts":U,"samples":S,"seconds":T,"procs":N}
 */
#define main ctts_reference_main
#include "ctts.c"
#undef main

#include <sys/wait.h>
#include <time.h>

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    float speed;
    char* text;
} Utt;

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s db config norm texts.tsv nproc [dump.bin]\n", argv[0]);
        return 2;
    }
    const char* db = argv[1];
    const char* cfg = strcmp(argv[2], "-") ? argv[2] : NULL;
    const char* norm = strcmp(argv[3], "-") ? argv[3] : NULL;
    int nproc = atoi(argv[5]);
    const char* dump = argc > 6 ? argv[6] : NULL;
    if (nproc < 1) nproc = 1;
    if (dump && nproc != 1) {
        fprintf(stderr, "dump needs nproc=1\n");
        return 2;
    }

    FILE* f = fopen(argv[4], "r");
    if (!f) {
        perror("texts");
        return 2;
    }
    size_t cap = 1024, n = 0;
    Utt* utts = malloc(cap * sizeof(Utt));
    char* line = NULL;
    size_t line_cap = 0;
    ssize_t got;
    while ((got = getline(&line, &line_cap, f)) > 0) {
        while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) line[--got] = 0;
        char* tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        if (n == cap) {
            cap *= 2;
            utts = realloc(utts, cap * sizeof(Utt));
        }
        utts[n].speed = strtof(line, NULL);
        utts[n].text = strdup(tab + 1);
        n++;
    }
    fclose(f);

    int (*ready)[2] = malloc(sizeof(int[2]) * nproc);
    int (*go)[2] = malloc(sizeof(int[2]) * nproc);
    int (*res)[2] = malloc(sizeof(int[2]) * nproc);
    pid_t* pids = malloc(sizeof(pid_t) * nproc);

    for (int r = 0; r < nproc; r++) {
        if (pipe(ready[r]) || pipe(go[r]) || pipe(res[r])) {
            perror("pipe");
            return 2;
        }
        pid_t pid = fork();
        if (pid < 0) {
            perror("fork");
            return 2;
        }
        if (pid == 0) {
            CTTS* e = ctts_init(db);
            if (!e) _exit(3);
            if (cfg) ctts_load_config(&e->config, cfg);
            if (norm) ctts_load_normalization(norm);
            else norm_rules_loaded = 1;
            duration_rules_loaded = 1;
            /* silence the per-rule warnings already printed; keep going */
            char c = 'r';
            if (write(ready[r][1], &c, 1) != 1) _exit(4);
            if (read(go[r][0], &c, 1) != 1) _exit(4);
            FILE* df = dump ? fopen(dump, "wb") : NULL;
            uint64_t total = 0, done = 0;
            for (size_t i = (size_t)r; i < n; i += (size_t)nproc) {
                int16_t* s = NULL;
                size_t cnt = 0;
                int err = ctts_synthesize(e, utts[i].text, &s, &cnt, utts[i].speed);
                if (err != CTTS_OK) _exit(5);
                total += cnt;
                done++;
                if (df) {
                    uint64_t c64 = cnt;
                    fwrite(&c64, sizeof c64, 1, df);
                    fwrite(s, sizeof(int16_t), cnt, df);
                }
                free(s);
            }
            if (df) fclose(df);
            uint64_t out[2] = {done, total};
            if (write(res[r][1], out, sizeof out) != (ssize_t)sizeof out) _exit(4);
            _exit(0);
        }
        pids[r] = pid;
    }

    char c;
    for (int r = 0; r < nproc; r++)
        if (read(ready[r][0], &c, 1) != 1) {
            fprintf(stderr, "worker %d failed to start\n", r);
            return 3;
        }
    double t0 = now_s();
    c = 'g';
    for (int r = 0; r < nproc; r++)
        if (write(go[r][1], &c, 1) != 1) return 3;
    uint64_t utt_total = 0, samp_total = 0;
    for (int r = 0; r < nproc; r++) {
        uint64_t out[2];
        if (read(res[r][0], out, sizeof out) != (ssize_t)sizeof out) {
            fprintf(stderr, "worker %d died\n", r);
            return 3;
        }
        utt_total += out[0];
        samp_total += out[1];
    }
    double t1 = now_s();
    for (int r = 0; r < nproc; r++) waitpid(pids[r], NULL, 0);
    printf("{\"utts\": %llu, \"samples\": %llu, \"seconds\": %.6f, \"procs\": %d}\n",
           (unsigned long long)utt_total, (unsigned long long)samp_total, t1 - t0, nproc);
    return 0;
}
