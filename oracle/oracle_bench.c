/*
 * oracle_bench.c -- multi-threaded timing driver around the CPU restatement.
 * TEST INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline / --impl reference legs.
 * Threading model = the reference's own: one decoder per thread, frames decoded independently
 * (thread_local Config/CCM, src/lib/cimb_translator/Config.h:11-15; TODO.md:14-16).
 */
#define _POSIX_C_SOURCE 200809L
#include "cimbar_oracle.h"

#include <malloc.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    const cbo_mode* mode; const uint8_t* frames; int nframes; int w, h; int tid, nthreads; int stage;
    uint8_t* out; size_t out_stride; uint64_t checksum;
} job_t;

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    size_t fsz = (size_t)j->w * j->h * 3;
    uint8_t* buf = (uint8_t*)malloc(16384);
    uint8_t ok[128];
    uint64_t cs = 0;
    for (int f = j->tid; f < j->nframes; f += j->nthreads) {
        const uint8_t* rgb = j->frames + (size_t)f * fsz;
        int n;
        if (j->stage == 0) n = cbo_decode_raw(j->mode, rgb, j->w, j->h, 0, 0, buf, NULL);   /* cell bits only */
        else n = cbo_decode(j->mode, rgb, j->w, j->h, 0, 1, buf, ok);                       /* + RS */
        for (int k = 0; k < n; ++k) cs = cs * 1099511628211ULL + buf[k];
        if (j->out) memcpy(j->out + (size_t)f * j->out_stride, buf, (size_t)n);
    }
    j->checksum = cs;
    free(buf);
    return NULL;
}

/* decode nframes frames with nthreads threads; stage 0 = raw cell bits, 1 = full decode incl. RS.
   returns elapsed seconds; out (optional) receives per-frame bytes at out_stride */
double cbo_bench_decode(int mode_val, const uint8_t* frames, int nframes, int w, int h, int nthreads, int stage,
                        uint8_t* out, size_t out_stride, uint64_t* checksum)
{
    cbo_mode mode; cbo_mode_init(&mode, mode_val);
    if (nthreads < 1) nthreads = 1;
    /* keep the per-frame MB-sized scratch buffers inside the per-thread malloc arenas: with the default
       thresholds every frame mmaps/munmaps ~5 MB, and the munmap TLB shootdowns serialise the threads */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 256);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].mode = &mode; jobs[t].frames = frames; jobs[t].nframes = nframes; jobs[t].w = w; jobs[t].h = h;
        jobs[t].tid = t; jobs[t].nthreads = nthreads; jobs[t].stage = stage; jobs[t].out = out; jobs[t].out_stride = out_stride;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    uint64_t cs = 0;
    for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); cs ^= jobs[t].checksum; }
    double t1 = now_s();
    if (checksum) *checksum = cs;
    free(th); free(jobs);
    return t1 - t0;
}
