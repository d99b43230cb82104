/* nbody_viewer_feed -- the viewer hand-off of the reference, in C over the C ABI.
 *
 * The reference runs the simulation on its own thread and hands the state to the renderer under a lock
 * (main.cpp:612-635):
 *     simulation->step();
 *     { lock_guard lock(UPDATE_LOCK); SHARED_BODIES = simulation->bodies; SHARED_QUADTREE = ...; }
 *     sleep 1 ms
 * while the render loop copies SHARED_BODIES under the same lock once per frame (main.cpp:~700).  Here the
 * simulation thread is
 *     nbody_gpu_step(ctx, dt, 1);
 *     lock; nbody_gpu_download(ctx, SHARED_BODIES, n, POS | VEL); unlock;      (the copy IS the download)
 *     sleep 1 ms
 * and a consumer thread plays the render loop at 60 Hz: lock, copy SHARED_BODIES, unlock, "draw" (a bounding box
 * and a checksum, so that torn frames would show).  One context, used by the simulation thread only (nbody_gpu.h:
 * a context is used by one host thread at a time).  Prints one JSON line with the cadences and the lock waits.
 *
 *   nbody_viewer_feed [--seconds S] [--n N] [--sleep-ms M] [--fields pos|posvel|all]
 */
#define _POSIX_C_SOURCE 200809L
#include "nbody_gpu.h"
#include "nbody_host.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static void sleep_ms(double ms)
{
    struct timespec ts;
    ts.tv_sec = (time_t)(ms / 1000.0);
    ts.tv_nsec = (long)((ms - 1000.0 * (double)ts.tv_sec) * 1e6);
    nanosleep(&ts, NULL);
}

typedef struct {
    nbody_ctx *ctx;
    size_t n;
    float dt;
    unsigned fields;
    double sleep_ms, seconds;
    pthread_mutex_t lock;                 /* UPDATE_LOCK, main.cpp:41 */
    nbody_body_t *shared;                 /* SHARED_BODIES, main.cpp:43 */
    unsigned long long frame;             /* Simulation::frame of the copy in `shared` */
    volatile int stop, failed;
    /* statistics */
    unsigned long long steps, publishes, frames, torn;
    double step_s, publish_s, sim_lock_wait_max, view_lock_wait_max, view_copy_s;
} feed_t;

static void *simulation_thread(void *arg)   /* main.cpp:612-635 */
{
    feed_t *f = (feed_t *)arg;
    const double t_end = now_s() + f->seconds;
    while (!f->stop && now_s() < t_end) {
        double t0 = now_s();
        if (nbody_gpu_step(f->ctx, f->dt, 1) != NBODY_OK) { f->failed = 1; break; }
        double t1 = now_s();
        pthread_mutex_lock(&f->lock);
        double t2 = now_s();
        int rc = nbody_gpu_download(f->ctx, f->shared, f->n, f->fields);   /* synchronises: the step is complete here */
        f->frame++;
        pthread_mutex_unlock(&f->lock);
        double t3 = now_s();
        if (rc != NBODY_OK) { f->failed = 1; break; }
        f->steps++; f->publishes++;
        f->step_s += t1 - t0; f->publish_s += t3 - t2;
        if (t2 - t1 > f->sim_lock_wait_max) f->sim_lock_wait_max = t2 - t1;
        sleep_ms(f->sleep_ms);
    }
    f->stop = 1;
    return NULL;
}

static void *render_thread(void *arg)       /* the render loop's copy, 60 Hz */
{
    feed_t *f = (feed_t *)arg;
    nbody_body_t *local = (nbody_body_t *)malloc(f->n * sizeof *local);
    unsigned long long last = 0;
    if (!local) { f->failed = 1; return NULL; }
    while (!f->stop) {
        double t0 = now_s();
        pthread_mutex_lock(&f->lock);
        double t1 = now_s();
        memcpy(local, f->shared, f->n * sizeof *local);
        unsigned long long frame = f->frame;
        pthread_mutex_unlock(&f->lock);
        double t2 = now_s();
        if (t1 - t0 > f->view_lock_wait_max) f->view_lock_wait_max = t1 - t0;
        f->view_copy_s += t2 - t1;
        /* "draw": every position must be finite and the frame counter must not run backwards */
        int bad = frame < last;
        for (size_t i = 0; i < f->n && !bad; ++i) bad = !(local[i].pos[0] == local[i].pos[0] && local[i].pos[1] == local[i].pos[1]);
        f->torn += (unsigned long long)bad;
        last = frame;
        f->frames++;
        sleep_ms(1000.0 / 60.0);
    }
    free(local);
    return NULL;
}

int main(int argc, char **argv)
{
    feed_t f;
    memset(&f, 0, sizeof f);
    f.n = 25000; f.dt = 0.01f; f.seconds = 2.0; f.sleep_ms = 1.0; f.fields = NBODY_FIELD_POS | NBODY_FIELD_VEL;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--seconds")) f.seconds = atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--n")) f.n = (size_t)strtoull(argv[i + 1], NULL, 10);
        else if (!strcmp(argv[i], "--sleep-ms")) f.sleep_ms = atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--fields")) f.fields = !strcmp(argv[i + 1], "pos") ? NBODY_FIELD_POS : !strcmp(argv[i + 1], "all") ? NBODY_FIELD_ALL : (NBODY_FIELD_POS | NBODY_FIELD_VEL);
        else { fprintf(stderr, "usage: %s [--seconds S] [--n N] [--sleep-ms M] [--fields pos|posvel|all]\n", argv[0]); return 2; }
    }
    f.shared = (nbody_body_t *)calloc(f.n, sizeof *f.shared);
    if (!f.shared || nbody_ic_reference_disc(f.shared, f.n) != 0) { fprintf(stderr, "initial conditions failed\n"); return 1; }
    nbody_params p;
    nbody_params_default(&p);                 /* the reference's shipped configuration: Simulation::step() bit for bit */
    p.force_algo = NBODY_FORCE_BARNES_HUT;
    p.rsqrt_mode = NBODY_RSQRT_REFCOMPAT;
    p.integ_flags = NBODY_INTEG_CLAMP | NBODY_INTEG_BOUNDARY;
    p.collide = 1;
    int rc = nbody_gpu_init(&f.ctx, &p, f.shared, f.n);
    if (rc != NBODY_OK) { fprintf(stderr, "nbody_gpu_init: %s: %s\n", nbody_gpu_strerror(rc), nbody_gpu_last_error(NULL)); return 1; }
    pthread_mutex_init(&f.lock, NULL);
    pthread_t sim, view;
    const double t0 = now_s();
    pthread_create(&sim, NULL, simulation_thread, &f);
    pthread_create(&view, NULL, render_thread, &f);
    pthread_join(sim, NULL);
    pthread_join(view, NULL);
    const double wall = now_s() - t0;
    nbody_gpu_shutdown(f.ctx);
    const double ns = f.steps ? (double)f.steps : 1.0, nf = f.frames ? (double)f.frames : 1.0;
    printf("{\"n\": %zu, \"seconds\": %.3f, \"steps\": %llu, \"steps_per_s\": %.1f, \"render_frames\": %llu, \"frames_per_s\": %.1f, "
           "\"step_enqueue_ms\": %.4f, \"publish_ms\": %.4f, \"sim_lock_wait_max_ms\": %.4f, \"view_lock_wait_max_ms\": %.4f, "
           "\"view_copy_ms\": %.4f, \"bad_frames\": %llu, \"sleep_ms\": %.2f, \"fields\": %u, \"failed\": %d}\n",
           f.n, wall, f.steps, f.steps / wall, f.frames, f.frames / wall, 1e3 * f.step_s / ns, 1e3 * f.publish_s / ns,
           1e3 * f.sim_lock_wait_max, 1e3 * f.view_lock_wait_max, 1e3 * f.view_copy_s / nf, f.torn, f.sleep_ms, f.fields, f.failed);
    free(f.shared);
    return f.failed ? 1 : 0;
}
