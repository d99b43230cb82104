/* nbody_run -- headless C driver over the C ABI (include/nbody_gpu.h).
 *
 * The reference has no command line at all: `int main()` takes no argv (main.cpp:637) and every
 * parameter is a compile-time literal.  This driver's flags are new
 * surface whose DEFAULTS are those literals, so a NO-ARGUMENT RUN IS THE REFERENCE'S SIMULATION,
 * headless (other scenes, --ic plummer etc., default to the fast all-pairs path): uniform_disc(25000) (Simulation.hpp:61,347-603), Barnes-Hut theta=1, eps=1 (:59),
 * dt=0.01 (main.cpp:39), clamp + soft boundary (:120-155), collide() (:216-346), the reference's
 * own rsqrt -- i.e. Simulation::step() bit for bit.  It replaces simulation_thread's loop
 * (main.cpp:612-635): step, then hand the bodies to a consumer (here: diagnostics / snapshot).
 */
#define _POSIX_C_SOURCE 200809L
#include "nbody_gpu.h"
#include "nbody_host.h"
#include <math.h>
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

static void usage(const char *a0)
{
    fprintf(stderr,
            "usage: %s [--n N] [--steps K] [--dt DT] [--eps E] [--ic plummer|sphere|galaxy|disc|reference]\n"
            "          [--seed S] [--dims 2|3] [--gpus G] [--precision f32|f64]\n"
            "          [--rsqrt fast|refcompat] [--clamp on|off] [--boundary on|off]\n"
            "          [--algo allpairs|bh] [--theta T] [--near-leaves on|off] [--collide on|off] [--exchange auto|nccl]\n"
            "          [--energy-every M] [--in snapshot] [--out snapshot] [--splits S]\n",
            a0);
}

int main(int argc, char **argv)
{
    size_t n = 25000;
    int steps = 100, dims = 2, gpus = 1, energy_every = 0, splits = 0;
    float dt = 0.01f, eps = 1.0f;
    unsigned long long seed = 0;
    const char *ic = "reference", *in_path = NULL, *out_path = NULL;
    nbody_params p;
    nbody_params_default(&p);
    /* explicit choices (-1 = not given): with the default scene (--ic reference) the unset ones fall back to
     * the reference's shipped configuration, with any other scene to the library defaults */
    int o_algo = -1, o_rsqrt = -1, o_clamp = -1, o_boundary = -1, o_collide = -1;

    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
#define NEED() do { if (!v) { usage(argv[0]); return 2; } ++i; } while (0)
        if (!strcmp(a, "--n")) { NEED(); n = (size_t)strtoull(v, NULL, 10); }
        else if (!strcmp(a, "--steps")) { NEED(); steps = atoi(v); }
        else if (!strcmp(a, "--dt")) { NEED(); dt = (float)atof(v); }
        else if (!strcmp(a, "--eps")) { NEED(); eps = (float)atof(v); }
        else if (!strcmp(a, "--ic")) { NEED(); ic = v; }
        else if (!strcmp(a, "--seed")) { NEED(); seed = strtoull(v, NULL, 10); }
        else if (!strcmp(a, "--dims")) { NEED(); dims = atoi(v); }
        else if (!strcmp(a, "--gpus")) { NEED(); gpus = atoi(v); }
        else if (!strcmp(a, "--splits")) { NEED(); splits = atoi(v); }
        else if (!strcmp(a, "--precision")) { NEED(); p.precision = !strcmp(v, "f64") ? NBODY_PRECISION_F64 : NBODY_PRECISION_F32; }
        else if (!strcmp(a, "--rsqrt")) { NEED(); o_rsqrt = !strcmp(v, "refcompat"); }
        else if (!strcmp(a, "--clamp")) { NEED(); o_clamp = !strcmp(v, "on"); }
        else if (!strcmp(a, "--boundary")) { NEED(); o_boundary = !strcmp(v, "on"); }
        else if (!strcmp(a, "--algo")) { NEED(); o_algo = !strcmp(v, "bh"); }
        else if (!strcmp(a, "--theta")) { NEED(); p.theta = (float)atof(v); }
        else if (!strcmp(a, "--near-leaves")) { NEED(); p.bh_fix_near_leaves = !strcmp(v, "on"); }
        else if (!strcmp(a, "--collide")) { NEED(); o_collide = !strcmp(v, "on"); }
        else if (!strcmp(a, "--exchange")) { NEED(); p.exchange = !strcmp(v, "nccl") ? 1 : 0; }
        else if (!strcmp(a, "--energy-every")) { NEED(); energy_every = atoi(v); }
        else if (!strcmp(a, "--in")) { NEED(); in_path = v; }
        else if (!strcmp(a, "--out")) { NEED(); out_path = v; }
        else { usage(argv[0]); return 2; }
#undef NEED
    }

    nbody_snapshot_header_t hdr;
    memset(&hdr, 0, sizeof hdr);
    if (in_path) {
        if (nbody_snapshot_read_header(in_path, &hdr) != 0) { fprintf(stderr, "cannot read %s\n", in_path); return 1; }
        n = (size_t)hdr.n;
        dims = (int)hdr.dims;
    }
    nbody_body_t *b = (nbody_body_t *)calloc(n, sizeof *b);
    if (!b) { fprintf(stderr, "out of host memory\n"); return 1; }
    int rc = 0;
    if (in_path) rc = nbody_snapshot_read(in_path, &hdr, b, n);
    else if (!strcmp(ic, "plummer")) rc = nbody_ic_plummer(b, n, seed, dims);
    else if (!strcmp(ic, "sphere")) rc = nbody_ic_uniform_sphere(b, n, seed, dims, 0.5);
    else if (!strcmp(ic, "galaxy")) rc = nbody_ic_two_galaxy(b, n, seed, dims);
    else if (!strcmp(ic, "reference")) rc = nbody_ic_reference_disc(b, n);   /* the reference's own uniform_disc scene */
    else if (!strcmp(ic, "disc")) rc = nbody_ic_spinning_disc(b, n, seed, 100.0f * sqrtf((float)n / 1024.0f), 0.3f / sqrtf((float)n / 1024.0f), 1.0f);
    else { usage(argv[0]); return 2; }
    if (rc != 0) { fprintf(stderr, "initial conditions failed (%d)\n", rc); return 1; }

    {
        /* the reference's configuration (BH theta=1, its own rsqrt, clamp + boundary, collide) for its own scene */
        const int refcfg = !in_path && !strcmp(ic, "reference") && dims == 2 && gpus == 1 && p.precision == NBODY_PRECISION_F32;
        const int algo = o_algo >= 0 ? o_algo : refcfg, rsq = o_rsqrt >= 0 ? o_rsqrt : refcfg;
        const int clamp = o_clamp >= 0 ? o_clamp : refcfg, bound = o_boundary >= 0 ? o_boundary : refcfg;
        const int coll = o_collide >= 0 ? o_collide : refcfg;
        p.force_algo = algo ? NBODY_FORCE_BARNES_HUT : NBODY_FORCE_ALLPAIRS;
        p.rsqrt_mode = rsq ? NBODY_RSQRT_REFCOMPAT : NBODY_RSQRT_FAST;
        p.integ_flags = (clamp ? NBODY_INTEG_CLAMP : 0u) | (bound ? NBODY_INTEG_BOUNDARY : 0u);
        p.collide = coll;
    }
    p.dims = dims;
    p.eps = eps;
    p.j_splits = splits;
    p.ngpus = gpus;
    for (int g = 0; g < gpus && g < NBODY_MAX_GPUS; ++g) p.device_ids[g] = g;

    nbody_ctx *ctx = NULL;
    rc = nbody_gpu_init(&ctx, &p, b, n);
    if (rc != NBODY_OK) {
        fprintf(stderr, "nbody_gpu_init: %s: %s\n", nbody_gpu_strerror(rc), nbody_gpu_last_error(NULL));
        return 1;
    }
    nbody_info info;
    nbody_gpu_get_info(ctx, &info);
    printf("%s\nn=%zu (padded %llu) dims=%d eps=%g dt=%g gpus=%d sms=%d splits=%d ctas=%d fused=%d exchange=%s\n",
           nbody_gpu_version(), n, (unsigned long long)info.n_padded, dims, eps, dt, gpus, info.sm_count,
           info.j_splits, info.force_ctas, info.fused, gpus > 1 ? (info.p2p_exchange ? "p2p-push" : "nccl") : "none");

    double K0 = 0, W0 = 0, P[3];
    if (energy_every > 0) {
        nbody_gpu_energy(ctx, &K0, &W0, P);
        printf("step %6d  K=%.9e  W=%.9e  E=%.9e\n", 0, K0, W0, K0 + W0);
    }
    const double t0 = now_s();
    int done = 0;
    while (done < steps) {
        int chunk = energy_every > 0 ? energy_every : steps;
        if (chunk > steps - done) chunk = steps - done;
        rc = nbody_gpu_step(ctx, dt, chunk);
        if (rc != NBODY_OK) { fprintf(stderr, "step: %s: %s\n", nbody_gpu_strerror(rc), nbody_gpu_last_error(ctx)); return 1; }
        done += chunk;
        if (energy_every > 0) {
            double K, W;
            nbody_gpu_energy(ctx, &K, &W, P);
            printf("step %6d  K=%.9e  W=%.9e  E=%.9e  dE/E0=%.3e  |P|=%.3e\n", done, K, W, K + W,
                   (K + W - K0 - W0) / fabs(K0 + W0), sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]));
        }
    }
    nbody_gpu_sync(ctx);
    const double t1 = now_s();
    nbody_gpu_get_info(ctx, &info);
    if (p.force_algo == NBODY_FORCE_BARNES_HUT)
        printf("%d steps in %.3f s  (%.3f ms/step incl. diagnostics; Barnes-Hut theta=%g%s)\n", steps, t1 - t0,
               1e3 * (t1 - t0) / (steps > 0 ? steps : 1), (double)p.theta, p.collide ? " + collision pass" : "");
    else
        printf("%d steps in %.3f s  (%.3f ms/step, %.2f G pair-interactions/s incl. diagnostics)\n", steps,
               t1 - t0, 1e3 * (t1 - t0) / (steps > 0 ? steps : 1), 1e-9 * (double)n * (double)n * steps / (t1 - t0));

    rc = nbody_gpu_download(ctx, b, n, NBODY_FIELD_ALL);
    if (rc != NBODY_OK) { fprintf(stderr, "download: %s\n", nbody_gpu_last_error(ctx)); return 1; }
    if (out_path) {
        memset(&hdr, 0, sizeof hdr);
        hdr.dims = (uint32_t)dims; hdr.n = n; hdr.step = (uint64_t)steps; hdr.time = (double)dt * steps;
        hdr.eps = eps; hdr.dt = dt;
        if (nbody_snapshot_write(out_path, &hdr, b) != 0) { fprintf(stderr, "cannot write %s\n", out_path); return 1; }
    }
    nbody_gpu_shutdown(ctx);
    free(b);
    return 0;
}
