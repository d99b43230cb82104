/* nbody_host.h -- host-side (plain C, no CUDA) helpers that sit either side of the hot path:
 * seeded initial-condition generators, the binary snapshot format and the shard plan used by the
 * multi-GPU driver.  None of this exists in the reference as reusable code: its only generator
 * is Simulation::uniform_disc (Simulation.hpp:347-603, a Lorenz-attractor trace, restated by
 * nbody_ic_reference_disc below), it has no file I/O and no multi-device code.
 */
#ifndef NBODY_HOST_H
#define NBODY_HOST_H
#include "nbody_body.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- deterministic RNG shared by all generators (splitmix64 seeding + xoshiro256**) ---- */
typedef struct { uint64_t s[4]; } nbody_rng_t;
void   nbody_rng_seed(nbody_rng_t *r, uint64_t seed);
double nbody_rng_uniform(nbody_rng_t *r);            /* [0,1) with 53 random bits */
double nbody_rng_normal(nbody_rng_t *r);             /* Box-Muller, one value per call */

/* ---- initial conditions (N-body units: G=1, total mass 1, radius=0 so collide() is inert) ----
 * dims=3: full 3-D.  dims=2: planar variant (z = vz = 0) for direct comparison with the 2-D
 * reference.  All return 0 or a negative NBODY_HOST_E* code. */
#define NBODY_HOST_EINVAL (-1)
#define NBODY_HOST_EIO    (-2)

/* Uniform ball (disc if dims=2) of radius 1; velocities isotropic Gaussian scaled so that
 * 2K/|W| = virial (0 -> cold start).  BASELINE.json configs[0]. */
int nbody_ic_uniform_sphere(nbody_body_t *b, size_t n, uint64_t seed, int dims, double virial);
/* Plummer model, Aarseth-Henon-Wielen sampling, scale radius a = 3*pi/16, cut at 10a,
 * equal masses 1/n.  BASELINE.json configs[1..3]. */
int nbody_ic_plummer(nbody_body_t *b, size_t n, uint64_t seed, int dims);
/* Two Plummer spheres of n/2 bodies, centres (+-5,+-1,0), approach velocity (-+0.5,0,0).
 * BASELINE.json configs[4]. */
int nbody_ic_two_galaxy(nbody_body_t *b, size_t n, uint64_t seed, int dims);
/* SURVEY.md section 4 KAT disc: unit-disc samples (x,y); pos = scale*(x,y), vel = spin*(-y,x),
 * mass = m each (reference-unit scale: scale=100, spin=0.3, m=1, eps=1). */
int nbody_ic_spinning_disc(nbody_body_t *b, size_t n, uint64_t seed, float scale, float spin,
                           float m);
/* The reference's own scene, Simulation::uniform_disc (Simulation.hpp:347-603), restated: body 0 is a
 * 1e9 central mass of radius 200; the others follow an Euler-integrated Lorenz attractor scaled by
 * sqrt(n)*300.7/10, masses from a three-bucket distribution drawn with std::mt19937(0), radius =
 * cbrt(mass), sorted by |pos|, tangential speed sqrt(M_enclosed/r) applied to the reference's
 * (double-divide) "normalised" direction (Vec2::normalize, Vec2.hpp:226-236).  Bit-identical to the
 * strict libstdc++ build of the reference (tests/test_reference_scene.py). */
int nbody_ic_reference_disc(nbody_body_t *b, size_t n);
/* Scale lengths/velocities/masses in place (reference-unit runs with eps=1). */
void nbody_ic_rescale(nbody_body_t *b, size_t n, float lscale, float vscale, float mscale);

/* ---- snapshot format: 64-byte header + n raw 64-byte Body records (little endian) ---- */
typedef struct {
    char     magic[8];      /* "NBODYB2\0" */
    uint32_t version;       /* 1 */
    uint32_t dims;          /* 2 or 3 */
    uint64_t n;
    uint64_t step;
    double   time;
    float    eps;
    float    dt;
    uint8_t  reserved[16];
} nbody_snapshot_header_t;
int nbody_snapshot_write(const char *path, const nbody_snapshot_header_t *h, const nbody_body_t *b);
int nbody_snapshot_read_header(const char *path, nbody_snapshot_header_t *h);
int nbody_snapshot_read(const char *path, nbody_snapshot_header_t *h, nbody_body_t *b, size_t cap);

/* ---- shard plan of the multi-GPU driver: targets [start, start+count) belong to `rank`.
 * n_padded is n rounded up so that every rank owns the same whole number of `granule`-body
 * tiles; bodies >= n are zero-mass padding that contribute exactly 0 to every sum. ---- */
int nbody_shard_plan(size_t n, int world, int rank, size_t granule, size_t *n_padded,
                     size_t *start, size_t *count);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_HOST_H */
