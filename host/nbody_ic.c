/* Seeded initial conditions, snapshot I/O and the shard plan -- plain C host code (no CUDA).
 * See include/nbody_host.h.  New surface: the reference has no reusable generator, no file I/O
 * and no multi-device code (SURVEY.md sections 2 and 5). */
#include "nbody_host.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ RNG */
static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

void nbody_rng_seed(nbody_rng_t *r, uint64_t seed)
{
    uint64_t x = seed;
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}
static uint64_t rng_next(nbody_rng_t *r)
{
    uint64_t *s = r->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}
double nbody_rng_uniform(nbody_rng_t *r) { return (double)(rng_next(r) >> 11) * 0x1.0p-53; }
double nbody_rng_normal(nbody_rng_t *r)
{
    double u1 = nbody_rng_uniform(r), u2 = nbody_rng_uniform(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* ------------------------------------------------------------------ helpers */
static void zero_body(nbody_body_t *b) { memset(b, 0, sizeof *b); }

static void unit_vector(nbody_rng_t *r, int dims, double v[3])
{
    if (dims == 2) {
        double a = 2.0 * M_PI * nbody_rng_uniform(r);
        v[0] = cos(a); v[1] = sin(a); v[2] = 0.0;
    } else {
        double z = 2.0 * nbody_rng_uniform(r) - 1.0, a = 2.0 * M_PI * nbody_rng_uniform(r);
        double s = sqrt(fmax(0.0, 1.0 - z * z));
        v[0] = s * cos(a); v[1] = s * sin(a); v[2] = z;
    }
}

/* shift to centre-of-mass frame (positions and velocities) */
static void recentre(nbody_body_t *b, size_t n)
{
    double M = 0, c[3] = {0, 0, 0}, p[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        double m = b[i].mass;
        M += m;
        c[0] += m * b[i].pos[0]; c[1] += m * b[i].pos[1]; c[2] += m * b[i].pos_z;
        p[0] += m * b[i].vel[0]; p[1] += m * b[i].vel[1]; p[2] += m * b[i].vel_z;
    }
    if (M <= 0) return;
    for (size_t i = 0; i < n; ++i) {
        b[i].pos[0] -= (float)(c[0] / M); b[i].pos[1] -= (float)(c[1] / M);
        b[i].pos_z -= (float)(c[2] / M);
        b[i].vel[0] -= (float)(p[0] / M); b[i].vel[1] -= (float)(p[1] / M);
        b[i].vel_z -= (float)(p[2] / M);
    }
}

/* ------------------------------------------------------------------ generators */
int nbody_ic_uniform_sphere(nbody_body_t *b, size_t n, uint64_t seed, int dims, double virial)
{
    if (!b || n == 0 || (dims != 2 && dims != 3) || virial < 0) return NBODY_HOST_EINVAL;
    nbody_rng_t r;
    nbody_rng_seed(&r, seed);
    const double m = 1.0 / (double)n;
    for (size_t i = 0; i < n; ++i) {
        zero_body(&b[i]);
        double x, y, z;
        do { /* rejection sampling inside the unit ball / disc */
            x = 2.0 * nbody_rng_uniform(&r) - 1.0;
            y = 2.0 * nbody_rng_uniform(&r) - 1.0;
            z = (dims == 3) ? 2.0 * nbody_rng_uniform(&r) - 1.0 : 0.0;
        } while (x * x + y * y + z * z > 1.0);
        b[i].pos[0] = (float)x; b[i].pos[1] = (float)y; b[i].pos_z = (float)z;
        b[i].mass = (float)m;
    }
    if (virial > 0) {
        /* |W| of a uniform ball of mass 1 radius 1 is 3/5 (G=1); disc: 8/(3 pi).  Target
         * K = virial*|W|/2 = n * m * sigma^2 * dims / 2  ->  sigma^2 = virial*|W|/dims. */
        const double W = (dims == 3) ? 0.6 : 8.0 / (3.0 * M_PI);
        const double sigma = sqrt(virial * W / (double)dims);
        for (size_t i = 0; i < n; ++i) {
            b[i].vel[0] = (float)(sigma * nbody_rng_normal(&r));
            b[i].vel[1] = (float)(sigma * nbody_rng_normal(&r));
            b[i].vel_z = (dims == 3) ? (float)(sigma * nbody_rng_normal(&r)) : 0.0f;
        }
    }
    recentre(b, n);
    if (dims == 2) for (size_t i = 0; i < n; ++i) b[i].pos_z = b[i].vel_z = 0.0f;
    return 0;
}

static void plummer_fill(nbody_body_t *b, size_t n, nbody_rng_t *r, int dims, double mass_each)
{
    const double a = 3.0 * M_PI / 16.0; /* virial radius 1 in N-body units */
    for (size_t i = 0; i < n; ++i) {
        zero_body(&b[i]);
        double rad;
        do { /* cumulative mass fraction -> radius; reject beyond 10 a */
            double X = nbody_rng_uniform(r);
            if (X < 1e-12) X = 1e-12;
            rad = a / sqrt(pow(X, -2.0 / 3.0) - 1.0);
        } while (rad > 10.0 * a);
        double u[3];
        unit_vector(r, 3, u);
        /* speed: q = v/v_esc by von Neumann rejection on g(q) = q^2 (1-q^2)^(7/2) */
        double q, g;
        do {
            q = nbody_rng_uniform(r);
            g = 0.1 * nbody_rng_uniform(r);
        } while (g > q * q * pow(1.0 - q * q, 3.5));
        const double vesc = sqrt(2.0) * pow(rad * rad + a * a, -0.25);
        double w[3];
        unit_vector(r, 3, w);
        b[i].pos[0] = (float)(rad * u[0]); b[i].pos[1] = (float)(rad * u[1]);
        b[i].vel[0] = (float)(q * vesc * w[0]); b[i].vel[1] = (float)(q * vesc * w[1]);
        if (dims == 3) {
            b[i].pos_z = (float)(rad * u[2]);
            b[i].vel_z = (float)(q * vesc * w[2]);
        }
        b[i].mass = (float)mass_each;
    }
}

int nbody_ic_plummer(nbody_body_t *b, size_t n, uint64_t seed, int dims)
{
    if (!b || n == 0 || (dims != 2 && dims != 3)) return NBODY_HOST_EINVAL;
    nbody_rng_t r;
    nbody_rng_seed(&r, seed);
    plummer_fill(b, n, &r, dims, 1.0 / (double)n);
    recentre(b, n);
    if (dims == 2) for (size_t i = 0; i < n; ++i) b[i].pos_z = b[i].vel_z = 0.0f;
    return 0;
}

int nbody_ic_two_galaxy(nbody_body_t *b, size_t n, uint64_t seed, int dims)
{
    if (!b || n < 2 || (dims != 2 && dims != 3)) return NBODY_HOST_EINVAL;
    nbody_rng_t r;
    nbody_rng_seed(&r, seed);
    const size_t h = n / 2;
    plummer_fill(b, h, &r, dims, 0.5 / (double)h);
    recentre(b, h);
    plummer_fill(b + h, n - h, &r, dims, 0.5 / (double)(n - h));
    recentre(b + h, n - h);
    for (size_t i = 0; i < n; ++i) {
        const float sgn = (i < h) ? 1.0f : -1.0f;
        b[i].pos[0] += sgn * 5.0f;
        b[i].pos[1] += sgn * 1.0f;
        b[i].vel[0] -= sgn * 0.5f;
        if (dims == 2) b[i].pos_z = b[i].vel_z = 0.0f;
    }
    return 0;
}

int nbody_ic_spinning_disc(nbody_body_t *b, size_t n, uint64_t seed, float scale, float spin,
                           float m)
{
    if (!b || n == 0) return NBODY_HOST_EINVAL;
    nbody_rng_t r;
    nbody_rng_seed(&r, seed);
    for (size_t i = 0; i < n; ++i) {
        zero_body(&b[i]);
        double x, y;
        do {
            x = 2.0 * nbody_rng_uniform(&r) - 1.0;
            y = 2.0 * nbody_rng_uniform(&r) - 1.0;
        } while (x * x + y * y > 1.0);
        b[i].pos[0] = scale * (float)x; b[i].pos[1] = scale * (float)y;
        b[i].vel[0] = spin * (float)(-y) * scale; b[i].vel[1] = spin * (float)x * scale;
        b[i].mass = m;
    }
    return 0;
}

/* ---- Simulation::uniform_disc restated (Simulation.hpp:347-603) ------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t *g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(mt19937_t *g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* std::uniform_real_distribution<float>(0,1) over mt19937 as libstdc++ computes it:
 * generate_canonical<float,24>: one draw, float(draw) / 2^32, clamped below 1. */
static float mt_canonical(mt19937_t *g)
{
    float r = (float)mt_next(g) / 4294967296.0f;
    if (r >= 1.0f) r = 0.99999994f; /* nextafter(1.0f, 0.0f) */
    return r;
}
/* std::sort is unstable and the scene has a handful of exact ties in |pos|^2 (7 pairs at n=25000), whose
 * final order then depends on the library's algorithm.  To stay bit-identical with the libstdc++ build
 * of the reference, the sort below follows libstdc++'s std::sort structure: introsort (median-of-three
 * to front, unguarded partition, recursion on the right part) down to runs of 16, then one insertion
 * pass.  The heap-sort fallback of the depth limit (2*log2 n levels) is a plain heap sort here; it is
 * never reached for this data. */
static int less_mag_sq(const nbody_body_t *a, const nbody_body_t *b)
{
    return (a->pos[0] * a->pos[0] + a->pos[1] * a->pos[1]) < (b->pos[0] * b->pos[0] + b->pos[1] * b->pos[1]);
}
static void body_swap(nbody_body_t *a, nbody_body_t *b) { nbody_body_t t = *a; *a = *b; *b = t; }
static void unguarded_linear_insert(nbody_body_t *last)
{
    nbody_body_t val = *last;
    nbody_body_t *next = last - 1;
    while (less_mag_sq(&val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort(nbody_body_t *first, nbody_body_t *last)
{
    if (first == last) return;
    for (nbody_body_t *i = first + 1; i != last; ++i) {
        if (less_mag_sq(i, first)) {
            nbody_body_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof *first);
            *first = val;
        } else {
            unguarded_linear_insert(i);
        }
    }
}
static void heap_sort_fallback(nbody_body_t *first, nbody_body_t *last)
{
    const ptrdiff_t n = last - first;
    for (ptrdiff_t start = n / 2 - 1; start >= 0; --start)
        for (ptrdiff_t r = start;;) {
            ptrdiff_t c = 2 * r + 1;
            if (c >= n) break;
            if (c + 1 < n && less_mag_sq(first + c, first + c + 1)) ++c;
            if (!less_mag_sq(first + r, first + c)) break;
            body_swap(first + r, first + c); r = c;
        }
    for (ptrdiff_t end = n - 1; end > 0; --end) {
        body_swap(first, first + end);
        for (ptrdiff_t r = 0;;) {
            ptrdiff_t c = 2 * r + 1;
            if (c >= end) break;
            if (c + 1 < end && less_mag_sq(first + c, first + c + 1)) ++c;
            if (!less_mag_sq(first + r, first + c)) break;
            body_swap(first + r, first + c); r = c;
        }
    }
}
static void introsort_loop(nbody_body_t *first, nbody_body_t *last, int depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) { heap_sort_fallback(first, last); return; }
        --depth_limit;
        /* median of (first+1, mid, last-1) moved to *first */
        nbody_body_t *a = first + 1, *b = first + (last - first) / 2, *c = last - 1;
        if (less_mag_sq(a, b)) {
            if (less_mag_sq(b, c)) body_swap(first, b);
            else if (less_mag_sq(a, c)) body_swap(first, c);
            else body_swap(first, a);
        } else if (less_mag_sq(a, c)) body_swap(first, a);
        else if (less_mag_sq(b, c)) body_swap(first, c);
        else body_swap(first, b);
        /* unguarded partition of [first+1, last) around the pivot *first */
        nbody_body_t *lo = first + 1, *hi = last;
        for (;;) {
            while (less_mag_sq(lo, first)) ++lo;
            --hi;
            while (less_mag_sq(first, hi)) --hi;
            if (!(lo < hi)) break;
            body_swap(lo, hi);
            ++lo;
        }
        introsort_loop(lo, last, depth_limit);
        last = lo;
    }
}
static void sort_like_libstdcxx(nbody_body_t *first, size_t n)
{
    if (n == 0) return;
    nbody_body_t *last = first + n;
    int lg = 0;
    for (size_t t = n; t > 1; t >>= 1) ++lg;
    introsort_loop(first, last, 2 * lg);
    if (n > 16) {
        insertion_sort(first, first + 16);
        for (nbody_body_t *i = first + 16; i != last; ++i) unguarded_linear_insert(i);
    } else {
        insertion_sort(first, last);
    }
}

int nbody_ic_reference_disc(nbody_body_t *b, size_t n)
{
    if (!b || n == 0) return NBODY_HOST_EINVAL;
    mt19937_t rng;
    mt_seed(&rng, 0); /* std::mt19937 rng(0), :349 */
    const float inner_radius = 200.0f;
    const float outer_radius = sqrtf((float)n) * 300.7f;
    zero_body(&b[0]);
    b[0].mass = 1e9f;
    b[0].radius = inner_radius;
    /* mass buckets :372-397: probabilities normalised, then cumulative, all in float */
    const float mn[3] = {0.00005f, 1.2f, 5.0f}, mx[3] = {0.8f, 2.5f, 50.0f};
    float pr[3] = {0.825f, 0.125f, 0.025f};
    float total = 0.0f;
    for (int i = 0; i < 3; ++i) total += pr[i];
    for (int i = 0; i < 3; ++i) pr[i] /= total;
    float cum[3], c = 0.0f;
    for (int i = 0; i < 3; ++i) { c += pr[i]; cum[i] = c; }
    /* Lorenz attractor :399-405, :523-535 */
    const float sigma = 10.0f, rho = 28.0f, beta = 8.0f / 3.0f;
    float x = 0.1f, y = 0.0f, z = 0.0f;
    for (size_t k = 1; k < n; ++k) {
        const float dt = 0.01f;
        const float dx = sigma * (y - x);
        const float dy = x * (rho - z) - y;
        const float dz = x * y - beta * z;
        x += dx * dt;
        y += dy * dt;
        z += dz * dt;
        const float scale = outer_radius / 10.0f;
        const float px = x * scale, py = y * scale;
        float vx = -py, vy = px;
        const float m = sqrtf(vx * vx + vy * vy); /* Vec2::normalize: x is divided TWICE (Vec2.hpp:231-232) */
        if (m > 0.0f) { vx /= m; vx /= m; vy /= m; }
        const float rp = mt_canonical(&rng);
        size_t sel = 0;
        for (size_t i = 0; i < 3; ++i)
            if (rp <= cum[i]) { sel = i; break; }
        const float mass = mt_canonical(&rng) * (mx[sel] - mn[sel]) + mn[sel];
        zero_body(&b[k]);
        b[k].pos[0] = px; b[k].pos[1] = py;
        b[k].vel[0] = vx; b[k].vel[1] = vy;
        b[k].mass = mass;
        b[k].radius = cbrtf(mass);
    }
    sort_like_libstdcxx(b, n); /* std::sort by |pos|^2, :585-590 */
    float total_mass = 0.0f;
    for (size_t i = 0; i < n; ++i) { /* :592-600 */
        total_mass += b[i].mass;
        if (b[i].pos[0] == 0.0f && b[i].pos[1] == 0.0f) continue;
        const float r = sqrtf(b[i].pos[0] * b[i].pos[0] + b[i].pos[1] * b[i].pos[1]);
        const float v = sqrtf(total_mass / r);
        b[i].vel[0] *= v;
        b[i].vel[1] *= v;
    }
    return 0;
}

void nbody_ic_rescale(nbody_body_t *b, size_t n, float ls, float vs, float ms)
{
    for (size_t i = 0; i < n; ++i) {
        b[i].pos[0] *= ls; b[i].pos[1] *= ls; b[i].pos_z *= ls;
        b[i].vel[0] *= vs; b[i].vel[1] *= vs; b[i].vel_z *= vs;
        b[i].mass *= ms;
    }
}

/* ------------------------------------------------------------------ snapshots */
static const char SNAP_MAGIC[8] = {'N', 'B', 'O', 'D', 'Y', 'B', '2', '\0'};

int nbody_snapshot_write(const char *path, const nbody_snapshot_header_t *h, const nbody_body_t *b)
{
    if (!path || !h || !b) return NBODY_HOST_EINVAL;
    FILE *f = fopen(path, "wb");
    if (!f) return NBODY_HOST_EIO;
    nbody_snapshot_header_t hh = *h;
    memcpy(hh.magic, SNAP_MAGIC, 8);
    hh.version = 1;
    int ok = fwrite(&hh, sizeof hh, 1, f) == 1 && fwrite(b, sizeof *b, hh.n, f) == hh.n;
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : NBODY_HOST_EIO;
}

int nbody_snapshot_read_header(const char *path, nbody_snapshot_header_t *h)
{
    if (!path || !h) return NBODY_HOST_EINVAL;
    FILE *f = fopen(path, "rb");
    if (!f) return NBODY_HOST_EIO;
    int ok = fread(h, sizeof *h, 1, f) == 1 && memcmp(h->magic, SNAP_MAGIC, 8) == 0 &&
             h->version == 1;
    fclose(f);
    return ok ? 0 : NBODY_HOST_EIO;
}

int nbody_snapshot_read(const char *path, nbody_snapshot_header_t *h, nbody_body_t *b, size_t cap)
{
    if (!path || !h || !b) return NBODY_HOST_EINVAL;
    FILE *f = fopen(path, "rb");
    if (!f) return NBODY_HOST_EIO;
    int ok = fread(h, sizeof *h, 1, f) == 1 && memcmp(h->magic, SNAP_MAGIC, 8) == 0 &&
             h->version == 1 && h->n <= cap && fread(b, sizeof *b, h->n, f) == h->n;
    fclose(f);
    return ok ? 0 : NBODY_HOST_EIO;
}

/* ------------------------------------------------------------------ shard plan */
int nbody_shard_plan(size_t n, int world, int rank, size_t granule, size_t *n_padded,
                     size_t *start, size_t *count)
{
    if (world < 1 || rank < 0 || rank >= world || granule == 0) return NBODY_HOST_EINVAL;
    const size_t per = granule * (size_t)world;
    const size_t np = ((n + per - 1) / per) * per;
    const size_t cnt = (np ? np : per) / (size_t)world;
    if (n_padded) *n_padded = np ? np : per;
    if (start) *start = cnt * (size_t)rank;
    if (count) *count = cnt;
    return 0;
}
