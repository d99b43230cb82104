/* TEST INFRASTRUCTURE ONLY -- see nbody_oracle.h.  Build: -O2 -ffp-contract=off (no FMA
 * contraction, no fast-math) so every float expression rounds exactly as the strict build of
 * the reference does. */
#include "nbody_oracle.h"
#include <math.h>
#include <string.h>

/* Quadtree.hpp:106-111.  bit_cast -> memcpy; `number * 0.5f * y * y` is left-associative. */
float orc_fast_inv_sqrt(float number)
{
    uint32_t u;
    memcpy(&u, &number, 4);
    u = 0x5f3759dfu - (u >> 1);
    float y;
    memcpy(&y, &u, 4);
    return y * (1.5f - (((number * 0.5f) * y) * y));
}

/* One target through the leaf loop, Quadtree.hpp:134-144.
 *   r = body.pos - pos                (Vec2 operator-,  Vec2.hpp:88-96)
 *   r_sq = r.x*r.x + r.y*r.y          (Vec2::mag_sq,    Vec2.hpp:216-219)
 *   if (r_sq > 0) { inv = fast_inv_sqrt(r_sq + e_sq); inv3 = inv*inv*inv;
 *                   acc += r * (body.mass * inv3); }   (operator* then operator+=) */
static void acc_one(const orc_body_t *b, size_t n, float e_sq, int dims, size_t i, float *out)
{
    const float px = b[i].px, py = b[i].py, pz = b[i].pz;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    for (size_t j = 0; j < n; ++j) {
        float rx = b[j].px - px;
        float ry = b[j].py - py;
        float rz = (dims == 3) ? (b[j].pz - pz) : 0.0f;
        float r_sq = rx * rx + ry * ry;
        if (dims == 3) r_sq = r_sq + rz * rz;
        if (r_sq > 0) {
            float inv = orc_fast_inv_sqrt(r_sq + e_sq);
            float inv3 = inv * inv * inv;
            float s = b[j].mass * inv3;
            ax += rx * s;
            ay += ry * s;
            if (dims == 3) az += rz * s;
        }
    }
    out[0] = ax;
    out[1] = ay;
    if (dims == 3) out[2] = az;
}

void orc_allpairs_acc(const orc_body_t *b, size_t n, float eps, int dims, size_t i0, size_t i1,
                      float *acc_out)
{
    const float e_sq = eps * eps; /* Quadtree ctor, Quadtree.hpp:19 */
#pragma omp parallel for schedule(static)
    for (long long i = (long long)i0; i < (long long)i1; ++i)
        acc_one(b, n, e_sq, dims, (size_t)i, acc_out + (size_t)dims * ((size_t)i - i0));
}

void orc_allpairs_acc_f64sum(const orc_body_t *b, size_t n, float eps, int dims, size_t i0,
                             size_t i1, double *acc_out)
{
    const float e_sq = eps * eps;
#pragma omp parallel for schedule(static)
    for (long long ii = (long long)i0; ii < (long long)i1; ++ii) {
        size_t i = (size_t)ii;
        double ax = 0, ay = 0, az = 0;
        for (size_t j = 0; j < n; ++j) {
            float rx = b[j].px - b[i].px, ry = b[j].py - b[i].py;
            float rz = (dims == 3) ? (b[j].pz - b[i].pz) : 0.0f;
            float r_sq = rx * rx + ry * ry;
            if (dims == 3) r_sq = r_sq + rz * rz;
            if (r_sq > 0) {
                float inv = orc_fast_inv_sqrt(r_sq + e_sq);
                float s = b[j].mass * (inv * inv * inv);
                ax += (double)(rx * s);
                ay += (double)(ry * s);
                az += (double)(rz * s);
            }
        }
        double *o = acc_out + (size_t)dims * (i - i0);
        o[0] = ax;
        o[1] = ay;
        if (dims == 3) o[2] = az;
    }
}

/* Body.hpp:34-38.  `acc * dt` is Vec2::operator* (one rounding), `+=` adds (second rounding). */
void orc_body_update(orc_body_t *b, size_t n, float dt, int dims)
{
    for (size_t i = 0; i < n; ++i) {
        b[i].vx += b[i].ax * dt;
        b[i].vy += b[i].ay * dt;
        b[i].px += b[i].vx * dt;
        b[i].py += b[i].vy * dt;
        if (dims == 3) {
            b[i].vz += b[i].az * dt;
            b[i].pz += b[i].vz * dt;
        }
    }
}

void orc_step_clean(orc_body_t *b, size_t n, float eps, float dt, int nsteps, int dims)
{
    float *acc = (float *)__builtin_malloc(sizeof(float) * 3 * n);
    for (int s = 0; s < nsteps; ++s) {
        orc_allpairs_acc(b, n, eps, dims, 0, n, acc);
        for (size_t i = 0; i < n; ++i) {
            b[i].ax = acc[dims * i];
            b[i].ay = acc[dims * i + 1];
            if (dims == 3) b[i].az = acc[dims * i + 2];
        }
        orc_body_update(b, n, dt, dims);
    }
    __builtin_free(acc);
}

/* Simulation.hpp:116-163 minus attract().  Constants :120-124. */
void orc_iterate_after_attract(orc_body_t *b, size_t n, float dt, unsigned flags, int dims)
{
    const float BOUNDARY_RADIUS = 100000.0f;
    const float SOFT_BOUNDARY = BOUNDARY_RADIUS * 0.8f;
    const float BOUNDARY_FORCE = 0.9f;
    const float DAMPING = 0.9995f;
    const float MAX_VELOCITY = 1000.0f;
    for (size_t i = 0; i < n; ++i) { /* :129-138 */
        b[i].vx += b[i].ax * dt;
        b[i].vy += b[i].ay * dt;
        if (dims == 3) b[i].vz += b[i].az * dt;
        if (flags & 1u) {
            float v2 = b[i].vx * b[i].vx + b[i].vy * b[i].vy;
            if (dims == 3) v2 = v2 + b[i].vz * b[i].vz;
            if (v2 > MAX_VELOCITY * MAX_VELOCITY) {
                float scale = MAX_VELOCITY / sqrtf(v2);
                b[i].vx *= scale;
                b[i].vy *= scale;
                if (dims == 3) b[i].vz *= scale;
            }
        }
    }
    if (flags & 2u) { /* :140-155 */
        const float SOFT_SQ = SOFT_BOUNDARY * SOFT_BOUNDARY;
        for (size_t i = 0; i < n; ++i) {
            float d2 = b[i].px * b[i].px + b[i].py * b[i].py;
            if (dims == 3) d2 = d2 + b[i].pz * b[i].pz;
            if (d2 > SOFT_SQ) {
                float dist = sqrtf(d2);
                float ratio = dist / SOFT_BOUNDARY;
                float force = BOUNDARY_FORCE * expf(ratio - 1.0f);
                float k = -1.0f / dist;          /* dir = pos * (-1/dist) */
                float fd = force * dt;           /* boundaryForce = dir * (force*dt) */
                b[i].vx += (b[i].px * k) * fd;
                b[i].vy += (b[i].py * k) * fd;
                if (dims == 3) b[i].vz += (b[i].pz * k) * fd;
                b[i].vx *= DAMPING;
                b[i].vy *= DAMPING;
                if (dims == 3) b[i].vz *= DAMPING;
            }
        }
    }
    for (size_t i = 0; i < n; ++i) { /* :160-163 */
        b[i].px += b[i].vx * dt;
        b[i].py += b[i].vy * dt;
        if (dims == 3) b[i].pz += b[i].vz * dt;
    }
}

void orc_exact_acc_f64(const orc_body_t *b, size_t n, double eps, double G, int dims, size_t i0,
                       size_t i1, double *acc_out)
{
    const double e2 = eps * eps;
#pragma omp parallel for schedule(static)
    for (long long ii = (long long)i0; ii < (long long)i1; ++ii) {
        size_t i = (size_t)ii;
        double ax = 0, ay = 0, az = 0;
        const double px = b[i].px, py = b[i].py, pz = (dims == 3) ? b[i].pz : 0.0;
        for (size_t j = 0; j < n; ++j) {
            double rx = (double)b[j].px - px, ry = (double)b[j].py - py;
            double rz = (dims == 3) ? ((double)b[j].pz - pz) : 0.0;
            double r2 = rx * rx + ry * ry + rz * rz;
            if (r2 > 0) {
                double d2 = r2 + e2;
                double s = G * (double)b[j].mass / (d2 * sqrt(d2));
                ax += rx * s;
                ay += ry * s;
                az += rz * s;
            }
        }
        acc_out[3 * (i - i0) + 0] = ax;
        acc_out[3 * (i - i0) + 1] = ay;
        acc_out[3 * (i - i0) + 2] = az;
    }
}

void orc_energy_f64(const orc_body_t *b, size_t n, double eps, double G, int dims, double *K,
                    double *W, double P[3])
{
    const double e2 = eps * eps;
    double k = 0, w = 0, p0 = 0, p1 = 0, p2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : k, w, p0, p1, p2)
    for (long long ii = 0; ii < (long long)n; ++ii) {
        size_t i = (size_t)ii;
        const double m = b[i].mass;
        const double vx = b[i].vx, vy = b[i].vy, vz = (dims == 3) ? b[i].vz : 0.0;
        k += 0.5 * m * (vx * vx + vy * vy + vz * vz);
        p0 += m * vx;
        p1 += m * vy;
        p2 += m * vz;
        double wi = 0;
        for (size_t j = i + 1; j < n; ++j) {
            double rx = (double)b[j].px - b[i].px, ry = (double)b[j].py - b[i].py;
            double rz = (dims == 3) ? ((double)b[j].pz - b[i].pz) : 0.0;
            wi += (double)b[j].mass / sqrt(rx * rx + ry * ry + rz * rz + e2);
        }
        w -= G * m * wi;
    }
    *K = k;
    *W = w;
    P[0] = p0;
    P[1] = p1;
    P[2] = p2;
}
