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

/* ============================================================================ Barnes-Hut path */
#include <stdlib.h>

typedef struct { orc_node_t *v; size_t n, cap; } nodevec_t;
typedef struct { size_t *v; size_t n, cap; } idxvec_t;

static void nv_push(nodevec_t *a, orc_node_t x)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (orc_node_t *)realloc(a->v, a->cap * sizeof *a->v);
    }
    a->v[a->n++] = x;
}
static void iv_push(idxvec_t *a, size_t x)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (size_t *)realloc(a->v, a->cap * sizeof *a->v);
    }
    a->v[a->n++] = x;
}

/* Node(next, quad, depth): data{Vec2::zero(), 0.0f, quad}, children 0.   Node.hpp:47-48 */
static orc_node_t make_node(uint64_t next, float cx, float cy, float size, uint64_t depth)
{
    orc_node_t nd;
    nd.px = 0.0f; nd.py = 0.0f; nd.mass = 0.0f;
    nd.cx = cx; nd.cy = cy; nd.size = size;
    nd.children = 0; nd.next = next; nd.depth = depth;
    return nd;
}

/* Quad::find_quadrant, Quad.hpp:47-49 */
static size_t find_quadrant(const orc_node_t *nd, float x, float y)
{
    return ((size_t)(y > nd->cy) << 1) | (size_t)(x > nd->cx);
}

/* Quad::into_quadrant, Quad.hpp:51-57: new_size = size*0.5f; offset = unit_x*((q&1)-0.5f) +
 * unit_y*((q>>1)-0.5f); new_center = center + offset*new_size. */
static void into_quadrant(const orc_node_t *nd, size_t q, float *cx, float *cy, float *size)
{
    const float new_size = nd->size * 0.5f;
    const float fx = (float)(q & 1) - 0.5f, fy = (float)(q >> 1) - 0.5f;
    const float ox = 1.0f * fx + 0.0f * fy; /* unit_x()*fx + unit_y()*fy, component-wise */
    const float oy = 0.0f * fx + 1.0f * fy;
    *cx = nd->cx + ox * new_size;
    *cy = nd->cy + oy * new_size;
    *size = new_size;
}

/* Quadtree::insert, Quadtree.hpp:35-93 */
static void bh_insert(nodevec_t *nodes, idxvec_t *parents, float x, float y, float mass)
{
    size_t node = 0;
    while (nodes->v[node].children != 0) {
        size_t q = find_quadrant(&nodes->v[node], x, y);
        node = nodes->v[node].children + q;
    }
    if (nodes->v[node].mass == 0.0f) { /* is_empty */
        nodes->v[node].px = x; nodes->v[node].py = y; nodes->v[node].mass = mass;
        return;
    }
    const float ex = nodes->v[node].px, ey = nodes->v[node].py, em = nodes->v[node].mass;
    if (x == ex && y == ey) { /* coincident: merge, :56-60 */
        nodes->v[node].mass += mass;
        return;
    }
    for (;;) {
        const size_t children = nodes->n;
        nodes->v[node].children = children;
        iv_push(parents, node);
        for (size_t i = 0; i < 4; ++i) {
            float cx, cy, sz;
            const orc_node_t parent = nodes->v[node]; /* copy: the vector may reallocate */
            into_quadrant(&parent, i, &cx, &cy, &sz);
            nv_push(nodes, make_node((i < 3) ? children + i + 1 : parent.next, cx, cy, sz, parent.depth + 1));
        }
        const size_t q1 = find_quadrant(&nodes->v[node], ex, ey);
        const size_t q2 = find_quadrant(&nodes->v[node], x, y);
        if (q1 == q2) {
            node = children + q1;
        } else {
            nodes->v[children + q1].px = ex; nodes->v[children + q1].py = ey; nodes->v[children + q1].mass = em;
            nodes->v[children + q2].px = x; nodes->v[children + q2].py = y; nodes->v[children + q2].mass = mass;
            return;
        }
    }
}

size_t orc_bh_build(const orc_body_t *b, size_t n, orc_node_t **nodes_out)
{
    nodevec_t nodes = {0, 0, 0};
    idxvec_t parents = {0, 0, 0};
    /* Quad::new_containing, Quad.hpp:31-45 */
    float minx = 3.402823466e+38f, miny = 3.402823466e+38f, maxx = -3.402823466e+38f, maxy = -3.402823466e+38f;
    for (size_t i = 0; i < n; ++i) {
        minx = b[i].px < minx ? b[i].px : minx; /* std::min(a,b) = (b<a)?b:a */
        miny = b[i].py < miny ? b[i].py : miny;
        maxx = maxx < b[i].px ? b[i].px : maxx; /* std::max(a,b) = (a<b)?b:a */
        maxy = maxy < b[i].py ? b[i].py : maxy;
    }
    const float cx = (minx + maxx) * 0.5f, cy = (miny + maxy) * 0.5f;
    const float sx = maxx - minx, sy = maxy - miny;
    const float size = sx < sy ? sy : sx; /* std::max(x, y) */
    nv_push(&nodes, make_node(0, cx, cy, size, 0)); /* clear(): Node(0, quad) */
    for (size_t i = 0; i < n; ++i) bh_insert(&nodes, &parents, b[i].px, b[i].py, b[i].mass);
    /* propagate, Quadtree.hpp:236-258 */
    for (size_t k = parents.n; k-- > 0;) {
        const size_t node = parents.v[k], child = nodes.v[node].children;
        float px = 0.0f, py = 0.0f, m = 0.0f;
        for (size_t i = 0; i < 4; ++i) {
            const orc_node_t *c = &nodes.v[child + i];
            px += c->px * c->mass;
            py += c->py * c->mass;
            m += c->mass;
        }
        if (m > 0) { /* Vec2::operator/=: inv = 1/scalar; x *= inv; y *= inv */
            const float inv = 1.0f / m;
            px *= inv;
            py *= inv;
        }
        nodes.v[node].px = px; nodes.v[node].py = py; nodes.v[node].mass = m;
    }
    free(parents.v);
    *nodes_out = nodes.v;
    return nodes.n;
}

/* Quadtree::acc, Quadtree.hpp:113-155 */
void orc_bh_acc(const orc_node_t *nodes, size_t nnodes, float theta, float eps, const orc_body_t *b,
                size_t i0, size_t i1, int fix_near_leaves, float *acc_out)
{
    (void)nnodes;
    const float t_sq = theta * theta, e_sq = eps * eps;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long ii = (long long)i0; ii < (long long)i1; ++ii) {
        const float px = b[ii].px, py = b[ii].py;
        float ax = 0.0f, ay = 0.0f;
        size_t node = 0;
        for (;;) {
            const orc_node_t *n = &nodes[node];
            const float dx = n->px - px, dy = n->py - py;
            const float d_sq = dx * dx + dy * dy;
            if (n->size * n->size < d_sq * t_sq) {
                if (d_sq > 0) {
                    const float inv = orc_fast_inv_sqrt(d_sq + e_sq);
                    const float inv3 = inv * inv * inv;
                    const float s = n->mass * inv3;
                    ax += dx * s;
                    ay += dy * s;
                }
                if (n->next == 0) break;
                node = n->next;
            } else if (n->children == 0) {
                /* near leaf: the reference loops over an EMPTY body range here */
                if (fix_near_leaves && d_sq > 0 && n->mass != 0.0f) {
                    const float inv = orc_fast_inv_sqrt(d_sq + e_sq);
                    const float s = n->mass * (inv * inv * inv);
                    ax += dx * s;
                    ay += dy * s;
                }
                if (n->next == 0) break;
                node = n->next;
            } else {
                node = n->children;
            }
        }
        acc_out[2 * (ii - (long long)i0)] = ax;
        acc_out[2 * (ii - (long long)i0) + 1] = ay;
    }
}

void orc_free(void *p) { free(p); }

/* ============================================================================ collision pass */
/* SpatialGrid::hash_position, Simulation.hpp:31-34: int arithmetic (wraps in practice), widened to size_t */
static uint64_t hash_position(int x, int y)
{
    const uint32_t h = (((uint32_t)x * 92837111u) ^ ((uint32_t)y * 689287499u)) * 15485863u;
    return (uint64_t)(int64_t)(int32_t)h;
}

/* Simulation::resolve, Simulation.hpp:293-346.  b1/b2 alias bodies[i]/bodies[j] in the reference, so
 * every read below happens at the same point of the update sequence as there. */
int orc_resolve(orc_body_t *B, size_t i, size_t j)
{
    orc_body_t *b1 = &B[i], *b2 = &B[j];
    float dx = b2->px - b1->px, dy = b2->py - b1->py;
    const float r = b1->radius + b2->radius;
    if (dx * dx + dy * dy > r * r) return 0;
    const float vx = b2->vx - b1->vx, vy = b2->vy - b1->vy;
    const float d_dot_v = dx * vx + dy * vy; /* Vec2::dot (dpps 0x31 == product, product, add) */
    const float m1 = b1->mass, m2 = b2->mass;
    const float weight1 = m2 / (m1 + m2);
    const float weight2 = m1 / (m1 + m2);
    if (d_dot_v >= 0.0f && !(dx == 0.0f && dy == 0.0f)) { /* separating: push apart, :312-318 */
        const float k = r / sqrtf(dx * dx + dy * dy) - 1.0f;
        const float tx = dx * k, ty = dy * k;
        b1->px -= tx * weight1; b1->py -= ty * weight1;
        b2->px += tx * weight2; b2->py += ty * weight2;
        return 1;
    }
    const float v_sq = vx * vx + vy * vy;
    const float d_sq = dx * dx + dy * dy;
    const float r_sq = r * r;
    float disc = d_dot_v * d_dot_v - v_sq * (d_sq - r_sq);
    if (disc < 0.0f) disc = 0.0f;
    const float t = (d_dot_v + sqrtf(disc)) / v_sq;
    b1->px -= b1->vx * t; b1->py -= b1->vy * t;           /* rewind to the moment of contact */
    b2->px -= b2->vx * t; b2->py -= b2->vy * t;
    const float ndx = b2->px - b1->px, ndy = b2->py - b1->py;
    const float nd_dot_v = ndx * vx + ndy * vy;
    const float nd_sq = ndx * ndx + ndy * ndy;
    const float k = 1.5f * nd_dot_v / nd_sq;              /* new_d * (1.5f * new_d_dot_v / new_d_sq) */
    const float tx = ndx * k, ty = ndy * k;
    const float v1x = b1->vx + tx * weight1, v1y = b1->vy + ty * weight1;
    const float v2x = b2->vx - tx * weight2, v2y = b2->vy - ty * weight2;
    b1->vx = v1x; b1->vy = v1y;
    b2->vx = v2x; b2->vy = v2y;
    b1->px += v1x * t; b1->py += v1y * t;
    b2->px += v2x * t; b2->py += v2y * t;
    return 1;
}

typedef struct { uint64_t key; uint32_t body; } cell_entry_t;
typedef struct { float value; uint32_t body; int is_end; } sweep_entry_t;
typedef struct { uint32_t a, b; } pair_t;

static int cmp_cell(const void *pa, const void *pb)
{
    const cell_entry_t *a = (const cell_entry_t *)pa, *b = (const cell_entry_t *)pb;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return (a->body > b->body) - (a->body < b->body);
}
static int cmp_sweep(const void *pa, const void *pb) /* SweepEntry::operator<, :43-46 */
{
    const sweep_entry_t *a = (const sweep_entry_t *)pa, *b = (const sweep_entry_t *)pb;
    if (a->value != b->value) return a->value < b->value ? -1 : 1;
    if (a->is_end != b->is_end) return a->is_end < b->is_end ? -1 : 1;
    return (a->body > b->body) - (a->body < b->body); /* exact ties: unspecified in the reference */
}
static int cmp_pair(const void *pa, const void *pb)
{
    const pair_t *a = (const pair_t *)pa, *b = (const pair_t *)pb;
    if (a->a != b->a) return a->a < b->a ? -1 : 1;
    return (a->b > b->b) - (a->b < b->b);
}

size_t orc_collide(orc_body_t *B, size_t n, size_t *resolved)
{
    const float CELL = 600.0f;
    size_t ne = 0, cap = 4 * n + 64;
    cell_entry_t *ent = (cell_entry_t *)malloc(cap * sizeof *ent);
    for (size_t i = 0; i < n; ++i) { /* :224-243 */
        const float r = B[i].radius;
        const int minX = (int)((B[i].px - r) / CELL), maxX = (int)((B[i].px + r) / CELL);
        const int minY = (int)((B[i].py - r) / CELL), maxY = (int)((B[i].py + r) / CELL);
        for (int y = minY; y <= maxY; ++y)
            for (int x = minX; x <= maxX; ++x) {
                if (ne == cap) { cap *= 2; ent = (cell_entry_t *)realloc(ent, cap * sizeof *ent); }
                ent[ne].key = hash_position(x, y);
                ent[ne].body = (uint32_t)i;
                ++ne;
            }
    }
    qsort(ent, ne, sizeof *ent, cmp_cell);
    size_t np = 0, pcap = 1024;
    pair_t *pairs = (pair_t *)malloc(pcap * sizeof *pairs);
    sweep_entry_t *sw = NULL; uint32_t *active = NULL; size_t swcap = 0;
    for (size_t s = 0; s < ne;) {
        size_t e = s;
        while (e < ne && ent[e].key == ent[s].key) ++e;
        const size_t k = e - s;
        if (k > 1) { /* :247-283 */
            if (2 * k > swcap) { swcap = 2 * k; sw = (sweep_entry_t *)realloc(sw, swcap * sizeof *sw); active = (uint32_t *)realloc(active, swcap * sizeof *active); }
            for (size_t q = 0; q < k; ++q) {
                const orc_body_t *bd = &B[ent[s + q].body];
                sw[2 * q] = (sweep_entry_t){bd->px - bd->radius, ent[s + q].body, 0};
                sw[2 * q + 1] = (sweep_entry_t){bd->px + bd->radius, ent[s + q].body, 1};
            }
            qsort(sw, 2 * k, sizeof *sw, cmp_sweep);
            size_t na = 0;
            for (size_t q = 0; q < 2 * k; ++q) {
                if (!sw[q].is_end) {
                    for (size_t a = 0; a < na; ++a) {
                        if (np == pcap) { pcap *= 2; pairs = (pair_t *)realloc(pairs, pcap * sizeof *pairs); }
                        pairs[np].a = active[a]; pairs[np].b = sw[q].body; ++np;
                    }
                    active[na++] = sw[q].body;
                } else {
                    size_t w = 0;
                    for (size_t a = 0; a < na; ++a) if (active[a] != sw[q].body) active[w++] = active[a];
                    na = w;
                }
            }
        }
        s = e;
    }
    qsort(pairs, np, sizeof *pairs, cmp_pair); /* canonical order (the reference's is unspecified) */
    size_t nres = 0;
    for (size_t p = 0; p < np; ++p) nres += (size_t)orc_resolve(B, pairs[p].a, pairs[p].b);
    if (resolved) *resolved = nres;
    free(ent); free(pairs); free(sw); free(active);
    return np;
}

/* ============================================================================ octree (3-D generalisation) */
typedef struct { orc_node3_t *v; size_t n, cap; } node3vec_t;
static void n3_push(node3vec_t *a, orc_node3_t x)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (orc_node3_t *)realloc(a->v, a->cap * sizeof *a->v);
    }
    a->v[a->n++] = x;
}
static orc_node3_t make_node3(uint64_t next, float cx, float cy, float cz, float size, uint64_t depth)
{
    orc_node3_t nd;
    nd.px = nd.py = nd.pz = 0.0f; nd.mass = 0.0f;
    nd.cx = cx; nd.cy = cy; nd.cz = cz; nd.size = size;
    nd.children = 0; nd.next = next; nd.depth = depth;
    return nd;
}
static size_t find_octant(const orc_node3_t *nd, float x, float y, float z)
{
    return ((size_t)(z > nd->cz) << 2) | ((size_t)(y > nd->cy) << 1) | (size_t)(x > nd->cx);
}
static void bh3_insert(node3vec_t *nodes, idxvec_t *parents, float x, float y, float z, float mass)
{
    size_t node = 0;
    while (nodes->v[node].children != 0) node = nodes->v[node].children + find_octant(&nodes->v[node], x, y, z);
    if (nodes->v[node].mass == 0.0f) {
        nodes->v[node].px = x; nodes->v[node].py = y; nodes->v[node].pz = z; nodes->v[node].mass = mass;
        return;
    }
    const float ex = nodes->v[node].px, ey = nodes->v[node].py, ez = nodes->v[node].pz, em = nodes->v[node].mass;
    if (x == ex && y == ey && z == ez) { nodes->v[node].mass += mass; return; }
    for (;;) {
        const size_t children = nodes->n;
        nodes->v[node].children = children;
        iv_push(parents, node);
        for (size_t i = 0; i < 8; ++i) {
            const orc_node3_t parent = nodes->v[node];
            const float ns = parent.size * 0.5f;
            const float cx = parent.cx + ((i & 1) ? 0.5f : -0.5f) * ns;
            const float cy = parent.cy + ((i & 2) ? 0.5f : -0.5f) * ns;
            const float cz = parent.cz + ((i & 4) ? 0.5f : -0.5f) * ns;
            n3_push(nodes, make_node3((i < 7) ? children + i + 1 : parent.next, cx, cy, cz, ns, parent.depth + 1));
        }
        const size_t q1 = find_octant(&nodes->v[node], ex, ey, ez), q2 = find_octant(&nodes->v[node], x, y, z);
        if (q1 == q2) { node = children + q1; continue; }
        orc_node3_t *a = &nodes->v[children + q1], *c = &nodes->v[children + q2];
        a->px = ex; a->py = ey; a->pz = ez; a->mass = em;
        c->px = x; c->py = y; c->pz = z; c->mass = mass;
        return;
    }
}

size_t orc_bh3_build(const orc_body_t *b, size_t n, orc_node3_t **nodes_out)
{
    node3vec_t nodes = {0, 0, 0};
    idxvec_t parents = {0, 0, 0};
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (size_t i = 0; i < n; ++i) {
        const float p[3] = {b[i].px, b[i].py, b[i].pz};
        for (int a = 0; a < 3; ++a) { mn[a] = p[a] < mn[a] ? p[a] : mn[a]; mx[a] = mx[a] < p[a] ? p[a] : mx[a]; }
    }
    float size = mx[0] - mn[0];
    for (int a = 1; a < 3; ++a) { const float e = mx[a] - mn[a]; size = size < e ? e : size; }
    n3_push(&nodes, make_node3(0, (mn[0] + mx[0]) * 0.5f, (mn[1] + mx[1]) * 0.5f, (mn[2] + mx[2]) * 0.5f, size, 0));
    for (size_t i = 0; i < n; ++i) bh3_insert(&nodes, &parents, b[i].px, b[i].py, b[i].pz, b[i].mass);
    for (size_t k = parents.n; k-- > 0;) {
        const size_t node = parents.v[k], child = nodes.v[node].children;
        float px = 0.0f, py = 0.0f, pz = 0.0f, m = 0.0f;
        for (size_t i = 0; i < 8; ++i) {
            const orc_node3_t *c = &nodes.v[child + i];
            px += c->px * c->mass; py += c->py * c->mass; pz += c->pz * c->mass;
            m += c->mass;
        }
        if (m > 0) { const float inv = 1.0f / m; px *= inv; py *= inv; pz *= inv; }
        nodes.v[node].px = px; nodes.v[node].py = py; nodes.v[node].pz = pz; nodes.v[node].mass = m;
    }
    free(parents.v);
    *nodes_out = nodes.v;
    return nodes.n;
}

void orc_bh3_acc(const orc_node3_t *nodes, float theta, float eps, const orc_body_t *b, size_t i0, size_t i1,
                 int fix_near_leaves, float *acc_out)
{
    const float t_sq = theta * theta, e_sq = eps * eps;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long ii = (long long)i0; ii < (long long)i1; ++ii) {
        const float px = b[ii].px, py = b[ii].py, pz = b[ii].pz;
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        size_t node = 0;
        for (;;) {
            const orc_node3_t *n = &nodes[node];
            const float dx = n->px - px, dy = n->py - py, dz = n->pz - pz;
            const float d_sq = (dx * dx + dy * dy) + dz * dz;
            const int far = n->size * n->size < d_sq * t_sq;
            if (far || n->children == 0) {
                if ((far || (fix_near_leaves && n->mass != 0.0f)) && d_sq > 0) {
                    const float inv = orc_fast_inv_sqrt(d_sq + e_sq);
                    const float s = n->mass * (inv * inv * inv);
                    ax += dx * s; ay += dy * s; az += dz * s;
                }
                if (n->next == 0) break;
                node = n->next;
            } else {
                node = n->children;
            }
        }
        float *o = acc_out + 3 * (ii - (long long)i0);
        o[0] = ax; o[1] = ay; o[2] = az;
    }
}
