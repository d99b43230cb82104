// TEST INFRASTRUCTURE ONLY -- never linked into, called from, or shipped with the product path.
//
// C-callable harness around the UNMODIFIED reference headers (7IBBE77S/nbodysim). The reference
// sources are #included from where they lie under /root/reference (see oracle/Makefile: the -I
// flag); nothing of theirs is copied into this repository. Outputs go to oracle/_ref/ only.
//
// What each entry point drives (reference file:line):
//   ref_fast_inv_sqrt   -> Quadtree::fast_inv_sqrt                   Quadtree.hpp:106-111
//   ref_allpairs_acc    -> Quadtree::acc leaf loop (direct sum)      Quadtree.hpp:133-144
//                          driven with theta=0 and a root leaf carrying Range(0,n), so the
//                          reference's own per-pair code performs the all-pairs sum; threaded
//                          exactly like Simulation::attract           Simulation.hpp:176-214
//   ref_step_clean      -> all-pairs force (above) + Body::update     Body.hpp:34-38
//   ref_bh_acc          -> Quadtree::build + Quadtree::acc            Quadtree.hpp:157-170,113-155
//   ref_bh_nodes*       -> node array after build (for tree-build parity)
//   ref_iterate/_collide/_step_full -> Simulation::iterate/collide/step Simulation.hpp:67-75,116-164,216-346
//   ref_uniform_disc    -> Simulation::uniform_disc                   Simulation.hpp:347-603
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <future>
#include <thread>
#include <atomic>
#include <algorithm>
#include <random>
#include <numeric>
#include <numbers>
#include <unordered_map>
#include <execution>
#include <mutex>
#include <array>
#include <iostream>
#include <limits>
#include <bit>

// Simulation keeps iterate/attract/collide/uniform_disc private (Simulation.hpp:77). The harness
// needs to call them one at a time; the macro changes access only, not a byte of reference code.
#define private public
#include "Simulation.hpp"
#undef private

// The reference declares this extern (Simulation.hpp:16) and defines it in main.cpp:39, which
// cannot be built here (raylib + a display). Same default value.
std::atomic<float> SIMULATION_DT{0.01f};

static_assert(sizeof(Body) == 64, "reference Body layout changed");
static_assert(offsetof(Body, pos) == 0 && offsetof(Body, vel) == 16 && offsetof(Body, acc) == 32 &&
              offsetof(Body, mass) == 48 && offsetof(Body, radius) == 52, "reference Body offsets");

namespace {

// One Simulation object reused by the full-step entry points; its constructor runs
// uniform_disc(25000) (Simulation.hpp:58-65), so build it lazily and only once.
Simulation *g_sim = nullptr;
Simulation &sim()
{
    if (!g_sim) g_sim = new Simulation();
    return *g_sim;
}

std::vector<Body> to_vec(const Body *b, size_t n) { return std::vector<Body>(b, b + n); }

unsigned pick_threads(int nthreads, size_t n)
{
    unsigned t = nthreads > 0 ? (unsigned)nthreads : std::max(1u, std::thread::hardware_concurrency());
    if (t > n) t = (unsigned)std::max<size_t>(1, n);
    return t;
}

} // namespace

extern "C" {

int ref_sizeof_body(void) { return (int)sizeof(Body); }
int ref_sizeof_node(void) { return (int)sizeof(Node); }
unsigned ref_hardware_threads(void) { return std::thread::hardware_concurrency(); }

float ref_fast_inv_sqrt(float x)
{
    Quadtree q(0.0f, 0.0f, 16);
    volatile float v = x; // defeat constant folding so the compiled arithmetic is what runs
    return q.fast_inv_sqrt(v);
}

// Accelerations of targets [i0,i1) against all n bodies through the reference's leaf loop.
// acc_out holds 2 floats per target, written at (i-i0).  nthreads<=0 -> all hardware threads.
void ref_allpairs_acc(const void *bodies_in, size_t n, float eps, size_t i0, size_t i1,
                      float *acc_out, int nthreads)
{
    std::vector<Body> bodies = to_vec((const Body *)bodies_in, n);
    Quadtree q(0.0f, eps, n);
    q.clear(Quad::new_containing(bodies));
    q.nodes[0].bodies = Range(0, n);
    const unsigned T = pick_threads(nthreads, i1 - i0);
    const size_t chunk = (i1 - i0 + T - 1) / T;
    std::vector<std::future<void>> futures;
    for (unsigned t = 0; t < T; ++t) {
        size_t s = i0 + t * chunk, e = std::min(s + chunk, i1);
        if (s >= e) break;
        futures.emplace_back(std::async(std::launch::async, [&, s, e]() {
            for (size_t i = s; i < e; ++i) {
                Vec2 a = q.acc(bodies[i].pos, bodies);
                acc_out[2 * (i - i0)] = a.x;
                acc_out[2 * (i - i0) + 1] = a.y;
            }
        }));
    }
    for (auto &f : futures) f.get();
}

// nsteps of { acc = all-pairs leaf loop ; Body::update(dt) } in place on the caller's bodies.
void ref_step_clean(void *bodies_io, size_t n, float eps, float dt, int nsteps, int nthreads)
{
    Body *B = (Body *)bodies_io;
    std::vector<Body> bodies = to_vec(B, n);
    std::vector<float> acc(2 * n);
    for (int s = 0; s < nsteps; ++s) {
        ref_allpairs_acc(bodies.data(), n, eps, 0, n, acc.data(), nthreads);
        for (size_t i = 0; i < n; ++i) {
            bodies[i].acc = Vec2(acc[2 * i], acc[2 * i + 1]);
            bodies[i].update(dt);
        }
    }
    std::memcpy((void *)B, (const void *)bodies.data(), n * sizeof(Body));
}

// Barnes-Hut accelerations exactly as Simulation::attract computes them (serial here).
// Returns the node count of the tree that was built.
size_t ref_bh_acc(const void *bodies_in, size_t n, float theta, float eps, float *acc_out)
{
    std::vector<Body> bodies = to_vec((const Body *)bodies_in, n);
    Quadtree q(theta, eps, n);
    q.build(bodies);
    for (size_t i = 0; i < n; ++i) {
        Vec2 a = q.acc(bodies[i].pos, bodies);
        acc_out[2 * i] = a.x;
        acc_out[2 * i + 1] = a.y;
    }
    return q.nodes.size();
}

// Flattened node array after Quadtree::build: 8 floats/doubles-free ints per node:
//   f[0..1]=com, f[2]=mass, f[3..4]=quad.center, f[5]=quad.size ; u[0]=children, u[1]=next, u[2]=depth
size_t ref_bh_nodes(const void *bodies_in, size_t n, float theta, float eps, float *f_out,
                    uint64_t *u_out, size_t cap)
{
    std::vector<Body> bodies = to_vec((const Body *)bodies_in, n);
    Quadtree q(theta, eps, n);
    q.build(bodies);
    size_t m = q.nodes.size();
    for (size_t k = 0; k < m && k < cap; ++k) {
        const Node &nd = q.nodes[k];
        f_out[6 * k + 0] = nd.data.pos.x;
        f_out[6 * k + 1] = nd.data.pos.y;
        f_out[6 * k + 2] = nd.data.mass;
        f_out[6 * k + 3] = nd.data.quad.center.x;
        f_out[6 * k + 4] = nd.data.quad.center.y;
        f_out[6 * k + 5] = nd.data.quad.size;
        u_out[3 * k + 0] = nd.children;
        u_out[3 * k + 1] = nd.next;
        u_out[3 * k + 2] = nd.depth;
    }
    return m;
}

// Simulation::iterate (BH theta/eps as given + clamp + soft boundary + drift), no collide.
void ref_iterate(void *bodies_io, size_t n, float theta, float eps, float dt, int nsteps)
{
    Simulation &S = sim();
    S.bodies = to_vec((const Body *)bodies_io, n);
    S.quadtree.t_sq = theta * theta;
    S.quadtree.e_sq = eps * eps;
    for (int s = 0; s < nsteps; ++s) S.iterate(dt);
    std::memcpy(bodies_io, (const void *)S.bodies.data(), n * sizeof(Body));
}

void ref_collide(void *bodies_io, size_t n)
{
    Simulation &S = sim();
    S.bodies = to_vec((const Body *)bodies_io, n);
    S.collide();
    std::memcpy(bodies_io, (const void *)S.bodies.data(), n * sizeof(Body));
}

// The reference's real step: Simulation::step() = iterate(SIMULATION_DT) ; collide() ; ++frame.
void ref_step_full(void *bodies_io, size_t n, float theta, float eps, float dt, int nsteps)
{
    Simulation &S = sim();
    S.bodies = to_vec((const Body *)bodies_io, n);
    S.quadtree.t_sq = theta * theta;
    S.quadtree.e_sq = eps * eps;
    SIMULATION_DT.store(dt);
    for (int s = 0; s < nsteps; ++s) S.step();
    std::memcpy(bodies_io, (const void *)S.bodies.data(), n * sizeof(Body));
}

void ref_uniform_disc(void *bodies_out, size_t n)
{
    std::vector<Body> b = sim().uniform_disc(n);
    std::memcpy(bodies_out, (const void *)b.data(), n * sizeof(Body));
}

} // extern "C"
