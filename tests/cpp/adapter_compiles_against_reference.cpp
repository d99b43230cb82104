// Compile-only check (no GPU needed): the C ABI headers and the C++ adapter are usable from the
// reference's own translation-unit context -- the reference's `Body` (Nbodysim/headers/Body.hpp) can be
// handed to nbody_gpu_* / GpuSimulation<Body> without conversion.  Built by tests/test_integration_compile.py
// with:  g++ -std=c++20 -msse4.1 -I /root/reference/Nbodysim/headers -I include -c ...
#include "Body.hpp"                    // the reference's type, included from where it lies
#include "nbody_gpu.h"
#include "nbody_gpu_simulation.hpp"
#include <cstddef>
#include <vector>

static_assert(sizeof(Body) == sizeof(nbody_body_t), "reference Body and nbody_body_t differ in size");
static_assert(offsetof(Body, pos) == offsetof(nbody_body_t, pos), "pos offset");
static_assert(offsetof(Body, vel) == offsetof(nbody_body_t, vel), "vel offset");
static_assert(offsetof(Body, acc) == offsetof(nbody_body_t, acc), "acc offset");
static_assert(offsetof(Body, mass) == offsetof(nbody_body_t, mass), "mass offset");
static_assert(offsetof(Body, radius) == offsetof(nbody_body_t, radius), "radius offset");
static_assert(alignof(Body) == 16, "reference Body alignment");

// the INTEGRATION.md patch, as code: Simulation::step() with the GPU path behind it
struct PatchedSimulation {
    float dt = 0.01f;
    std::size_t frame = 0;
    std::vector<Body> bodies;
    nbody_ctx *gpu = nullptr;
    int init()
    {
        nbody_params p;
        nbody_params_default(&p);
        p.integ_flags = NBODY_INTEG_CLAMP | NBODY_INTEG_BOUNDARY;
        p.force_algo = NBODY_FORCE_BARNES_HUT;
        p.rsqrt_mode = NBODY_RSQRT_REFCOMPAT;
        p.collide = 1;
        return nbody_gpu_init(&gpu, &p, reinterpret_cast<const nbody_body_t *>(bodies.data()), bodies.size());
    }
    int step()
    {
        int rc = nbody_gpu_step(gpu, dt, 1);
        ++frame;
        return rc;
    }
    int publish(std::vector<Body> &shared)
    {
        shared.resize(bodies.size());
        return nbody_gpu_download(gpu, reinterpret_cast<nbody_body_t *>(shared.data()), shared.size(),
                                  NBODY_FIELD_POS | NBODY_FIELD_VEL);
    }
};

int adapter_smoke(std::vector<Body> initial)
{
    GpuSimulation<Body> sim(std::move(initial));   // template instantiates against the reference's Body
    sim.step();
    sim.sync_bodies();
    return static_cast<int>(sim.frame);
}
