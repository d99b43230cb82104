// nbody_gpu_simulation.hpp -- header-only C++ adapter with the member names of the reference's
// `Simulation` (Nbodysim/headers/Simulation.hpp:49-75) over the C ABI of nbody_gpu.h.
//
//   reference                                  this adapter
//   ---------------------------------------    -----------------------------------------------
//   Simulation()            :58-65             GpuSimulation(bodies, params)  -> nbody_gpu_init
//   float dt; size_t frame; :52-53             same members
//   std::vector<Body> bodies; :54              same member (host mirror, refreshed by sync_bodies())
//   void step()             :67-75             step(): nbody_gpu_step(ctx, SIMULATION_DT, 1); ++frame
//   attract()               :176-214 (private) attract(): nbody_gpu_accel_only
//
// `BodyT` is any 64-byte type laid out like the reference's Body (Body.hpp:6-14) -- e.g. the
// reference's own `Body`, so existing code that walks `simulation->bodies` keeps compiling.
#pragma once
#include "nbody_gpu.h"
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

template <typename BodyT = nbody_body_t>
class GpuSimulation {
    static_assert(sizeof(BodyT) == sizeof(nbody_body_t), "BodyT must be the reference's 64-byte Body");

public:
    float dt = 0.01f;   // SIMULATION_DT default, main.cpp:39; re-read on every step() like Simulation.hpp:69
    std::size_t frame = 0;
    std::vector<BodyT> bodies;

    explicit GpuSimulation(std::vector<BodyT> initial, const nbody_params *params = nullptr)
        : bodies(std::move(initial))
    {
        nbody_params p;
        if (params) p = *params; else nbody_params_default(&p);
        const int rc = nbody_gpu_init(&ctx_, &p, reinterpret_cast<const nbody_body_t *>(bodies.data()), bodies.size());
        if (rc != NBODY_OK)
            throw std::runtime_error(std::string("nbody_gpu_init: ") + nbody_gpu_strerror(rc) + ": " +
                                     nbody_gpu_last_error(nullptr));
    }
    GpuSimulation(const GpuSimulation &) = delete;
    GpuSimulation &operator=(const GpuSimulation &) = delete;
    ~GpuSimulation() { nbody_gpu_shutdown(ctx_); }

    // Simulation::step(): iterate(current_dt) ; ++frame.  (collide() stays on the host if wanted:
    // sync_bodies(); collide(); push_bodies();)
    void step(int nsteps = 1)
    {
        check(nbody_gpu_step(ctx_, dt, nsteps), "nbody_gpu_step");
        frame += static_cast<std::size_t>(nsteps);
    }
    // Simulation::attract(): fill bodies[i].acc for the current positions (visible after sync_bodies)
    void attract() { check(nbody_gpu_accel_only(ctx_), "nbody_gpu_accel_only"); }
    // the analogue of `SHARED_BODIES = simulation->bodies` (main.cpp:625)
    void sync_bodies(unsigned fields = NBODY_FIELD_ALL)
    {
        check(nbody_gpu_download(ctx_, reinterpret_cast<nbody_body_t *>(bodies.data()), bodies.size(), fields),
              "nbody_gpu_download");
    }
    void push_bodies()
    {
        check(nbody_gpu_upload(ctx_, reinterpret_cast<const nbody_body_t *>(bodies.data()), bodies.size()),
              "nbody_gpu_upload");
    }
    void energy(double &K, double &W, double P[3]) { check(nbody_gpu_energy(ctx_, &K, &W, P), "nbody_gpu_energy"); }
    nbody_ctx *handle() { return ctx_; }

private:
    void check(int rc, const char *what)
    {
        if (rc != NBODY_OK)
            throw std::runtime_error(std::string(what) + ": " + nbody_gpu_strerror(rc) + ": " + nbody_gpu_last_error(ctx_));
    }
    nbody_ctx *ctx_ = nullptr;
};
