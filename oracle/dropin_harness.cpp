// TEST INFRASTRUCTURE ONLY -- never linked into, called from, or shipped with the product path.
//
// The drop-in, executed: the reference's own `Simulation` class (Nbodysim/headers/Simulation.hpp) WITH the
// INTEGRATION.md patch applied (oracle/integration.patch: the body of Simulation::step() -- iterate(dt) +
// collide(), Simulation.hpp:67-75 -- replaced by nbody_gpu_step + nbody_gpu_download; init in the constructor),
// compiled against the reference's unmodified Body / Quadtree / Vec2 headers and linked with libnbody_gpu.so.
// The patched header is produced by `patch -o` into a scratch directory at build time and removed afterwards;
// only the executable lands in oracle/_ref/.  The reference tree is never written to.
//
//   dropin_patched <nsteps> <out.bin> [threaded]
//
// runs the patched Simulation -- constructor = uniform_disc(25000) + nbody_gpu_init -- for nsteps x step() and
// writes the raw 64-byte Body records of `simulation.bodies`.  tests/test_dropin.py compares them with the
// UNPATCHED Simulation::step() (oracle/_ref/libnbody_ref_strict.so: ref_uniform_disc + ref_step_full).
// With `threaded`, step() runs on a simulation thread that publishes `bodies` under a lock while the main
// thread consumes the copies at ~60 Hz, as simulation_thread / the render loop do (main.cpp:612-635, :659-660).
#include <algorithm>
#include <array>
#include <atomic>
#include <bit>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <numbers>
#include <numeric>
#include <random>
#include <thread>
#include <unordered_map>
#include <vector>

#include "Simulation.hpp"   // the PATCHED copy: first on the include path; it includes the reference's other headers

std::atomic<float> SIMULATION_DT{0.01f};   // main.cpp:39 (cannot be built here: raylib + a display)

static std::mutex UPDATE_LOCK;             // main.cpp:41-43
static std::vector<Body> SHARED_BODIES;

int main(int argc, char **argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s <nsteps> <out.bin> [threaded]\n", argv[0]); return 2; }
    const int nsteps = std::atoi(argv[1]);
    const bool threaded = argc > 3 && std::strcmp(argv[3], "threaded") == 0;
    try {
        auto simulation = std::make_shared<Simulation>();
        const auto t0 = std::chrono::steady_clock::now();
        size_t frames = 0;
        if (!threaded) {
            for (int s = 0; s < nsteps; ++s) simulation->step();
        } else {
            std::atomic<bool> done{false};
            std::thread sim([&] {                                   // simulation_thread, main.cpp:612-635
                for (int s = 0; s < nsteps; ++s) {
                    simulation->step();
                    {
                        std::lock_guard<std::mutex> lock(UPDATE_LOCK);
                        SHARED_BODIES = simulation->bodies;
                    }
                    std::this_thread::sleep_for(std::chrono::milliseconds(1));
                }
                done = true;
            });
            std::vector<Body> local;                                // the render loop's copy, main.cpp:~700
            while (!done) {
                {
                    std::lock_guard<std::mutex> lock(UPDATE_LOCK);
                    local = SHARED_BODIES;
                }
                ++frames;
                std::this_thread::sleep_for(std::chrono::microseconds(16667));
            }
            sim.join();
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::FILE *f = std::fopen(argv[2], "wb");
        if (!f) { std::perror(argv[2]); return 1; }
        std::fwrite(simulation->bodies.data(), sizeof(Body), simulation->bodies.size(), f);
        std::fclose(f);
        std::printf("{\"steps\": %d, \"frame\": %zu, \"n\": %zu, \"ms_per_step\": %.4f, \"threaded\": %s, \"render_frames\": %zu}\n", nsteps,
                    simulation->frame, simulation->bodies.size(), 1e3 * sec / std::max(1, nsteps), threaded ? "true" : "false", frames);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "dropin_patched: %s\n", e.what());
        return 1;
    }
    return 0;
}
