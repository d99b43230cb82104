/* nbody_body.h -- the body-state record handed across the drop-in boundary.
 *
 * Byte-for-byte the reference's `struct alignas(16) Body` (Nbodysim/headers/Body.hpp:6-14):
 * three `Vec2` (each `alignas(16) {float x, y;}`, i.e. 16 bytes with 8 bytes of tail padding,
 * Nbodysim/headers/Vec2.hpp:17-20) followed by mass and radius.  sizeof == 64; offsets
 * pos 0 / vel 16 / acc 32 / mass 48 / radius 52.  A `std::vector<Body>::data()` pointer from the
 * reference can be passed wherever `nbody_body_t *` is expected, with no conversion.
 *
 * 3-D runs (the reference itself is 2-D) carry z in the first padding float of each Vec2, so
 * every reference offset is preserved and a 2-D caller that leaves padding at zero gets the
 * 2-D result term for term.
 */
#ifndef NBODY_BODY_H
#define NBODY_BODY_H
#include <stddef.h>
#include <stdint.h>

typedef struct nbody_body {
    float pos[2];  float pos_z;  float _pad0;   /* Body::pos    @0  */
    float vel[2];  float vel_z;  float _pad1;   /* Body::vel    @16 */
    float acc[2];  float acc_z;  float _pad2;   /* Body::acc    @32 */
    float mass;                                 /* Body::mass   @48 */
    float radius;                               /* Body::radius @52 */
    float _pad3[2];
} nbody_body_t;

#if defined(__cplusplus)
static_assert(sizeof(nbody_body_t) == 64, "nbody_body_t must match the reference Body (64 B)");
static_assert(offsetof(nbody_body_t, vel) == 16 && offsetof(nbody_body_t, acc) == 32 &&
              offsetof(nbody_body_t, mass) == 48 && offsetof(nbody_body_t, radius) == 52,
              "nbody_body_t offsets must match Body.hpp:6-14");
#else
_Static_assert(sizeof(nbody_body_t) == 64, "nbody_body_t must match the reference Body (64 B)");
_Static_assert(offsetof(nbody_body_t, vel) == 16 && offsetof(nbody_body_t, acc) == 32 &&
               offsetof(nbody_body_t, mass) == 48 && offsetof(nbody_body_t, radius) == 52,
               "nbody_body_t offsets must match Body.hpp:6-14");
#endif
#endif /* NBODY_BODY_H */
