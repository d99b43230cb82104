// Host-side arithmetic of csrc/radix_sort.cuh (no GPU needed): the scratch reserved for a bound n must hold the status words
// of EVERY smaller input -- workspaces are allocated once for their capacity and then sort fewer items, which the
// size-dependent tile (512 / 1,024 / 2,048 / 4,096 keys) turns into MORE tiles per key at small sizes.
#include "../../nbodysim_b200/csrc/radix_sort.cuh"
#include <cstdio>
using namespace nb;

static size_t tiles_of(size_t n) { const size_t tile = (size_t)RS_THREADS * rs_rows_for(n); return (n + tile - 1) / tile; }

int main()
{
    int bad = 0;
    const size_t head = (size_t)(RS_MAX_PASSES * 256 + RS_MISC_WORDS) * sizeof(unsigned);
    const size_t caps[] = {1, 511, 512, 513, 25000, 40000, 40001, 104096, 200000, 200001, 1000000, 6291456, 6291457, 16777216, 70000000};
    for (size_t cap : caps) {
        const size_t words = (radix_sort_temp_bytes(cap) - head) / sizeof(unsigned long long) / 256;      // tiles the scratch holds
        // every tier boundary below the capacity, the capacity itself, and a sweep
        size_t probes[64]; int np = 0;
        const size_t marks[] = {1, RS_ROWS2_MAX_N - 1, RS_ROWS2_MAX_N, RS_ROWS2_MAX_N + 1, RS_ROWS4_MAX_N, RS_ROWS4_MAX_N + 1,
                                (size_t)RS_SMALL_TILE_MAX_N, (size_t)RS_SMALL_TILE_MAX_N + 1, cap};
        for (size_t m : marks) if (m <= cap) probes[np++] = m;
        for (int k = 1; k <= 32; ++k) probes[np++] = cap * k / 32 ? cap * k / 32 : 1;
        for (int q = 0; q < np; ++q)
            if (tiles_of(probes[q]) > words) { printf("capacity %zu: %zu items need %zu tiles, scratch holds %zu\n", cap, probes[q], tiles_of(probes[q]), words); ++bad; }
    }
    // the tiers themselves
    if (rs_rows_for(25000) != 2 || rs_rows_for(100000) != 4 || rs_rows_for(1000000) != RS_ROWS_SMALL || rs_rows_for(16777216) != RS_ROWS_LARGE) { printf("tiers\n"); ++bad; }
    // pass-id space: high digits first uses up to 2 x RS_MAX_PASSES tickets below the barrier word, the stamps start above the flag
    if (!(2 * RS_MAX_PASSES <= RS_MISC_BARRIER && RS_MISC_FLAG < RS_MISC_STAMPS && RS_MISC_STAMPS + 2 * RS_MISC_NSTAMPS <= RS_MISC_WORDS)) { printf("misc layout\n"); ++bad; }
    printf(bad ? "FAILED\n" : "OK\n");
    return bad ? 1 : 0;
}
