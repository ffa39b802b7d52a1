// CPU check of the SWAR primitives every hot-path kernel is built from (fastx_toolkit_b200/csrc/fxg_device.cuh): the header
// is compiled for the host (PRMT emulated below; the PTX wrappers are never instantiated) and each primitive is compared,
// byte by byte and exhaustively over the byte values, with the reference's per-character rule:
//   bases   src/libfastx/fastx.c:45-84    (A C G T N, upper case only)
//   quality src/libfastx/fastx.c:118-135  (-15 <= byte - Q <= 93; bytes >= 128 are negative chars)
//   complement src/fastx_reverse_complement/fastx_reverse_complement.c:43-72
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
#include "fxg_device.cuh"
using namespace fxg;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }
static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails++ < 10) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

static int legal_base(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }
static uint8_t comp(uint8_t c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'N'; }

int main(void)
{
    const char legal[5] = { 'A', 'C', 'G', 'T', 'N' };
    // ---- bases: every byte value in every lane, the other lanes legal or arbitrary
    for (int pos = 0; pos < 4; pos++)
        for (int b = 0; b < 256; b++)
            for (int rep = 0; rep < 64; rep++) {
                uint8_t w[4];
                for (int k = 0; k < 4; k++) w[k] = (rep & 1) ? (uint8_t)rnd() : (uint8_t)legal[rnd() % 5];
                w[pos] = (uint8_t)b;
                uint32_t x; memcpy(&x, w, 4);
                const uint32_t bad = seq_bad_bits(x);
                uint32_t bad2 = 0;
                const uint32_t c = seq_complement(x, bad2);
                for (int k = 0; k < 4; k++) {
                    const int isbad = ((bad >> (8 * k)) & 0xFF) != 0, isbad2 = ((bad2 >> (8 * k)) & 0xFF) != 0;
                    CHECK(isbad == !legal_base(w[k]), "seq_bad_bits byte %d of %08x", k, x);
                    CHECK(isbad2 == !legal_base(w[k]), "seq_complement bad flag byte %d of %08x", k, x);
                    if (legal_base(w[k])) CHECK(((c >> (8 * k)) & 0xFF) == comp(w[k]), "complement byte %d of %08x -> %08x", k, x, c);
                }
            }
    // ---- qualities: every -Q the tools accept, every byte value in every lane
    for (int Q = 15; Q <= 127; Q++) {
        const int lo = Q - 15, hi = Q + 93 > 127 ? 127 : Q + 93;
        for (int thr_q = -20; thr_q <= 130; thr_q += (thr_q > 60 ? 7 : 1)) {
            const QualK k = make_qualk(Q, thr_q);
            int thr = thr_q + Q; if (thr < 0) thr = 0; if (thr > 128) thr = 128;
            for (int pos = 0; pos < 4; pos++)
                for (int b = 0; b < 256; b++) {
                    uint8_t w[4];
                    for (int j = 0; j < 4; j++) w[j] = (rnd() & 3) ? (uint8_t)(rnd() & 127) : (uint8_t)rnd();
                    w[pos] = (uint8_t)b;
                    uint32_t x; memcpy(&x, w, 4);
                    const uint32_t bad = qual_bad_bits(x, x | HI, k) & HI;
                    const uint32_t ge = qual_ge_bits(x | HI, k) & HI;
                    // the kernels OR the verdicts of a read together, so what must hold is: the word is flagged iff some byte
                    // is illegal; byte-exact attribution holds while no byte is >= 128 (a byte >= 128, illegal itself, may
                    // carry into its upper neighbour's "> hi" test)
                    int want_any = 0, any_high = 0;
                    for (int j = 0; j < 4; j++) { const int v = w[j]; want_any |= (v >= 128 || v < lo || v > hi); any_high |= v >= 128; }
                    CHECK((bad != 0) == (want_any != 0), "qual_bad_bits Q=%d word %08x", Q, x);
                    for (int j = 0; j < 4; j++) {
                        const int v = w[j];
                        const int want_bad = v >= 128 || v < lo || v > hi;
                        if (!any_high) CHECK((((bad >> (8 * j)) & 0x80) != 0) == want_bad, "qual_bad_bits Q=%d byte %d of %08x", Q, j, x);
                        if (v < 128) CHECK((((ge >> (8 * j)) & 0x80) != 0) == (v >= thr), "qual_ge_bits Q=%d t=%d byte %d of %08x", Q, thr_q, j, x);
                    }
                }
        }
    }
    // ---- head_mask
    for (int n = -5; n <= 9; n++) {
        const uint32_t m = head_mask(n);
        for (int j = 0; j < 4; j++) CHECK(((m >> (8 * j)) & 0xFF) == (j < n ? 0xFFu : 0u), "head_mask(%d) = %08x", n, m);
    }
    if (fails) { printf("%d failures\n", fails); return 1; }
    printf("swar primitives ok\n");
    return 0;
}
