// Host build of the packed-s16x2 clipper DP (fastx_toolkit_b200/csrc/fxg_clip_dpx.cuh, portable emulation of the
// DPX ops) checked against the oracle's literal restatement of the reference aligner (fxo_align).
// Usage: clip_dpx_host <seed> <n_pairs> <L> <adapter>     exit status 0 = all fields equal for all reads
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "fxg_clip_dpx.cuh"
extern "C" {
#include "fastx_oracle.h"
}

static uint64_t sm64(uint64_t &s)
{
    uint64_t x = (s += 0x9E3779B97F4A7C15ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int HMAX>
static int run(uint64_t seed, int npairs, int L, const char *adapter)
{
    const int H = (int)strlen(adapter);
    std::vector<uint8_t> r0(L + 8), r1(L + 8);
    uint64_t st = seed;
    int bad = 0, flagged = 0;
    for (int p = 0; p < npairs; p++) {
        uint8_t *rows[2] = { r0.data(), r1.data() };
        for (int r = 0; r < 2; r++) {
            uint8_t *row = rows[r];
            for (int i = 0; i < L; i++) row[i] = "ACGT"[sm64(st) & 3];
            const uint64_t k = sm64(st);
            if ((k & 3) != 0 && L > 4) {      // plant a (possibly mutated / truncated / gapped) adapter
                int start = (int)((k >> 8) % (uint64_t)L);
                int j = 0;
                for (int i = start; i < L && j < H; i++) {
                    const uint64_t e = sm64(st) % 40;
                    if (e == 0) { row[i] = "ACGT"[sm64(st) & 3]; j++; }       // substitution
                    else if (e == 1) { j++; i--; }                            // deletion in the read
                    else if (e == 2) { row[i] = "ACGT"[sm64(st) & 3]; }       // insertion in the read
                    else row[i] = (uint8_t)adapter[j++];
                }
            }
            if ((k >> 40) % 50 == 0) row[(k >> 20) % (uint64_t)L] = 'N';      // goes to the exact path
            for (int i = L; i < L + 8; i++) row[i] = 0;
        }
        fxg::dpx::PairOut o;
        fxg::dpx::align_pair<HMAX, 256>(rows[0], rows[1], L, (const uint8_t *)adapter, H, o);
        for (int r = 0; r < 2; r++) {
            bool hasN = memchr(rows[r], 'N', (size_t)L) != NULL;
            if (((o.exact >> r) & 1u) != (hasN ? 1u : 0u)) { bad++; continue; }
            if (hasN) { flagged++; continue; }
            fxo_align_result a;
            fxo_align(rows[r], L, L, (const uint8_t *)adapter, H, &a);
            const int m = (int)(o.lo[r] & 127u), x = (int)((o.lo[r] >> 7) & 127u), ts = (int)((o.lo[r] >> 21) & 127u);
            const int g = (int)(o.hi[r] & 0x7FFFu), qs = (int)(o.hi[r] >> 15);
            if (m != a.matches || x != a.mismatches || a.neutral != 0 || g != a.gaps || qs != a.query_start || ts != a.target_start ||
                o.bx[r] != a.query_end || o.by[r] != a.target_end) {
                if (bad < 5)
                    fprintf(stderr, "MISMATCH pair %d read %d: dpx m%d x%d g%d qs%d ts%d end(%d,%d)  ref m%d x%d n%d g%d qs%d ts%d end(%d,%d) %.*s\n",
                            p, r, m, x, g, qs, ts, o.bx[r], o.by[r], a.matches, a.mismatches, a.neutral, a.gaps, a.query_start,
                            a.target_start, a.query_end, a.target_end, L, rows[r]);
                bad++;
            }
        }
    }
    printf("pairs=%d L=%d H=%d flagged=%d mismatches=%d\n", npairs, L, H, flagged, bad);
    return bad ? 1 : 0;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const uint64_t seed = strtoull(argv[1], NULL, 10);
    const int np = atoi(argv[2]), L = atoi(argv[3]);
    const char *ad = argv[4];
    const int H = (int)strlen(ad);
    if (H < 1 || H > 16 || L < 1 || L > 256) return 2;
    switch ((H + 3) / 4) {
    case 1: return run<4>(seed, np, L, ad);
    case 2: return run<8>(seed, np, L, ad);
    case 3: return run<12>(seed, np, L, ad);
    default: return run<16>(seed, np, L, ad);
    }
}
