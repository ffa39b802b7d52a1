/*
 * fxg_synth.h — deterministic, integer-only synthetic read generator.
 *
 * Counter-based splitmix64 so that plain C (host tools, oracle), CUDA (on-device generation
 * for the 100 M – 500 M read configs) and Python (tests) all produce the same bytes for
 * (seed, read_idx, pos).  No floating point anywhere.
 *
 * Usable from C, C++ and CUDA (functions are static inline / __host__ __device__).
 */
#ifndef FXG_SYNTH_H
#define FXG_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define FXG_HD __host__ __device__ __forceinline__
#else
#define FXG_HD static inline
#endif

/* Workload kinds */
#define FXG_SYNTH_PLAIN     0   /* uniform ACGT, quality decays along the read                        */
#define FXG_SYNTH_WITH_N    1   /* as PLAIN, plus an 'N' with probability 1/1024 per base             */
#define FXG_SYNTH_ADAPTER   2   /* 30 % of reads carry AGATCGGAAGAGC at a uniform start in [20, L-1]  */
#define FXG_SYNTH_DUPS      3   /* ~40 % of reads repeat a member of a skewed pool (collapser)        */

#define FXG_SYNTH_SEED_BASE 20260925ull

FXG_HD uint64_t fxg_sm64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* Per-read random word. For FXG_SYNTH_DUPS the *content key* of a read may be shared. */
FXG_HD uint64_t fxg_synth_read_key(uint64_t seed, uint64_t read_idx, int kind, uint64_t n_reads)
{
    if (kind == FXG_SYNTH_DUPS) {
        uint64_t g = fxg_sm64(seed ^ 0x5bd1e995c0ffeeull ^ fxg_sm64(read_idx));
        if ((g % 100ull) < 40ull) {
            uint64_t pool = n_reads / 8ull + 1ull;
            uint64_t u = (g >> 8) & 0xFFFFFFull;              /* 24-bit uniform          */
            uint64_t pid = ((u * u) >> 24) * pool >> 24;     /* squared => skewed to 0  */
            return fxg_sm64(seed ^ fxg_sm64(0x4000000000000000ull | pid));
        }
    }
    return fxg_sm64(seed ^ fxg_sm64(read_idx));
}

FXG_HD uint8_t fxg_synth_base(uint64_t r, int pos, int L, int kind)
{
    uint64_t h = fxg_sm64(r + (uint64_t)pos);
    uint8_t b = (uint8_t)(0x54474341u >> (8u * (unsigned)(h & 3ull)));   /* "ACGT"[h & 3] */
    if (kind == FXG_SYNTH_WITH_N && ((h >> 2) & 1023ull) == 0ull) b = (uint8_t)'N';
    if (kind == FXG_SYNTH_ADAPTER && ((r >> 32) % 100ull) < 30ull && L > 20) {
        int start = 20 + (int)((r >> 40) % (uint64_t)(L - 20));
        int off = pos - start;
        /* "AGATCGGAAGAGC"[off], packed little-endian so device code needs no string table */
        if (off >= 0 && off < 8) b = (uint8_t)(0x4147474354414741ull >> (8 * off));
        else if (off >= 8 && off < 13) b = (uint8_t)(0x4347414741ull >> (8 * (off - 8)));
    }
    return b;
}

/* Phred score (not yet offset by -Q) */
FXG_HD int fxg_synth_phred(uint64_t r, int pos, int L)
{
    uint64_t h = fxg_sm64(r + (uint64_t)pos);
    int noise = (int)((h >> 12) & 15ull) + (int)((h >> 16) & 15ull) + (int)((h >> 20) & 15ull) - 22;
    int q = 38 - (int)((18ll * pos * pos) / ((long long)L * L)) + noise;
    if (q < 2) q = 2;
    if (q > 40) q = 40;
    return q;
}

#endif /* FXG_SYNTH_H */
