// fxg_device.cuh — sm_100a device primitives shared by the FASTX hot-path kernels:
//   * TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier transaction barriers
//   * SWAR (4 bytes per 32-bit lane) validation / compare helpers used by every kernel
//
// Validation restates src/libfastx/fastx.c:45-84 (bases in {A,C,G,T,N}; all six hot-path tools open
// the reader with ALLOW_N, REQUIRE_UPPERCASE) and fastx.c:118-135 (-15 <= byte-Q <= 93).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxg {

constexpr uint32_t ONES = 0x01010101u;
constexpr uint32_t HI   = 0x80808080u;

// ------------------------------------------------------------------------------------------------
// mbarrier + bulk-copy wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, bulk-group completion
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// SWAR helpers.  A 32-bit register holds 4 consecutive bytes of a read.
// ------------------------------------------------------------------------------------------------

// Quality constants for one launch, replicated into all 4 byte lanes.
struct QualK {
    uint32_t lo4;    // (Q-15) * ONES                       : smallest legal byte
    uint32_t hik4;   // (127 - min(Q+93,127)) * ONES         : x + hik4 sets bit7 iff x > hi (x < 128)
    uint32_t thr4;   // clamp(t+Q, 0, 128) * ONES            : op threshold in byte domain
};

__host__ __device__ inline QualK make_qualk(int q_offset, int thr_q)
{
    int lo = q_offset - 15;
    int hi = q_offset + 93; if (hi > 127) hi = 127;
    int thr = thr_q + q_offset; if (thr < 0) thr = 0; if (thr > 128) thr = 128;
    QualK k;
    k.lo4 = (uint32_t)lo * ONES;
    k.hik4 = (uint32_t)(127 - hi) * ONES;
    k.thr4 = (uint32_t)thr * ONES;
    return k;
}

// bit7 of each byte of the result is set iff that quality byte is illegal
// (outside [lo,hi] or >= 128; bytes >= 128 are negative `char`s in the reference).
__device__ __forceinline__ uint32_t qual_bad_bits(uint32_t x, uint32_t xh /* x | HI */, const QualK &k)
{
    uint32_t d = xh - k.lo4;        // bit7 = (x >= lo)      (no inter-byte borrow: xh >= 128 >= lo)
    uint32_t e = x + k.hik4;        // bit7 = (x > hi)       (x < 128; otherwise x's own bit7 flags it)
    return (~d) | e | x;
}

// bit7 of each byte set iff byte >= thr (valid for bytes < 128; thr in [0,128])
__device__ __forceinline__ uint32_t qual_ge_bits(uint32_t xh, const QualK &k) { return xh - k.thr4; }

// Base validation by table lookup: the low 3 bits of A,C,G,T,N are 1,3,7,4,6 — all distinct — so
// PRMT (byte permute) against an 8-entry byte table returns, for every byte, the only legal
// character with that code; any difference from the input marks an illegal byte.
//   table[code]: 0:-, 1:'A', 2:-, 3:'C', 4:'T', 5:-, 6:'N', 7:'G'   (unused codes hold 0xFF, whose own
//   code is 7, so they can never equal the byte that selected them — NUL included)
constexpr uint32_t VLUT_LO = 0x43FF41FFu;  // bytes 3..0 = 'C', -, 'A', -
constexpr uint32_t VLUT_HI = 0x474EFF54u;  // bytes 7..4 = 'G','N', -, 'T'
// complement table (src/fastx_reverse_complement/fastx_reverse_complement.c:43-72):
//   1:'A'->'T', 3:'C'->'G', 4:'T'->'A', 6:'N'->'N', 7:'G'->'C'
constexpr uint32_t CLUT_LO = 0x47005400u;
constexpr uint32_t CLUT_HI = 0x434E0041u;

// selector with one 3-bit code per nibble (low 16 bits), from the low 3 bits of each byte
__device__ __forceinline__ uint32_t base_selector(uint32_t x)
{
    uint32_t y = x & 0x07070707u;
    uint32_t z = y | (y >> 4);              // byte0 = c0|c1<<4, byte2 = c2|c3<<4
    return __byte_perm(z, 0u, 0x4420u);     // low 16 bits = nibbles c0,c1,c2,c3
}
// non-zero iff some byte of x is not one of A,C,G,T,N
__device__ __forceinline__ uint32_t seq_bad_bits(uint32_t x)
{
    uint32_t e = __byte_perm(VLUT_LO, VLUT_HI, base_selector(x));
    return x ^ e;
}
__device__ __forceinline__ uint32_t seq_complement(uint32_t x, uint32_t &bad)
{
    uint32_t sel = base_selector(x);
    bad |= x ^ __byte_perm(VLUT_LO, VLUT_HI, sel);
    return __byte_perm(CLUT_LO, CLUT_HI, sel);
}

// raw PRMT (PTX default mode): selector nibble bit 3 replicates the selected byte's sign bit instead of
// copying the byte — __byte_perm() only documents the low 3 bits of each nibble, so spell it in PTX.
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// byte mask with 0xFF for the first `nbytes` (0..4) bytes of a word
__device__ __forceinline__ uint32_t head_mask(int nbytes)
{
    return nbytes >= 4 ? 0xFFFFFFFFu : (nbytes <= 0 ? 0u : ((1u << (8 * nbytes)) - 1u));
}

__device__ __forceinline__ uint4 lds128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }

}  // namespace fxg
