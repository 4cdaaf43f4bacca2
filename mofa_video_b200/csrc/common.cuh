// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 / TMEM PTX wrappers.
// Everything here is inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MOFA_DEVICE __device__ __forceinline__

namespace mofa {

// ---------------------------------------------------------------------------------------------
// error plumbing shared by every C-ABI entry point
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

// ---------------------------------------------------------------------------------------------
// basic
// ---------------------------------------------------------------------------------------------
MOFA_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
MOFA_DEVICE uint32_t lane_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
    return r;
}
MOFA_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
MOFA_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MOFA_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MOFA_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

MOFA_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
// non-blocking probe of a phase (for a thread that multiplexes several barriers)
MOFA_DEVICE bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
MOFA_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MOFA_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
MOFA_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
MOFA_DEVICE void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
MOFA_DEVICE void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
MOFA_DEVICE void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// TMA stores (shared -> global, bulk-group completion); out-of-bounds parts of the box are clipped
MOFA_DEVICE void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
MOFA_DEVICE void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
MOFA_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the shared-memory source of all but the newest `kPending` bulk groups has been read
template <int kPending>
MOFA_DEVICE void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending>
MOFA_DEVICE void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
MOFA_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MOFA_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole-warp: allocate `cols` (power of two >= 32) TMEM columns, base address written to *dst (smem)
MOFA_DEVICE void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MOFA_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate; issued by ONE thread
MOFA_DEVICE void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is a 128-lane x (K/2)-column block of packed fp16 pairs in tensor memory
MOFA_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
MOFA_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE tcgen05.mma of M = 256.  Each CTA
// stages its own 128 rows of A and HALF of the B rows, so the shared-memory traffic per SM and flop drops by a quarter to
// a third; the accumulator rows 0..127 land in the even CTA's tensor memory, rows 128..255 in the odd CTA's.
// Shared-memory addresses of the odd CTA carry bit 24 in the shared::cluster window; clearing it names the same offset
// in the even (leader) CTA -- the convention the 2-SM TMA / commit forms rely on.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
MOFA_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
MOFA_DEVICE void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// loads into THIS CTA's shared memory, completion bytes counted on the LEADER CTA's mbarrier
MOFA_DEVICE void tma_load_2d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
MOFA_DEVICE void tma_load_4d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// arrive on the mbarrier at this offset in the leader CTA (from either CTA of the pair)
MOFA_DEVICE void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// executed by one warp of EACH CTA of the pair (same warp index, same destination offset)
MOFA_DEVICE void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
MOFA_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[256 x N] (+)= A[256 x 16] * B[N x 16]^T: issued by ONE thread of the leader CTA; the descriptors name offsets that are
// valid in both CTAs (A: that CTA's 128 rows, B: that CTA's N/2 rows)
MOFA_DEVICE void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once the pair's previously issued MMAs have completed
MOFA_DEVICE void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t holds row lane_base+t)
MOFA_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
MOFA_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
MOFA_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the reverse: 32 registers per thread -> 32 lanes x 32 consecutive columns
MOFA_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]),
        "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
        : "memory");
}
// 8-column variants (rarely taken paths that must not inflate the register budget)
MOFA_DEVICE void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
MOFA_DEVICE void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(taddr)
                 : "memory");
}
MOFA_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor for a K-major tile stored as 128-byte rows with SWIZZLE_128B
// (the layout a TMA box with inner extent 64 fp16 and CU_TENSOR_MAP_SWIZZLE_128B produces):
// 8-row x 128 B swizzle atoms stacked along M/N with a 1024 B stride (SBO); LBO unused.
MOFA_DEVICE uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address, 16 B units
    d |= static_cast<uint64_t>(1024 >> 4) << 32;            // stride byte offset between 8-row groups
    d |= 1ull << 46;                                        // descriptor version (sm_100)
    d |= 2ull << 61;                                        // SWIZZLE_128B
    return d;
}
// MN-major operand (the contiguous dimension is M/N, e.g. V[kv][d] used as B with N=d, K=kv):
// atoms are 8 K-rows x 128 B (64 fp16 along N); consecutive K-groups 1024 B apart (SBO);
// LBO = byte distance between 64-wide N atoms (unused when the tile is exactly 64 wide).
MOFA_DEVICE uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, M = 128 (or 256 for a CTA pair)
MOFA_DEVICE uint32_t umma_idesc_f16_m(uint32_t n, uint32_t m) {
    uint32_t d = 0;
    d |= 1u << 4;           // D format: F32
    d |= (n >> 3) << 17;    // N / 8
    d |= (m >> 4) << 24;    // M / 16
    return d;
}
MOFA_DEVICE uint32_t umma_idesc_f16(uint32_t n, bool b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                     // D format: F32
    d |= 0u << 7;                     // A format: F16
    d |= 0u << 10;                    // B format: F16
    d |= (b_mn_major ? 1u : 0u) << 16;  // B major
    d |= (n >> 3) << 17;              // N / 8
    d |= (128u >> 4) << 24;           // M / 16
    return d;
}

MOFA_DEVICE float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for x <= 0 on the FMA / integer pipes only (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, cubic
// minimax for 2^f (relative error 7.5e-5, below the fp16 resolution of the probabilities it feeds), n added to the
// exponent field.  x is clamped at -126 (results there are ~1e-38, i.e. zero after the fp16 conversion).
MOFA_DEVICE float poly_exp2(float x) {
    x = fmaxf(x, -126.0f);
    const float r = x + 12582912.0f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (r - 12582912.0f);
    float q = fmaf(0.05517027f, f, 0.24260795f);
    q = fmaf(q, f, 0.69326093f);
    q = fmaf(q, f, 0.99992828f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}
// two at a time on the packed-fp32 FMA pipe (FFMA2 / FADD2, sm_100): 5 instructions per element instead of 8
MOFA_DEVICE float2 poly_exp2x2(float2 x) {
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
    const float2 r = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
    const float2 t = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(t, make_float2(-1.0f, -1.0f), x);
    float2 q = __ffma2_rn(make_float2(0.05517027f, 0.05517027f), f, make_float2(0.24260795f, 0.24260795f));
    q = __ffma2_rn(q, f, make_float2(0.69326093f, 0.69326093f));
    q = __ffma2_rn(q, f, make_float2(0.99992828f, 0.99992828f));
    float2 o;
    o.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
    o.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
    return o;
}
MOFA_DEVICE float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x * sigmoid(x).  Raw ex2 / rcp approximations: the CUDA intrinsics (__expf, __fdividef) wrap each MUFU in 4-5 extra
// range-fixing instructions when the file is not built with -use_fast_math, and those dominated the epilogues.
MOFA_DEVICE float sigmoid_f(float x) { return fast_rcp(1.0f + fast_exp2(x * -1.4426950408889634f)); }
MOFA_DEVICE float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x * sigmoid(x) with ONE MUFU: sigmoid(x) = 0.5 tanh(x/2) + 0.5, so SiLU = h tanh(h) + h with h = x/2 (MUFU.TANH, relative
// error 2^-11 = the fp16 output resolution).  GroupNorm+SiLU apply is MUFU / issue bound, not HBM bound: 2 MUFU per element
// were 78 us of a 141 us level-0 launch (295 M operations at 16 per clock per SM).
MOFA_DEVICE float silu_f(float x) {
    const float h = 0.5f * x;
    return fmaf(h, fast_tanh(h), h);
}
// GELU(x) = x * Phi(x), Phi by Abramowitz & Stegun 7.1.26 on |x|/sqrt(2) (|error| <= 1.5e-7 on erf, far below fp16
// output resolution):  q = 0.5 * t * poly(t) * exp(-x^2/2),  t = 1 / (1 + p |x| / sqrt(2)),  Phi = x < 0 ? q : 1 - q.
// 2 MUFU + 13 FMA/ALU instructions; the negative branch has no cancellation.  The GEGLU epilogue is issue-bound.
MOFA_DEVICE float gelu_phi(float x) {
    const float ax = fabsf(x);
    const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float e = fast_exp2((x * x) * (-0.5f * 1.4426950408889634f));
    const float q = (poly * t) * e;
    return x < 0.0f ? q : 1.0f - q;
}
MOFA_DEVICE float gelu_erf_f(float x) { return x * gelu_phi(x); }
// The same GELU without the sign select: with q = 1 - Phi(|x|) (the A&S tail),  x Phi(x) = max(x, 0) - |x| q  for either sign.
// 2 MUFU + 12 FMA-pipe instructions (|x| is an operand modifier); used by the GEGLU epilogue, which is issue-bound.
MOFA_DEVICE float gelu_erf_relu_form(float x) {
    const float ax = fabsf(x);
    const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float e = fast_exp2((x * x) * (-0.5f * 1.4426950408889634f));
    const float q = (poly * t) * e;
    return fmaf(-ax, q, fmaxf(x, 0.0f));
}
// GEGLU of two adjacent columns on the packed-fp32 pipe (FFMA2 / FMUL2 / FADD2): (v + bv) * gelu(g + bg), the same
// arithmetic as gelu_erf_relu_form per element, ~13 issue slots per element instead of ~19 (the GEGLU epilogues are
// issue-bound: 16 epilogue warps at K = 320)
MOFA_DEVICE float2 geglu2(float2 v, float2 g, __half2 bv, __half2 bg) {
    const float2 x = __fadd2_rn(g, __half22float2(bg));
    const float2 val = __fadd2_rn(v, __half22float2(bv));
    const float2 nax = make_float2(-fabsf(x.x), -fabsf(x.y));
    const float kk = 0.3275911f * 0.70710678118654752440f;
    const float2 d = __ffma2_rn(make_float2(-kk, -kk), nax, make_float2(1.0f, 1.0f));
    const float2 t = make_float2(fast_rcp(d.x), fast_rcp(d.y));
    float2 poly = __ffma2_rn(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t,
                             make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    const float ce = -0.5f * 1.4426950408889634f;
    const float2 ea = __fmul2_rn(__fmul2_rn(x, x), make_float2(ce, ce));
    const float2 e = make_float2(fast_exp2(ea.x), fast_exp2(ea.y));
    const float2 q = __fmul2_rn(__fmul2_rn(poly, t), e);
    const float2 gelu = __ffma2_rn(nax, q, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
    return __fmul2_rn(val, gelu);
}
MOFA_DEVICE float erf_fast(float x) { return 2.0f * gelu_phi(x * 1.41421356237309504880f) - 1.0f; }

}  // namespace mofa
