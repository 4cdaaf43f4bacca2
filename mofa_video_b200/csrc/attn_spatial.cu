// Spatial self-attention (head_dim 64), two 128-query tiles per CTA.
//
// With head_dim 64 the kernel is bound by the exponential (MUFU.EX2: one warp instruction per 8 clk per SM
// sub-partition, measured by tools/microbench/mufu_rate.cu -> 2048 clk per 256 x 128 score block) and by the latency
// chain around it, not by the tensor core (1024 clk for the same block).  Structure:
//   * one query row per thread; the 128 scores of a tile are read from TMEM ONCE into registers and the S buffer is
//     released right after that read (s_free), so S_{j+1} = Q K_{j+1}^T is computed while the exponentials of tile j
//     run -- the registers act as the second S buffer;
//   * P never touches shared memory: the packed fp16 probabilities are written back to tensor memory (tcgen05.st)
//     and consumed as the A operand of the P V MMA (tcgen05.mma with A in TMEM); V is the MN-major B operand, read in
//     place from the TMA tile.  Shared-memory traffic per 256 x 128 block drops from 256 KB to 128 KB;
//   * O accumulates in TMEM across KV tiles, no per-tile read-back; the running maximum is lazy (moved only when it
//     grows by more than 2^8, P <= 256 stays inside fp16), so the O rescale -- a TMEM load / scale / store by the
//     owning warp -- happens on the first one or two tiles only;
//   * the MMA-issuing thread multiplexes four barriers (s_free / p_full of both query tiles) with non-blocking probes
//     and issues whatever is ready;
//   * the two softmax warps that share an SM sub-partition (one per query tile) hand the MUFU over explicitly
//     (named barriers): their exp phases alternate instead of drifting into lockstep (+8 %);
//   * masking only on the ragged last tile, 3-input max, four independent max / sum chains.
// TMEM: S 2 x 128 | O 2 x 64 | P 2 x 64 columns = 512.
// Tried and measured slower on B200 (profiles/r1_attention_experiments.md): exp2 on the FMA pipe for 1/4..1/2 of the
// scores (issue-bound), two threads per row (16 softmax warps), row sums through 16 extra all-ones B columns,
// softmax warpgroups issuing their own MMAs.
//
//   warp 0 (1 lane) : TMA producer   Q0,Q1 once; K/V tiles through a 5-stage ring
//   warp 1 (1 lane) : MMA issuer     S_w = Q_w K^T, O_w += P_w V
//   warps 2..5      : softmax warpgroup 0 (query tile 0), one row per thread
//   warps 6..9      : softmax warpgroup 1 (query tile 1)
#include "../../include/mofa_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace mofa {

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

namespace v2 {

constexpr int kKVStages = 5;
constexpr uint32_t kTile = 128 * 64 * 2;  // 16 KB
constexpr uint32_t kPBytes = 2 * kTile;   // one 128 x 128 fp16 P tile

struct Params {
    __half* out;
    int L, C, heads, n_kv;
    float scale_log2;
    int idle_ns;  // MMA issuer: nanosleep between barrier probes when nothing was ready (0 = spin)
};

MOFA_DEVICE float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// kPolyEvery: every kPolyEvery-th pair of probabilities is computed by poly_exp2 on the FMA pipe instead of MUFU.EX2
// (0 = none); kHandoff: alternate the exp phases of the two warps that share an SM sub-partition.
// kSplit: threads per query row (1: 8 softmax warps, 128 scores per thread; 2: 16 softmax warps, 64 scores per thread,
// the two halves of a row exchange their maxima / sums through shared memory).
template <int kPolyEvery, bool kHandoff, int kSplit>
__global__ void __launch_bounds__(64 + 256 * kSplit, 1)
attn_spatial2_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                              // 2 x 16 KB
    uint8_t* sKV = sQ + 2 * kTile;                   // kKVStages x (K | V)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kKVStages * 2 * kTile);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + kKVStages;
    uint64_t* s_full = kv_empty + kKVStages;  // [2] per query tile: S_j landed in TMEM
    uint64_t* s_free = s_full + 2;            // [2] S_j copied to registers, the buffer may be overwritten
    uint64_t* p_full = s_free + 2;            // [2] P_j in shared memory (and O rescaled if needed)
    uint64_t* o_full = p_full + 2;            // [2] P_j V_j accumulated
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);
    float* fence_slots = reinterpret_cast<float*>(tmem_ptr_smem + 4);  // one word per thread (see the MUFU hand-off)
    float* xchg = fence_slots + 64 + 256 * 2;  // [2 slots][2 tiles][2 halves][128 rows] (kSplit == 2)
    static_assert(kSplit == 1 || kSplit == 2, "kSplit");
    static_assert(!(kHandoff && kSplit == 2), "the hand-off pairs the two warps of a sub-partition: kSplit == 1 only");
    constexpr int kCols = 128 / kSplit;  // scores per thread and KV tile
    constexpr int kOc = 64 / kSplit;     // O columns per thread

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256;
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int n_kv = p.n_kv;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < kKVStages; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128 * kSplit);
            mbar_init(&p_full[i], 128 * kSplit);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;  // 2 x 64 columns
    const uint32_t tmem_P = tmem_base + 384;  // 2 x 64 columns: P as packed fp16 pairs, the A operand of P V

    if (threadIdx.x == 0) {
        // ===================== TMA producer =====================
        mbar_arrive_expect_tx(q_full, 2 * kTile);
        tma_load_3d(&tmQKV, q_full, sQ, head * 64, q0, frame);
        tma_load_3d(&tmQKV, q_full, sQ + kTile, head * 64, q0 + 128, frame);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            uint8_t* sk = sKV + stage * 2 * kTile;
            mbar_arrive_expect_tx(&kv_full[stage], 2 * kTile);
            tma_load_3d(&tmQKV, &kv_full[stage], sk, p.C + head * 64, j * 128, frame);
            tma_load_3d(&tmQKV, &kv_full[stage], sk + kTile, 2 * p.C + head * 64, j * 128, frame);
            if (++stage == kKVStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the schedule (every decision is a warp vote, so it stays warp-uniform) and one elected
        // lane issues: descriptors and loop state live in uniform registers.  Inside a single-thread branch the compiler
        // re-elects a lane and broadcasts every operand before each tcgen05 instruction -- ~16 instructions per MMA on a
        // scheduler shared with softmax warps, more than the 32..64 cycles one of these MMAs occupies the tensor pipe.
        const bool leader = elect_one();
        const uint32_t idesc_s = umma_idesc_f16(128, false);
        const uint32_t idesc_o = umma_idesc_f16(64, true);
        const uint64_t dq0 = umma_desc_sw128_kmajor(smem_u32(sQ));
        const uint64_t dq1 = umma_desc_sw128_kmajor(smem_u32(sQ + kTile));
        const uint64_t dk0 = umma_desc_sw128_kmajor(smem_u32(sKV));
        const uint64_t dv0 = umma_desc_sw128_mnmajor(smem_u32(sKV + kTile), kTile);
        constexpr uint64_t kStageStep = (2 * kTile) >> 4;   // descriptor address units per K/V ring stage
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        if (leader) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const uint64_t dq = w ? dq1 : dq0;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S + w * 128, dq + 2 * k, dk0 + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[w]);
            }
        }
        // per query tile: next S tile to issue (with its ring stage / phase) and next PV tile to issue
        int js[2] = {1, 1}, js_stage[2] = {1 % kKVStages, 1 % kKVStages};
        uint32_t js_phase[2] = {(1 / kKVStages) & 1u, (1 / kKVStages) & 1u};
        int jp[2] = {0, 0}, jp_stage[2] = {0, 0};
        int released = 0, rel_stage = 0;
        while (jp[0] < n_kv || jp[1] < n_kv) {
            bool progress = false;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const uint64_t dq = w ? dq1 : dq0;
                if (js[w] < n_kv && __all_sync(0xffffffffu, mbar_test(&s_free[w], (js[w] - 1) & 1) &&
                                                                mbar_test(&kv_full[js_stage[w]], js_phase[w]))) {
                    tc_fence_after();
                    if (leader) {
                        const uint64_t dk = dk0 + static_cast<uint64_t>(js_stage[w]) * kStageStep;
                        umma_f16_ss(tmem_S + w * 128, dq, dk, idesc_s, 0u);
                        umma_f16_ss(tmem_S + w * 128, dq + 2, dk + 2, idesc_s, 1u);
                        umma_f16_ss(tmem_S + w * 128, dq + 4, dk + 4, idesc_s, 1u);
                        umma_f16_ss(tmem_S + w * 128, dq + 6, dk + 6, idesc_s, 1u);
                        umma_commit(&s_full[w]);
                    }
                    ++js[w];
                    if (++js_stage[w] == kKVStages) {
                        js_stage[w] = 0;
                        js_phase[w] ^= 1u;
                    }
                    progress = true;
                }
                if (jp[w] < n_kv && __all_sync(0xffffffffu, mbar_test(&p_full[w], jp[w] & 1))) {
                    tc_fence_after();
                    ++jp[w];
                    const int done = jp[0] < jp[1] ? jp[0] : jp[1];
                    if (leader) {
                        const uint64_t dv = dv0 + static_cast<uint64_t>(jp_stage[w]) * kStageStep;
                        const uint32_t acc0 = jp[w] > 1 ? 1u : 0u;
                        umma_f16_ts(tmem_O + w * 64, tmem_P + w * 64, dv, idesc_o, acc0);
#pragma unroll
                        for (int kk = 1; kk < 8; ++kk)  // 16 keys per step = 8 TMEM columns of packed P
                            umma_f16_ts(tmem_O + w * 64, tmem_P + w * 64 + kk * 8, dv + 128 * kk, idesc_o, 1u);
                        umma_commit(&o_full[w]);
                        // a K/V stage is free once both query tiles have issued its PV (the commit covers every earlier MMA)
                        if (done > released) umma_commit(&kv_empty[rel_stage]);
                    }
                    if (++jp_stage[w] == kKVStages) jp_stage[w] = 0;
                    if (done > released) {
                        ++released;
                        if (++rel_stage == kKVStages) rel_stage = 0;
                    }
                    progress = true;
                }
            }
            if (!progress && p.idle_ns > 0) __nanosleep(p.idle_ns);
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ===================== softmax warps =====================
        const int sw = warp - 2;
        const int w = sw / (4 * kSplit);   // query tile
        const int half = (sw >> 2) % kSplit;  // which kCols-wide slice of the score row
        const int qd = warp & 3;           // TMEM lane quarter this warp may access
        const int r = qd * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ts = tmem_S + lane_addr + w * 128 + half * kCols;
        const uint32_t to = tmem_O + lane_addr + w * 64 + half * kOc;
        const uint32_t tp = tmem_P + lane_addr + w * 64 + half * (kCols / 2);
        const float sl2 = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        const uint32_t fence_slot = smem_u32(fence_slots + threadIdx.x);
        // MUFU hand-off between the two warps that share an SM sub-partition (this warp and warp +-4): their exp
        // phases strictly alternate, so one streams ex2 at the pipe's full rate while the other loads S / finds the
        // maximum / stores P.  Left alone they fall into lockstep and the MUFU idles ~40 % of the time.
        const int bar_mine = 1 + 2 * qd + w, bar_other = 1 + 2 * qd + (1 - w);
        if (kHandoff && w == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
        // kSplit == 2: the two warps holding the halves of the same rows meet at named barrier 1 + 4w + qd
        const int bar_pair = 1 + 4 * w + qd;
        const uint32_t x_mine = smem_u32(xchg + (w * 2 + half) * 128 + r);
        const uint32_t x_peer = smem_u32(xchg + (w * 2 + (1 - half)) * 128 + r);

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[w], j & 1);
            tc_fence_after();
            uint32_t s[kCols];
#pragma unroll
            for (int c = 0; c < kCols / 32; ++c)
                tmem_ld_32x32(ts + c * 32, reinterpret_cast<uint32_t(&)[32]>(s[c * 32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_free[w]);  // S_{j+1} may now overwrite the TMEM buffer
            const int kv_valid = p.L - j * 128 - half * kCols;
            if (kv_valid < kCols) {  // warp-uniform: ragged last tile
#pragma unroll
                for (int i = 0; i < kCols; ++i)
                    if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: FMNMX3 latency, not issue
#pragma unroll
            for (int i = 0; i < kCols; i += 8) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    mx4[q] = max3(mx4[q], __uint_as_float(s[i + 2 * q]), __uint_as_float(s[i + 2 * q + 1]));
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            if (kSplit == 2) {
                // both halves must use the same reference maximum; slots alternate with j so a fast warp cannot
                // overwrite a value its peer has not read yet
                const uint32_t slot = static_cast<uint32_t>(j & 1) * (2 * 2 * 128 * 4);
                float other;
                asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(x_mine + slot), "f"(mx) : "memory");
                asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");
                asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(other) : "r"(x_peer + slot) : "memory");
                mx = fmaxf(mx, other);
            }
            mx *= sl2;
            // lazy rescale: move the reference maximum only when it grows by more than 2^8
            float alpha = 1.0f;
            if (mx > m + 8.0f) {
                alpha = fast_exp2(m - mx);  // first tile: m = -inf -> 0
                m = mx;
            }
            l *= alpha;
            // exponentials first (registers only): they overlap P_{j-1} V_{j-1}, which still reads the P buffer
            float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
            // kPolyEvery < 0: row sums of the ROUNDED probabilities with packed-half adds (one HADD2 per two scores
            // instead of two FADD); four accumulators of <= 16 values <= 256 each, widened to fp32 once per tile
            __half2 hs[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hs[q] = __floats2half2_rn(0.f, 0.f);
            uint32_t packed[kCols / 2];
            // the reference maximum takes a round trip through shared memory across the barrier (and the row sum is
            // stored before the arrive below): ptxas keeps shared-memory accesses ordered around BAR, so the
            // exponentials can neither be hoisted above the hand-off nor sunk below it
            float m_x = m;
            if (kHandoff) {
                asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(fence_slot), "f"(m) : "memory");
                asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");
                asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(m_x) : "r"(fence_slot) : "memory");
            }
            // kPolyEvery: -1 packed-half row sums, no polynomial; 0 / 2 / 4 fp32 sums (and 1/2, 1/4 of the exponentials on
            // the FMA pipe); 12 / 13 / 14 packed-half sums AND every 2nd / 3rd / 4th PAIR of exponentials on the packed
            // fp32 FMA pipe (poly_exp2x2).  The score scaling is one FFMA2 per pair in every variant.
            constexpr bool kHalfSum = kPolyEvery < 0 || kPolyEvery >= 10;
            constexpr int kPE = kPolyEvery >= 10 ? kPolyEvery - 10 : (kPolyEvery > 0 ? kPolyEvery : 0);
            const float2 sl2v = make_float2(sl2, sl2), nmv = make_float2(-m_x, -m_x);
#pragma unroll
            for (int i = 0; i < kCols / 2; ++i) {
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sl2v, nmv);
                const bool poly = kPE > 0 && (i % (kPE > 0 ? kPE : 1)) == 0;
                float p0, p1;
                if (poly) {
                    if (kPolyEvery >= 10) {
                        const float2 pp = poly_exp2x2(x);
                        p0 = pp.x;
                        p1 = pp.y;
                    } else {
                        p0 = poly_exp2(x.x);
                        p1 = poly_exp2(x.y);
                    }
                } else {
                    p0 = fast_exp2(x.x);
                    p1 = fast_exp2(x.y);
                }
                const __half2 h = __floats2half2_rn(p0, p1);
                if (kHalfSum) {
                    hs[i & 3] = __hadd2(hs[i & 3], h);
                } else if (i & 1) {
                    sum2 += p0;
                    sum3 += p1;
                } else {
                    sum0 += p0;
                    sum1 += p1;
                }
                packed[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            if (kHalfSum) {
                const float2 a = __half22float2(hs[0]), b = __half22float2(hs[1]);
                const float2 c = __half22float2(hs[2]), d = __half22float2(hs[3]);
                sum0 = a.x + a.y;
                sum1 = b.x + b.y;
                sum2 = c.x + c.y;
                sum3 = d.x + d.y;
            }
            const float l_tile = (sum0 + sum1) + (sum2 + sum3);
            if (kHandoff) {
                asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(fence_slot), "f"(l_tile) : "memory");
                if (w == 0 || j + 1 < n_kv) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
            }
            if (j > 0) {
                // P_{j-1} V_{j-1} done: the P buffer may be overwritten and O is stable
                mbar_wait(&o_full[w], (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int c = 0; c < kOc / 8; ++c) {
                        uint32_t v[8];
                        tmem_ld_32x8(to + c * 8, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x8(to + c * 8, v);
                    }
                    tmem_st_wait();
                }
            }
#pragma unroll
            for (int c = 0; c < kCols / 64; ++c)
                tmem_st_32x32(tp + c * 32, reinterpret_cast<const uint32_t(&)[32]>(packed[c * 32]));
            tmem_st_wait();
            l += l_tile;
            tc_fence_before();
            mbar_arrive(&p_full[w]);
        }
        if (kSplit == 2) {  // total row sum = sum of the two halves (both are relative to the same maximum)
            const uint32_t slot = static_cast<uint32_t>(n_kv & 1) * (2 * 2 * 128 * 4);
            float other;
            asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(x_mine + slot), "f"(l) : "memory");
            asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");
            asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(other) : "r"(x_peer + slot) : "memory");
            l += other;
        }
        mbar_wait(&o_full[w], (n_kv - 1) & 1);
        tc_fence_after();
        const int qrow = q0 + w * 128 + r;
        const float inv = 1.0f / l;
        __half* dst = p.out + (static_cast<long long>(frame) * p.L + qrow) * p.C + head * 64 + half * kOc;
#pragma unroll
        for (int c = 0; c < kOc / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(to + c * 32, v);
            tmem_ld_wait();
            if (qrow < p.L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t wv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2 h = __floats2half2_rn(__uint_as_float(v[g * 8 + 2 * i]) * inv,
                                                            __uint_as_float(v[g * 8 + 2 * i + 1]) * inv);
                        wv[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace v2
}  // namespace mofa

using namespace mofa;

extern "C" int mofa_attn_spatial(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                                 mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !out || frames <= 0 || L <= 0 || heads <= 0) {
        set_last_error("mofa_attn_spatial: bad arguments");
        return MOFA_ERR_ARG;
    }
    const int C = heads * 64;
    CUtensorMap tm;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(L), static_cast<uint64_t>(frames)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * C) * 2, static_cast<uint64_t>(L) * 3 * C * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_f16(&tm, qkv, 3, dims, strides, box);
    if (rc) return rc;
    v2::Params p;
    p.out = static_cast<__half*>(out);
    p.L = L;
    p.C = C;
    p.heads = heads;
    p.n_kv = (L + 127) / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    static int idle_ns = -1;
    if (idle_ns < 0) {
        const char* e = getenv("MOFA_ATTN_IDLE_NS");
        idle_ns = e ? atoi(e) : 20;
    }
    p.idle_ns = idle_ns;
    const size_t smem_bytes = 2 * v2::kTile + v2::kKVStages * 2 * v2::kTile + 16 * 8 + 16 +
                              (64 + 256 * 2) * 4 + 2 * 2 * 2 * 128 * 4 + 1024;
    // variants (environment, read once): MOFA_ATTN_SPLIT = 1 | 2 threads per query row, MOFA_ATTN_HANDOFF = 0 | 1
    // (kSplit 1 only), MOFA_ATTN_POLY = -1 | 0 | 2 | 4 | 12 | 13 | 14 (-1: packed-half row sums; 0: fp32 row sums; 2 / 4: fp32
    // sums and 1/2 / 1/4 of the exponentials on the FMA pipe; 12 / 13 / 14: packed-half sums and every 2nd / 3rd / 4th pair
    // of exponentials as a packed-fp32 polynomial).
    // Defaults are the measured best on B200 (profiles/).
    using Kern = void (*)(const CUtensorMap, const v2::Params);
    static Kern kern = nullptr;
    static int threads = 0;
    if (!kern) {
        const char* ep = getenv("MOFA_ATTN_POLY");
        const char* eh = getenv("MOFA_ATTN_HANDOFF");
        const char* es = getenv("MOFA_ATTN_SPLIT");
        const int poly = ep ? atoi(ep) : 14;  // packed-half row sums, every 3rd pair of exponentials on the packed FMA pipe
                                               // (same-box A/B at L = 9216: -1 734, 12 746, 13 770, 14 760 TFLOP/s)
        const int split = es ? atoi(es) : 1;
        // (the MUFU hand-off paid while every exponential was a MUFU op: 734 vs ~700; with a third of them on the FMA pipe
        //  the pipe is no longer saturated and the barrier only serialises: 771 with, 813 without)
        const bool handoff = eh ? (eh[0] == '1') : false;
        Kern k;
        if (split == 2) {
            k = poly == 2    ? v2::attn_spatial2_kernel<2, false, 2>
                : poly == 4  ? v2::attn_spatial2_kernel<4, false, 2>
                : poly == 13 ? v2::attn_spatial2_kernel<13, false, 2>
                : poly == 14 ? v2::attn_spatial2_kernel<14, false, 2>
                : poly == -1 ? v2::attn_spatial2_kernel<-1, false, 2>
                             : v2::attn_spatial2_kernel<0, false, 2>;
        } else if (handoff) {
            k = poly == 2    ? v2::attn_spatial2_kernel<2, true, 1>
                : poly == 4  ? v2::attn_spatial2_kernel<4, true, 1>
                : poly == -1 ? v2::attn_spatial2_kernel<-1, true, 1>
                : poly == 12 ? v2::attn_spatial2_kernel<12, true, 1>
                : poly == 13 ? v2::attn_spatial2_kernel<13, true, 1>
                : poly == 14 ? v2::attn_spatial2_kernel<14, true, 1>
                             : v2::attn_spatial2_kernel<0, true, 1>;
        } else {
            k = poly == 2    ? v2::attn_spatial2_kernel<2, false, 1>
                : poly == 4  ? v2::attn_spatial2_kernel<4, false, 1>
                : poly == 12 ? v2::attn_spatial2_kernel<12, false, 1>
                : poly == 13 ? v2::attn_spatial2_kernel<13, false, 1>
                : poly == 14 ? v2::attn_spatial2_kernel<14, false, 1>
                : poly == 15 ? v2::attn_spatial2_kernel<15, false, 1>
                : poly == 16 ? v2::attn_spatial2_kernel<16, false, 1>
                : poly == -1 ? v2::attn_spatial2_kernel<-1, false, 1>
                             : v2::attn_spatial2_kernel<0, false, 1>;
        }
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_bytes));
        if (e != cudaSuccess) {
            set_last_error("mofa_attn_spatial: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        kern = k;
        threads = 64 + 256 * (split == 2 ? 2 : 1);
    }
    dim3 grid((L + 255) / 256, heads, frames);
    kern<<<grid, threads, smem_bytes, stream>>>(tm, p);
    return check_launch("mofa_attn_spatial");
}
