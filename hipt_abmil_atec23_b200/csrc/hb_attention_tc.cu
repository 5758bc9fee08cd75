// hb_attention_tc.cu — tcgen05 / TMEM attention for the ViT-256 shape (257 tokens, head_dim 64): softmax(q k^T * scale) v of
// Attention.forward (HIPT_4K/vision_transformer.py:119-128) without the [B, heads, 257, 257] matrix.
//
// One persistent CTA per SM walks (sequence, head) items.  An item is THREE query tiles — queries 0..127, 128..255 and a
// tail tile whose only real row is query 256 — and every tile is TWO independent units, one per key half (keys 0..127,
// keys 128..255 plus key 256).  A unit owns 128 TMEM columns and goes through
//     S = Q K_half^T            tcgen05.mma, A and B from shared memory, 128 x 128 fp32 in the unit's 128 columns
//     softmax warpgroup         thread = query row: tcgen05.ld all 128 scores, row max, exp2, row sum; P is rounded to bf16
//                               and written BACK to tensor memory (tcgen05.st) over the unit's columns 0..63
//     O = P V_half              tcgen05.mma with the A operand read from TMEM, V read MN-major in place, into columns 64..127
// with its own (max, sum); the two halves are merged in the tile epilogue (out = (O_a w_a + O_b w_b), w from the two maxima
// and sums), so no unit ever waits for another one's statistics and nothing is rescaled in flight.  Key 256 does not fit a
// 128-key half: its score is a 64-term dot product in the row's thread, it joins half B's statistics, and p_256 v_256 is
// added in the epilogue.  Four units are resident (4 x 128 = all 512 TMEM columns): two softmax warpgroups each alternate
// between their two unit buffers, so the S of a warpgroup's NEXT tile is computed while it is still busy with the current
// one and the MUFU pipe always has a second warp to run.
//   warp 0        TMA producer: K / V [272 x 64] per item (two stages) and one Q tile per warpgroup slot; the tail tile
//                 loads the 16-row box holding query 256 at row 32 * ((item / 2) % 4) of the slot, so that successive tail
//                 tiles of a warpgroup land on different warps (= different scheduler partitions)
//   warps 1, 2    MMA issuers of warpgroups 0 and 1 (whole warp walks the protocol, one elected lane issues)
//   warps 4-7     softmax warpgroup 0 (tiles 0, 2, 4, ... of the CTA's tile sequence), warps 8-11 warpgroup 1
// Softmax statistics stay fp32; P is rounded to bf16 relative to its half's maximum.
#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

#ifdef HB_EXP_TRACE
// timing experiment only (tools/build_exp.sh NAME -DHB_EXP_TRACE, tools/exp_at2_trace.py): clock64 stamps of CTA 0
__device__ long long g_at2_trace[4 * 64 * 16];
#define AT2_TRACE(role, it, k) do { if (blockIdx.x == 0 && (it) < 64 && lane == 0 && (warp < 4 || (warp & 3) == 0)) g_at2_trace[((role) * 64 + (it)) * 16 + (k)] = clock64(); } while (0)
#define AT2_TRACE_UNIT(k) do { if (blockIdx.x == 0 && tr_j < 64 && (threadIdx.x & 127) == 0) g_at2_trace[(tr_w * 64 + tr_j) * 16 + (k)] = clock64(); } while (0)
#else
#define AT2_TRACE(role, it, k) do { } while (0)
#define AT2_TRACE_UNIT(k) do { } while (0)
#endif

constexpr int AT2_THREADS = 384;
constexpr int AT2_S = 257;
constexpr int AT2_KV_BYTES = 272 * 128;                 // [272 keys][64 bf16], 128-byte swizzled rows
constexpr int AT2_TILE_BYTES = 128 * 128;               // one 128-row tile of 128-byte rows
constexpr int AT2_SMEM = 4 * AT2_KV_BYTES + 4 * AT2_TILE_BYTES + 512 + 1024;
constexpr int AT2_TMEM_COLS = 512;

__device__ __forceinline__ void at2_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float at2_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float at2_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float at2_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// One unit of a softmax thread: its row of S (128 scores in the unit's TMEM columns) -> P (bf16, back into columns 0..63).
// `extra` is the score against key 256 (half B) or -inf (half A).  Returns the row's maximum (raw score units) and sum.
__device__ __forceinline__ void at2_softmax_unit(uint32_t t_unit, float extra, float scale_log2, float& m_out, float& l_out,
                                                 float& p_extra, int tr_w = 0, int tr_j = 99, int tr_k = 0) {
    uint32_t sv[128];
#pragma unroll
    for (int c = 0; c < 4; ++c) at2_ld32(t_unit + c * 32, sv + c * 32);
    tmem_ld_wait();
    AT2_TRACE_UNIT(tr_k);
    float m0 = extra, m1 = __uint_as_float(sv[1]), m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
    m0 = fmaxf(m0, __uint_as_float(sv[0]));
#pragma unroll
    for (int j = 4; j < 128; j += 4) {
        m0 = fmaxf(m0, __uint_as_float(sv[j]));     m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(sv[j + 2])); m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
    }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    const float neg_m = -m * scale_log2;
    AT2_TRACE_UNIT(tr_k + 1);
    const f32x2_t c2 = f2_pack(scale_log2, scale_log2), n2 = f2_pack(neg_m, neg_m);
    f32x2_t sum2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {                       // 32 scores -> 16 packed columns per store
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int e = q * 32 + 2 * j;
            float x0, x1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2, n2), x0, x1);
            const float p0 = at2_ex2(x0), p1 = at2_ex2(x1);
            sum2 = f2_add(sum2, f2_pack(p0, p1));
            pk[j] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x16(t_unit + q * 16, pk);
    }
    float s0, s1;
    f2_unpack(sum2, s0, s1);
    AT2_TRACE_UNIT(tr_k + 2);
    p_extra = at2_ex2(fmaf(extra, scale_log2, neg_m));  // exp2(-inf) = 0 for half A
    m_out = m;
    l_out = s0 + s1 + p_extra;
    tmem_st_wait();
}

__global__ void __launch_bounds__(AT2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map16,
                     const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, int n_items, int heads,
                     float scale_log2) {
    extern __shared__ uint8_t smem_raw_at2[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at2) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                                   // [2 stages][34816]
    uint8_t* sV = sK + 2 * AT2_KV_BYTES;                  // [2 stages][34816]
    uint8_t* sQ = sV + 2 * AT2_KV_BYTES;                  // [2 warpgroup slots][16384]
    uint8_t* sO = sQ + 2 * AT2_TILE_BYTES;                // [2 warpgroups][16384]  output staging for the TMA store
    uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * AT2_TILE_BYTES);
    uint64_t* kv_full = bars;            // [2 stages]   TMA bytes of K and V
    uint64_t* kv_empty = bars + 2;       // [2 stages]   3 tiles x (1 MMA commit + 128 softmax threads)
    uint64_t* q_full = bars + 4;         // [2 slots]
    uint64_t* q_empty = bars + 6;        // [2 slots]    1 MMA commit (both S MMAs done) + 128 softmax threads (key-256 dot)
    uint64_t* s_full = bars + 8;         // [4 units]    S of the unit is in TMEM
    uint64_t* p_full = bars + 12;        // [4 units]    128 softmax threads: P is in TMEM
    uint64_t* o_full = bars + 16;        // [4 units]    O of the unit is in TMEM
    uint64_t* u_free = bars + 20;        // [4 units]    128 softmax threads: O has been read, the unit's columns are free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = heads * 64;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 3 * 129);
            mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 129);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); mbar_init(&u_free[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, AT2_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int my_items = 0;
    if (static_cast<int>(blockIdx.x) < n_items) my_items = (n_items - 1 - blockIdx.x) / gridDim.x + 1;
    const int n_tiles = 3 * my_items;                    // tile n = 3 * it + t is handled by warpgroup n & 1

    // register budget: the producer / MMA warpgroup (warps 0-3) gives its registers to the two softmax warpgroups, whose
    // threads hold a full 128-score row (128 x 56 + 256 x 224 = 64,512 of the SM's 65,536 registers); the setmaxnreg sits
    // inside each role's branch so that the compiler sees the budget of that role only
    if (warp == 0) {
        setmaxnreg_dec<56>();
        // ------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int n = 0; n < n_tiles; ++n) {
                const int it = n / 3, t = n - 3 * it;
                const int item = blockIdx.x + it * gridDim.x;
                const int seq = item / heads, h = item - seq * heads;
                const int row0 = seq * AT2_S;
                if (t == 0) {                                           // K and V of the item, stage it & 1
                    const int st = it & 1;
                    mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * AT2_KV_BYTES);
                    uint8_t* k = sK + st * AT2_KV_BYTES;
                    uint8_t* v = sV + st * AT2_KV_BYTES;
                    tma_load_2d(k, &map128, &kv_full[st], D + h * 64, row0);
                    tma_load_2d(k + AT2_TILE_BYTES, &map128, &kv_full[st], D + h * 64, row0 + 128);
                    tma_load_2d(k + 2 * AT2_TILE_BYTES, &map16, &kv_full[st], D + h * 64, row0 + 256);
                    tma_load_2d(v, &map128, &kv_full[st], 2 * D + h * 64, row0);
                    tma_load_2d(v + AT2_TILE_BYTES, &map128, &kv_full[st], 2 * D + h * 64, row0 + 128);
                    tma_load_2d(v + 2 * AT2_TILE_BYTES, &map16, &kv_full[st], 2 * D + h * 64, row0 + 256);
                }
                const int w = n & 1, j = n >> 1;                        // warpgroup slot, index in its tile sequence
                mbar_wait(&q_empty[w], (j & 1) ^ 1);
                uint8_t* q = sQ + w * AT2_TILE_BYTES;
                if (t < 2) {
                    mbar_arrive_expect_tx(&q_full[w], AT2_TILE_BYTES);
                    tma_load_2d(q, &map128, &q_full[w], h * 64, row0 + t * 128);
                } else {                                                // tail tile: query 256 at row 32 * ((it / 2) % 4)
                    mbar_arrive_expect_tx(&q_full[w], 16 * 128);
                    tma_load_2d(q + ((it >> 1) & 3) * 32 * 128, &map16, &q_full[w], h * 64, row0 + 256);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        setmaxnreg_dec<56>();
        const int w = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
        const uint32_t t_a = tmem_base + (2 * w) * 128, t_b = t_a + 128;       // the warpgroup's two unit buffers
        const uint64_t dq = umma_desc_k128(smem_u32(sQ + w * AT2_TILE_BYTES));
        const int my_tiles = (n_tiles - w + 1) >> 1;

        auto issue_s = [&](int j) {                     // S of both halves of the warpgroup's tile j
            const int n = 2 * j + w, it = n / 3, st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            mbar_wait(&q_full[w], j & 1);
            const uint64_t dk = umma_desc_k128(smem_u32(sK + st * AT2_KV_BYTES));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (j > 0) mbar_wait(&u_free[2 * w + half], (j - 1) & 1);
                AT2_TRACE(2 + w, j, 4 + 2 * half);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(half ? t_b : t_a, dq + 2 * kk, dk + half * (AT2_TILE_BYTES >> 4) + 2 * kk, idesc_s, kk != 0);
                    umma_commit(&s_full[2 * w + half]);
                    if (half) umma_commit(&q_empty[w]);
                }
                __syncwarp();
                AT2_TRACE(2 + w, j, 5 + 2 * half);
            }
        };

        if (my_tiles > 0) issue_s(0);
        for (int j = 0; j < my_tiles; ++j) {
            const int n = 2 * j + w, it = n / 3, st = it & 1;
            const uint32_t v_base = smem_u32(sV + st * AT2_KV_BYTES);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&p_full[2 * w + half], j & 1);
                AT2_TRACE(2 + w, j, 2 * half);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t t_u = half ? t_b : t_a;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {                    // 8 x 16 keys: P columns 8 kk .., V rows 128 half + 16 kk ..
                        const uint64_t dv = umma_desc_k128(v_base + (half * 128 + kk * 16) * 128);
                        umma_bf16_ts(t_u + 64, t_u + 8 * kk, dv, idesc_pv, kk != 0);
                    }
                    umma_commit(&o_full[2 * w + half]);
                    if (half) umma_commit(&kv_empty[st]);
                }
                __syncwarp();
                AT2_TRACE(2 + w, j, 1 + 2 * half);
            }
            if (j + 1 < my_tiles) issue_s(j + 1);
        }
    } else if (warp == 3) {
        setmaxnreg_dec<56>();                                 // idle warp of the first warpgroup (the instruction is warpgroup-wide)
#ifdef HB_EXP_TRACE
        // observer: completion times of warpgroup 0's MMAs (s_full / o_full of its two units), role 2 slots 8..11
        if (lane == 0) {
            const int my_tiles0 = (n_tiles + 1) >> 1;
            for (int j = 0; j < my_tiles0; ++j) {
                mbar_wait(&s_full[0], j & 1); AT2_TRACE(2, j, 8);
                mbar_wait(&s_full[1], j & 1); AT2_TRACE(2, j, 9);
                mbar_wait(&o_full[0], j & 1); AT2_TRACE(2, j, 10);
                mbar_wait(&o_full[1], j & 1); AT2_TRACE(2, j, 11);
            }
        }
#endif
    } else {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        setmaxnreg_inc<224>();
        const int w = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // query row inside the tile = TMEM lane
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (2 * w) * 128;
        const int sw = r & 7;
        const bool leader = (quad == 0 && lane == 0);
        const uint8_t* q_row = sQ + w * AT2_TILE_BYTES + r * 128;
        uint8_t* o_row = sO + w * AT2_TILE_BYTES + r * 128;
        const int my_tiles = (n_tiles - w + 1) >> 1;
        for (int j = 0; j < my_tiles; ++j) {
            const int n = 2 * j + w, it = n / 3, t = n - 3 * it, st = it & 1;
            const int item = blockIdx.x + it * gridDim.x;
            const int seq = item / heads, h = item - seq * heads;
            const bool tail = (t == 2);
            const bool active = !tail || quad == ((it >> 1) & 3);       // warp-uniform: the tail tile has one real row

            // ---- score against key 256: q_r . k_256 (row 256 of K is row 0 of its own swizzle atom: unswizzled)
            AT2_TRACE(w, j, 0);
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            mbar_wait(&q_full[w], j & 1);
            AT2_TRACE(w, j, 1);
            float s256 = 0.f;
            if (active) {
                const uint4* k256 = reinterpret_cast<const uint4*>(sK + st * AT2_KV_BYTES + 2 * AT2_TILE_BYTES);
                f32x2_t acc = f2_pack(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 qv = *reinterpret_cast<const uint4*>(q_row + ((c ^ sw) << 4));
                    const uint4 kv = k256[c];
                    acc = f2_fma(f2_pack(at2_lo(qv.x), at2_hi(qv.x)), f2_pack(at2_lo(kv.x), at2_hi(kv.x)), acc);
                    acc = f2_fma(f2_pack(at2_lo(qv.y), at2_hi(qv.y)), f2_pack(at2_lo(kv.y), at2_hi(kv.y)), acc);
                    acc = f2_fma(f2_pack(at2_lo(qv.z), at2_hi(qv.z)), f2_pack(at2_lo(kv.z), at2_hi(kv.z)), acc);
                    acc = f2_fma(f2_pack(at2_lo(qv.w), at2_hi(qv.w)), f2_pack(at2_lo(kv.w), at2_hi(kv.w)), acc);
                }
                float a0, a1;
                f2_unpack(acc, a0, a1);
                s256 = a0 + a1;
            }
            mbar_arrive(&q_empty[w]);

            // ---- the two units: S -> P, each with its own statistics
            float m_a = 0.f, l_a = 1.f, m_b = 0.f, l_b = 1.f, p256 = 0.f, unused;
            mbar_wait(&s_full[2 * w], j & 1);
            AT2_TRACE(w, j, 2);
            tc_fence_after();
#ifdef HB_EXP_TRACE
            if (active) at2_softmax_unit(t_row, -INFINITY, scale_log2, m_a, l_a, unused, w, j, 11);
#else
            if (active) at2_softmax_unit(t_row, -INFINITY, scale_log2, m_a, l_a, unused);
#endif
            tc_fence_before();
            mbar_arrive(&p_full[2 * w]);
            AT2_TRACE(w, j, 3);
            mbar_wait(&s_full[2 * w + 1], j & 1);
            AT2_TRACE(w, j, 4);
            tc_fence_after();
            if (active) at2_softmax_unit(t_row + 128, s256, scale_log2, m_b, l_b, p256);
            tc_fence_before();
            mbar_arrive(&p_full[2 * w + 1]);
            AT2_TRACE(w, j, 5);

            // ---- merge weights of the two halves
            const float m = fmaxf(m_a, m_b);
            const float e_a = at2_ex2((m_a - m) * scale_log2), e_b = at2_ex2((m_b - m) * scale_log2);
            const float inv = 1.0f / fmaf(l_a, e_a, l_b * e_b);
            const float w_a = e_a * inv, w_b = e_b * inv, w_256 = p256 * w_b;

            // ---- epilogue: O_a w_a + O_b w_b + p_256 w_b v_256 (O_b in two halves of 32 columns to bound the registers)
            uint32_t oa[64];
            uint4 res[8];
            mbar_wait(&o_full[2 * w], j & 1);
            AT2_TRACE(w, j, 6);
            tc_fence_after();
            if (active) { at2_ld32(t_row + 64, oa); at2_ld32(t_row + 96, oa + 32); tmem_ld_wait(); }
            tc_fence_before();
            mbar_arrive(&u_free[2 * w]);
            AT2_TRACE(w, j, 7);
            mbar_wait(&o_full[2 * w + 1], j & 1);
            AT2_TRACE(w, j, 8);
            tc_fence_after();
            if (active) {
                const uint4* v256 = reinterpret_cast<const uint4*>(sV + st * AT2_KV_BYTES + 2 * AT2_TILE_BYTES);
                const f32x2_t wa2 = f2_pack(w_a, w_a), wb2 = f2_pack(w_b, w_b), w2 = f2_pack(w_256, w_256);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t ob[32];
                    at2_ld32(t_row + 128 + 64 + hf * 32, ob);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 vv = v256[hf * 4 + q];
                        const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
                        uint32_t o4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 8 * q + 2 * e;
                            f32x2_t acc = f2_mul(f2_pack(__uint_as_float(oa[hf * 32 + c]), __uint_as_float(oa[hf * 32 + c + 1])), wa2);
                            acc = f2_fma(f2_pack(__uint_as_float(ob[c]), __uint_as_float(ob[c + 1])), wb2, acc);
                            acc = f2_fma(f2_pack(at2_lo(vw[e]), at2_hi(vw[e])), w2, acc);
                            float x0, x1;
                            f2_unpack(acc, x0, x1);
                            o4[e] = pack_bf16x2(x0, x1);
                        }
                        res[hf * 4 + q] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&u_free[2 * w + 1]);
            AT2_TRACE(w, j, 9);
            mbar_arrive(&kv_empty[st]);

            if (!tail) {
                // the TMA store of the warpgroup's previous full tile must have drained the staging buffer
                if (leader) tma_store_wait_read<0>();
                named_bar_sync(4 + w, 128);
#pragma unroll
                for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(o_row + ((q ^ sw) << 4)) = res[q];
                fence_proxy_async_smem();
                named_bar_sync(4 + w, 128);
                AT2_TRACE(w, j, 10);
                if (leader) {
                    tma_store_2d(&map_out, sO + w * AT2_TILE_BYTES, h * 64, seq * AT2_S + t * 128);
                    tma_store_commit();
                }
            } else if (active && lane == 0) {
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(seq) * AT2_S + 256) * D + h * 64);
#pragma unroll
                for (int q = 0; q < 8; ++q) dst[q] = res[q];
            }
        }
        if (leader) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, AT2_TMEM_COLS);
}

int attention_tc2_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int heads, float scale, cudaStream_t stream) {
    const int D = heads * 64;
    const uint64_t rows = static_cast<uint64_t>(n_seq) * AT2_S;
    CUtensorMap map128, map16, map_out;
    if (encode_tmap_2d(&map128, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 128, 64)) return -1;
    if (encode_tmap_2d(&map16, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 16, 64)) return -1;
    if (encode_tmap_2d(&map_out, TMAP_BF16, out_bf16, rows, D, static_cast<uint64_t>(D) * 2, 128, 64)) return -1;
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(attention_tc2_kernel), AT2_SMEM)) return -1;
    const int n_items = n_seq * heads;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_tc2_kernel<<<grid, AT2_THREADS, AT2_SMEM, stream>>>(map128, map16, map_out, static_cast<__nv_bfloat16*>(out_bf16),
                                                                  n_items, heads, scale * 1.4426950408889634f);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

#ifdef HB_EXP_TRACE
extern "C" int hb_exp_read_at2_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_at2_trace, sizeof(long long) * 4 * 64 * 16) == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace hb
