// hb_attention_tc.cu — tcgen05 / TMEM attention for the ViT-256 shape (257 tokens, head_dim 64): softmax(q k^T * scale) v of
// Attention.forward (HIPT_4K/vision_transformer.py:119-128) without the [B, heads, 257, 257] matrix.
//
// One persistent CTA per SM walks (sequence, head) items.  An item is two 128-query tiles plus query 256; every tile is TWO
// independent units, one per key half (keys 0..127, keys 128..255 plus key 256).  A unit owns 128 TMEM columns and goes through
//     S = Q K_half^T            tcgen05.mma, A and B from shared memory, 128 x 128 fp32 in the unit's 128 columns
//     softmax warpgroup         thread = query row: tcgen05.ld all 128 scores, row max, exp2, row sum; P is rounded to bf16
//                               and written BACK to tensor memory (tcgen05.st) over the unit's columns 0..63
//     O = P V_half              tcgen05.mma with the A operand read from TMEM, V read MN-major in place, into columns 64..127
// with its own (max, sum).  The two halves are merged by a separate EPILOGUE warpgroup (out = O_a w_a + O_b w_b, weights from
// the two maxima and sums), so a softmax warpgroup goes from one unit straight to the next — it never waits for a P V product,
// nothing is rescaled in flight — and the S of its next tile is issued as soon as the epilogue has drained the unit.
// Key 256 does not fit a 128-key half: its score against every query is a 64-term dot product on the CUDA cores, it joins
// half B's statistics, and p_256 v_256 is added in the epilogue.  Query 256 has no tile either: one row on the CUDA cores
// (two keys per thread, warp-shuffle softmax, a 4-way split P V).  Both run in the EPILOGUE warpgroup at the START of an
// item, while the softmax warpgroups are busy with its first units: a few hundred issue slots per scheduler partition that
// are off the softmax critical path, and K / V are released as soon as the item's last P V product has completed.
//   warp 0        TMA producer: K / V [272 x 64] per item (two stages), one Q tile per softmax warpgroup, the 16-row box that
//                 starts at query 256
//   warps 1, 2    MMA issuers of softmax warpgroups 0 and 1 (whole warp walks the protocol, one elected lane issues)
//   warps 4-7     softmax warpgroup 0 (queries 0..127 of every item), warps 8-11 warpgroup 1 (queries 128..255)
//   warps 12-15   epilogue warpgroup: drains O_a / O_b of both tiles, merges, stores rows straight to global memory; query 256
// Four units are resident (4 x 128 = all 512 TMEM columns).  Softmax statistics stay fp32; P is rounded to bf16 relative to
// its half's maximum.
#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

#ifdef HB_EXP_TRACE
// timing experiment only (tools/build_exp.sh NAME -DHB_EXP_TRACE, tools/exp_at2_trace.py): clock64 stamps of CTA 0
__device__ long long g_at2_trace[4 * 64 * 16];
#define AT2_TRACE(role, it, k) do { if (blockIdx.x == 0 && (it) < 64 && (threadIdx.x & 127) == 0) g_at2_trace[((role) * 64 + (it)) * 16 + (k)] = clock64(); } while (0)
#else
#define AT2_TRACE(role, it, k) do { } while (0)
#endif

constexpr int AT2_THREADS = 512;
constexpr int AT2_S = 257;
constexpr int AT2_KV_BYTES = 272 * 128;                 // [272 keys][64 bf16], 128-byte swizzled rows
constexpr int AT2_TILE_BYTES = 128 * 128;               // one 128-row tile of 128-byte rows
constexpr int AT2_Q256_BYTES = 16 * 128;                // the 16-row box that starts at query 256 (row 0 is unswizzled)
constexpr int AT2_STATS_FLOATS = 128 * 8;               // per warpgroup: [128 rows][m_a, l_a, m_b, l_b, p_256, pad]
constexpr int AT2_ROW_FLOATS = 272 + 4 * 64 + 16 + 64;   // query 256: p[272], partial O [4][64], reduction scratch;
                                                                   // copies of v_256 (2 stages x 64 bf16)
constexpr int AT2_SMEM = 4 * AT2_KV_BYTES + 2 * AT2_TILE_BYTES + 2 * AT2_Q256_BYTES + 2 * AT2_STATS_FLOATS * 4 +
                         AT2_ROW_FLOATS * 4 + 512 + 1024;
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

// ---- mma.sync helpers for the two odd ones out (key 256 and query 256), which have no 128-wide tile of their own
__device__ __forceinline__ uint32_t at2_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void at2_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void at2_ldsm4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void at2_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// q_r . k_256 for the 32 rows [row0, row0 + 32) of a swizzled Q tile, one warp: two m16 tiles x four k16 steps against a
// B fragment whose only non-zero column is k_256 (an unswizzled 128-byte row).  Lane l returns the score of row row0 + l.
__device__ __forceinline__ float at2_warp_dot_k256(uint32_t q_tile, int row0, const uint32_t* k256_words, int lane) {
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint32_t b0 = lane < 4 ? k256_words[kk * 8 + lane] : 0u;          // B[k = 2 (l % 4) .., n = l / 4]: column 0 only
        const uint32_t b1 = lane < 4 ? k256_words[kk * 8 + 4 + lane] : 0u;
        uint32_t a[4], a2[4];
        at2_ldsm4(q_tile + at2_swz(row0 + (lane & 15), kk * 2 + (lane >> 4)), a[0], a[1], a[2], a[3]);
        at2_ldsm4(q_tile + at2_swz(row0 + 16 + (lane & 15), kk * 2 + (lane >> 4)), a2[0], a2[1], a2[2], a2[3]);
        at2_mma16816(c0, a, b0, b1);
        at2_mma16816(c1, a2, b0, b1);
    }
    // column 0 of the C fragments lives in lanes 4 i: rows i (c[0]) and i + 8 (c[2]) of each m16 tile
    const int src = (lane & 7) * 4;
    const float v00 = __shfl_sync(0xffffffffu, c0[0], src), v01 = __shfl_sync(0xffffffffu, c0[2], src);
    const float v10 = __shfl_sync(0xffffffffu, c1[0], src), v11 = __shfl_sync(0xffffffffu, c1[2], src);
    return (lane & 16) ? ((lane & 8) ? v11 : v10) : ((lane & 8) ? v01 : v00);
}

// One unit of a softmax thread: its row of S (128 scores in the unit's TMEM columns) -> P (bf16, back into columns 0..63).
// `extra` is the score against key 256 (half B) or -inf (half A).  Returns the row's maximum (raw score units) and sum.
__device__ __forceinline__ void at2_softmax_unit(uint32_t t_unit, float extra, float scale_log2, float& m_out, float& l_out,
                                                 float& p_extra) {
    uint32_t sv[128];
#pragma unroll
    for (int c = 0; c < 4; ++c) at2_ld32(t_unit + c * 32, sv + c * 32);
    tmem_ld_wait();
    float m0 = extra, m1 = __uint_as_float(sv[1]), m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
    m0 = fmaxf(m0, __uint_as_float(sv[0]));
#pragma unroll
    for (int j = 4; j < 128; j += 4) {
        m0 = fmaxf(m0, __uint_as_float(sv[j]));     m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(sv[j + 2])); m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
    }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    const float neg_m = -m * scale_log2;
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
    p_extra = at2_ex2(fmaf(extra, scale_log2, neg_m));  // exp2(-inf) = 0 for half A
    m_out = m;
    l_out = s0 + s1 + p_extra;
    tmem_st_wait();
}

__global__ void __launch_bounds__(AT2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map16,
                     __nv_bfloat16* __restrict__ out, int n_items, int heads, float scale_log2) {
    extern __shared__ uint8_t smem_raw_at2[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at2) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                                   // [2 stages][34816]
    uint8_t* sV = sK + 2 * AT2_KV_BYTES;                  // [2 stages][34816]
    uint8_t* sQ = sV + 2 * AT2_KV_BYTES;                  // [2 softmax warpgroups][16384]
    uint8_t* sQ256 = sQ + 2 * AT2_TILE_BYTES;             // [2 stages][2048]
    float* sStats = reinterpret_cast<float*>(sQ256 + 2 * AT2_Q256_BYTES);       // [2 warpgroups][128][8]
    float* sRow = sStats + 2 * AT2_STATS_FLOATS;          // query 256 scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRow + AT2_ROW_FLOATS);
    uint64_t* kv_full = bars;            // [2 stages]   TMA bytes of K, V and the query-256 box
    uint64_t* kv_empty = bars + 2;       // [2 stages]   2 MMA commits (the item's last P V) + 256 softmax threads (key 256) + 128 epilogue threads (query 256)
    uint64_t* q_full = bars + 4;         // [2 warpgroups]
    uint64_t* q_empty = bars + 6;        // [2]          1 MMA commit (both S MMAs done) + 128 softmax threads (key-256 dot)
    uint64_t* s_full = bars + 8;         // [4 units]    S of the unit is in TMEM
    uint64_t* p_full = bars + 12;        // [4 units]    128 softmax threads: P is in TMEM
    uint64_t* o_full = bars + 16;        // [4 units]    O of the unit is in TMEM
    uint64_t* u_free = bars + 20;        // [4 units]    128 epilogue threads: O has been read, the unit's columns are free
    uint64_t* st_full = bars + 24;       // [2 warpgroups]  128 softmax threads: the tile's row statistics are in shared memory
    uint64_t* st_empty = bars + 26;      // [2]          128 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
    uint4* sV256 = reinterpret_cast<uint4*>(sRow + 272 + 4 * 64 + 16);   // [2 stages][8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = heads * 64;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2 + 256 + 128);
            mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 129);
            mbar_init(&st_full[i], 128); mbar_init(&st_empty[i], 128);
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

    // register budget (64 K registers per SM): 128 x 48 (producer / MMA warpgroup) + 256 x 168 (softmax: a full 128-score
    // row per thread) + 128 x 120 (epilogue: O_a whole + O_b by halves) = 64,512.  setmaxnreg sits inside each role's
    // branch so that the compiler sees that role's budget only.
    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        setmaxnreg_dec<48>();
        if (lane == 0) {
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int seq = item / heads, h = item - seq * heads;
                const int row0 = seq * AT2_S;
                const int st = it & 1;
                mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * AT2_KV_BYTES + AT2_Q256_BYTES);
                uint8_t* k = sK + st * AT2_KV_BYTES;
                uint8_t* v = sV + st * AT2_KV_BYTES;
                tma_load_2d(k, &map128, &kv_full[st], D + h * 64, row0);
                tma_load_2d(k + AT2_TILE_BYTES, &map128, &kv_full[st], D + h * 64, row0 + 128);
                tma_load_2d(k + 2 * AT2_TILE_BYTES, &map16, &kv_full[st], D + h * 64, row0 + 256);
                tma_load_2d(v, &map128, &kv_full[st], 2 * D + h * 64, row0);
                tma_load_2d(v + AT2_TILE_BYTES, &map128, &kv_full[st], 2 * D + h * 64, row0 + 128);
                tma_load_2d(v + 2 * AT2_TILE_BYTES, &map16, &kv_full[st], 2 * D + h * 64, row0 + 256);
                tma_load_2d(sQ256 + st * AT2_Q256_BYTES, &map16, &kv_full[st], h * 64, row0 + 256);
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    mbar_wait(&q_empty[w], (it & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[w], AT2_TILE_BYTES);
                    tma_load_2d(sQ + w * AT2_TILE_BYTES, &map128, &q_full[w], h * 64, row0 + w * 128);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        setmaxnreg_dec<48>();
        const int w = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
        const uint32_t t_a = tmem_base + (2 * w) * 128, t_b = t_a + 128;       // the warpgroup's two unit buffers
        const uint64_t dq = umma_desc_k128(smem_u32(sQ + w * AT2_TILE_BYTES));

        auto issue_s = [&](int it) {                    // S of both key halves of the warpgroup's tile of item `it`
            const int st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            mbar_wait(&q_full[w], it & 1);
            const uint64_t dk = umma_desc_k128(smem_u32(sK + st * AT2_KV_BYTES));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (it > 0) mbar_wait(&u_free[2 * w + half], (it - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(half ? t_b : t_a, dq + 2 * kk, dk + half * (AT2_TILE_BYTES >> 4) + 2 * kk, idesc_s, kk != 0);
                    umma_commit(&s_full[2 * w + half]);
                    if (half) umma_commit(&q_empty[w]);
                }
                __syncwarp();
            }
        };

        if (my_items > 0) issue_s(0);
        for (int it = 0; it < my_items; ++it) {
            const int st = it & 1;
            const uint32_t v_base = smem_u32(sV + st * AT2_KV_BYTES);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&p_full[2 * w + half], it & 1);
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
            }
            if (it + 1 < my_items) issue_s(it + 1);
        }
    } else if (warp == 3) {
        setmaxnreg_dec<48>();                                 // idle warp of the first warpgroup (the instruction is warpgroup-wide)
    } else if (warp < 12) {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        setmaxnreg_inc<168>();
        const int w = (warp - 4) >> 2;                        // warpgroup = query tile of every item
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // query row inside the tile = TMEM lane
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (2 * w) * 128;
        // The hand-off of the row statistics uses shared-space st / ld: they and the mbarrier operations travel the same
        // in-order path.  Generic-address accesses (ST.E / LD.E) immediately followed by the arrive were observed to be
        // overtaken by it about once in 10^5 items (the last word of the record, p_256, arrived stale).
        const uint32_t my_stats = smem_u32(sStats + w * AT2_STATS_FLOATS + r * 8);
        for (int it = 0; it < my_items; ++it) {
            AT2_TRACE(w, it, 0);
            // ---- score of every row against key 256 (row 256 of K is row 0 of its own swizzle atom: unswizzled)
            const int st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            mbar_wait(&q_full[w], it & 1);
            const float s256 = at2_warp_dot_k256(smem_u32(sQ + w * AT2_TILE_BYTES), quad * 32,
                                                 reinterpret_cast<const uint32_t*>(sK + st * AT2_KV_BYTES + 2 * AT2_TILE_BYTES), lane);
            mbar_arrive(&q_empty[w]);
            mbar_arrive(&kv_empty[st]);
            AT2_TRACE(w, it, 1);
            // ---- the two units: S -> P, each with its own statistics
            float m_a, l_a, m_b, l_b, p256, unused;
            mbar_wait(&s_full[2 * w], it & 1);
            AT2_TRACE(w, it, 2);
            tc_fence_after();
            at2_softmax_unit(t_row, -INFINITY, scale_log2, m_a, l_a, unused);
            tc_fence_before();
            mbar_arrive(&p_full[2 * w]);
            AT2_TRACE(w, it, 3);
            mbar_wait(&s_full[2 * w + 1], it & 1);
            AT2_TRACE(w, it, 4);
            tc_fence_after();
            at2_softmax_unit(t_row + 128, s256, scale_log2, m_b, l_b, p256);
            tc_fence_before();
            mbar_arrive(&p_full[2 * w + 1]);
            AT2_TRACE(w, it, 5);

            // ---- hand the row statistics to the epilogue warpgroup
            mbar_wait(&st_empty[w], (it & 1) ^ 1);
            sts_f4(my_stats, make_float4(m_a, l_a, m_b, l_b));
            sts_f1(my_stats + 16, p256);
            mbar_arrive(&st_full[w]);
            AT2_TRACE(w, it, 6);
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue warpgroup
        setmaxnreg_dec<120>();                                // the kernel launches with 128 registers per thread
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // row of the tile this thread finishes; thread id for query 256
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        float* sPart = sRow + 272;                            // [4][64]
        float* sRed = sPart + 4 * 64;                         // [8]
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int seq = item / heads, h = item - seq * heads;
            const int st = it & 1;
            AT2_TRACE(3, it, 0);
            mbar_wait(&kv_full[st], (it >> 1) & 1);           // acquire the TMA writes of K, V and the query-256 box
            AT2_TRACE(3, it, 1);
            const uint8_t* kbase = sK + st * AT2_KV_BYTES;
            const uint8_t* vbase = sV + st * AT2_KV_BYTES;
            // ---- query 256 on mma.sync, spread over the four warps: warp `quad` takes keys [64 quad, 64 quad + 64), warp 3
            //      also key 256.  Only row 0 of the m16 fragments is real (lanes 0-3); v_256 is copied aside for the merges.
            {
                if (r < 8) sV256[st * 8 + r] = reinterpret_cast<const uint4*>(vbase + 2 * AT2_TILE_BYTES)[r];   // two copies: a slow
                                                              // warp may still be merging the previous item
                const uint32_t* q256 = reinterpret_cast<const uint32_t*>(sQ256 + st * AT2_Q256_BYTES);
                const uint32_t k_addr = smem_u32(kbase), v_addr = smem_u32(vbase);
                const int n_groups = (quad == 3) ? 5 : 4;     // 16-key groups; the fifth holds key 256 (+ 15 masked rows)
                const int key0 = quad * 64;
                const int t2 = (lane & 3) * 2;
                uint32_t qf[4][4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    qf[kk][0] = lane < 4 ? q256[kk * 8 + lane] : 0u;
                    qf[kk][2] = lane < 4 ? q256[kk * 8 + 4 + lane] : 0u;
                    qf[kk][1] = 0u; qf[kk][3] = 0u;
                }
                float sc[10][2];
                float mx = -INFINITY;
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < n_groups) {
                        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            uint32_t b0, b1, b2, b3;
                            at2_ldsm4(k_addr + at2_swz(key0 + g * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                            at2_mma16816(c0, qf[kk], b0, b1);
                            at2_mma16816(c1, qf[kk], b2, b3);
                        }
                        sc[2 * g][0] = c0[0] * scale_log2; sc[2 * g][1] = c0[1] * scale_log2;
                        sc[2 * g + 1][0] = c1[0] * scale_log2; sc[2 * g + 1][1] = c1[1] * scale_log2;
                        if (g == 4) {                         // keys 256..271: only key 256 exists
                            if (t2 != 0) sc[8][0] = -INFINITY;
                            sc[8][1] = -INFINITY; sc[9][0] = -INFINITY; sc[9][1] = -INFINITY;
                        }
                    } else {
                        sc[2 * g][0] = sc[2 * g][1] = sc[2 * g + 1][0] = sc[2 * g + 1][1] = -INFINITY;
                    }
                    mx = fmaxf(mx, fmaxf(fmaxf(sc[2 * g][0], sc[2 * g][1]), fmaxf(sc[2 * g + 1][0], sc[2 * g + 1][1])));
                }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                if (lane == 0) sRed[quad] = mx;
                AT2_TRACE(3, it, 2);
                named_bar_sync(4, 128);                       // also publishes sV256
                AT2_TRACE(3, it, 3);
                mx = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    sc[j][0] = at2_ex2(sc[j][0] - mx); sc[j][1] = at2_ex2(sc[j][1] - mx);
                    sum += sc[j][0] + sc[j][1];
                }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                AT2_TRACE(3, it, 4);
                float o[8][4];
#pragma unroll
                for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < n_groups) {
                        uint32_t pa[4];
                        pa[0] = pack_bf16x2(sc[2 * g][0], sc[2 * g][1]); pa[1] = 0u;
                        pa[2] = pack_bf16x2(sc[2 * g + 1][0], sc[2 * g + 1][1]); pa[3] = 0u;
#pragma unroll
                        for (int dd = 0; dd < 4; ++dd) {
                            uint32_t b0, b1, b2, b3;
                            at2_ldsm4_t(v_addr + at2_swz(key0 + g * 16 + (lane & 15), dd * 2 + (lane >> 4)), b0, b1, b2, b3);
                            at2_mma16816(o[2 * dd], pa, b0, b1);
                            at2_mma16816(o[2 * dd + 1], pa, b2, b3);
                        }
                    }
                }
                if (lane < 4) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) { sPart[quad * 64 + d * 8 + t2] = o[d][0]; sPart[quad * 64 + d * 8 + t2 + 1] = o[d][1]; }
                }
                if (lane == 0) sRed[4 + quad] = sum;
                mbar_arrive(&kv_empty[st]);                   // this warpgroup is done with K, V and the query-256 box of the item
                AT2_TRACE(3, it, 5);
                named_bar_sync(4, 128);
                AT2_TRACE(3, it, 6);
                if (quad == 0) {
                    const float inv = 1.0f / (sRed[4] + sRed[5] + sRed[6] + sRed[7]);
                    const float x0 = (sPart[2 * lane] + sPart[64 + 2 * lane] + sPart[128 + 2 * lane] + sPart[192 + 2 * lane]) * inv;
                    const float x1 = (sPart[2 * lane + 1] + sPart[64 + 2 * lane + 1] + sPart[128 + 2 * lane + 1] + sPart[192 + 2 * lane + 1]) * inv;
                    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(seq) * AT2_S + 256) * D + h * 64);
                    dst[lane] = pack_bf16x2(x0, x1);
                }
                named_bar_sync(4, 128);                       // sPart / sRed are rewritten by the next item
                AT2_TRACE(3, it, 7);
            }
            const uint4* v256 = sV256 + st * 8;
#pragma unroll 1
            for (int w = 0; w < 2; ++w) {
                const uint32_t t_row = t_lane + (2 * w) * 128;
                // ---- O_a as soon as it exists (the softmax warpgroup is still busy with half B): frees the unit early
                uint32_t oa[64];
                mbar_wait(&o_full[2 * w], it & 1);
                tc_fence_after();
                at2_ld32(t_row + 64, oa); at2_ld32(t_row + 96, oa + 32);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&u_free[2 * w]);
                AT2_TRACE(2, it, 2 * w);
                // ---- merge weights of the two halves from the row statistics
                mbar_wait(&st_full[w], it & 1);
                const uint32_t stats = smem_u32(sStats + w * AT2_STATS_FLOATS + r * 8);
                const float4 s4 = lds_f4(stats);
                const float p256 = __uint_as_float(lds_u1(stats + 16));
                mbar_arrive(&st_empty[w]);
                const float m = fmaxf(s4.x, s4.z);
                const float e_a = at2_ex2((s4.x - m) * scale_log2), e_b = at2_ex2((s4.z - m) * scale_log2);
                const float inv = 1.0f / fmaf(s4.y, e_a, s4.w * e_b);
                const float w_a = e_a * inv, w_b = e_b * inv, w_256 = p256 * w_b;
                const f32x2_t wa2 = f2_pack(w_a, w_a), wb2 = f2_pack(w_b, w_b), w2 = f2_pack(w_256, w_256);
                // ---- O_b by halves: out = O_a w_a + O_b w_b + p_256 w_b v_256, rows straight to global memory
                mbar_wait(&o_full[2 * w + 1], it & 1);
                tc_fence_after();
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(seq) * AT2_S + w * 128 + r) * D + h * 64);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t ob[32];
                    at2_ld32(t_row + 128 + 64 + hf * 32, ob);
                    tmem_ld_wait();
                    if (hf == 1) { tc_fence_before(); mbar_arrive(&u_free[2 * w + 1]); }
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
                        dst[hf * 4 + q] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
                    }
                }
                AT2_TRACE(2, it, 2 * w + 1);
            }

        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, AT2_TMEM_COLS);
}

#ifdef HB_EXP_TRACE
extern "C" int hb_exp_read_at2_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_at2_trace, sizeof(long long) * 4 * 64 * 16) == cudaSuccess ? 0 : -1;
}
#endif

int attention_tc2_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int heads, float scale, cudaStream_t stream) {
    const int D = heads * 64;
    const uint64_t rows = static_cast<uint64_t>(n_seq) * AT2_S;
    CUtensorMap map128, map16;
    if (encode_tmap_2d(&map128, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 128, 64)) return -1;
    if (encode_tmap_2d(&map16, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 16, 64)) return -1;
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(attention_tc2_kernel), AT2_SMEM)) return -1;
    const int n_items = n_seq * heads;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_tc2_kernel<<<grid, AT2_THREADS, AT2_SMEM, stream>>>(map128, map16, static_cast<__nv_bfloat16*>(out_bf16), n_items,
                                                                  heads, scale * 1.4426950408889634f);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
