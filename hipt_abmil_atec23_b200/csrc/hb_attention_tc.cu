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
// Key 256 does not fit a 128-key half: its score against every query is a 64-term dot product on the CUDA cores (thread =
// query row, k_256 broadcast as fp32), it joins half B's statistics, and p_256 v_256 is added in the epilogue.  Query 256
// has no tile either: the epilogue warps compute that row at the START of an item, while the softmax warpgroups are in its
// first unit: each warp takes 64 keys (scores: lane = two keys, q_256 broadcast as fp32; softmax against the warp's own
// maximum; P V: lane = two output dimensions) and the four (max, sum, partial output) records are merged after one named
// barrier.  Nothing runs on mma.sync: the trace of the round showed legacy HMMA at 50-75 cycles apiece on this part, and
// the 300-odd HMMAs per item the first version spent on the two odd rows set the item period.
//   warp 0        TMA producer: K / V [272 x 64] per item (two stages), one Q tile per softmax warpgroup, the 16-row box that
//                 starts at query 256
//   warps 1, 2    MMA issuers of softmax warpgroups 0 and 1 (whole warp walks the protocol, one elected lane issues); the
//                 next item's half-A score tile is issued as soon as the epilogue has drained O_a
//   warp 3        per item: k_256 and q_256 converted to fp32 in shared memory (broadcast operands of the CUDA-core dots)
//   warps 4-7     softmax warpgroup 0 (queries 0..127 of every item), warps 8-11 warpgroup 1 (queries 128..255)
//   warps 12-15   epilogue warpgroup: query 256; then drains O_a / O_b of both tiles, merges, stages each warp's 32 rows in
//                 swizzled shared memory and hands them to a TMA store (16-byte scattered row stores and the HMMA version
//                 of query 256 made this warpgroup, not the softmax, set the item period)
// Four units are resident (4 x 128 = all 512 TMEM columns).  Softmax statistics stay fp32; P is rounded to bf16 relative to
// its half's maximum.
#include <stdlib.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

#ifdef HB_EXP_TRACE
// timing experiment only (tools/build_exp.sh NAME -DHB_EXP_TRACE, tools/exp_at2_trace.py): clock64 stamps of CTA 0
__device__ long long g_at2_trace[4 * 64 * 16];
#define AT2_TRACE(role, it, k) do { if (blockIdx.x == 0 && (it) < 64 && (threadIdx.x & 127) == 0) g_at2_trace[((role) * 64 + (it)) * 16 + (k)] = clock64(); } while (0)
#define AT2_TRACE3(it, k) do { if (blockIdx.x == 0 && (it) < 64 && (threadIdx.x & 31) == 0) g_at2_trace[(3 * 64 + (it)) * 16 + (k)] = clock64(); } while (0)
#else
#define AT2_TRACE(role, it, k) do { } while (0)
#define AT2_TRACE3(it, k) do { } while (0)
#endif
#ifdef HB_EXP_AT2_DEBUG
// race hunt only: per (item, query row 0..255): s_256 and p_256 as the softmax thread computed them, p_256 as the epilogue read it
__device__ float g_at2_dbg[1536 * 256 * 4];
extern "C" int hb_exp_read_at2_dbg(float* out) {
    return cudaMemcpyFromSymbol(out, g_at2_dbg, sizeof(float) * 1536 * 256 * 4) == cudaSuccess ? 0 : -1;
}
#endif

constexpr int AT2_THREADS = 512;
constexpr int AT2_S = 257;
constexpr int AT2_KV_BYTES = 272 * 128;                 // [272 keys][64 bf16], 128-byte swizzled rows
constexpr int AT2_TILE_BYTES = 128 * 128;               // one 128-row tile of 128-byte rows
constexpr int AT2_Q256_BYTES = 16 * 128;                // the 16-row box that starts at query 256 (row 0 is unswizzled)
constexpr int AT2_STATS_FLOATS = 128 * 8;               // per warpgroup: [128 rows][m_a, l_a, m_b, l_b, p_256, pad]
// odd-row scratch (floats): [2 stages][k_256 64 | q_256 64] fp32 copies, [4 warps][68] probabilities of the warp's keys,
// [2 items][4 warps][64 partial outputs + max + sum + pad], [4 warps][2 stages][32] copies of v_256
constexpr int AT2_CVT = 0, AT2_PW = 256, AT2_PART = AT2_PW + 4 * 68, AT2_PART_W = 68, AT2_V256 = AT2_PART + 2 * 4 * AT2_PART_W;
constexpr int AT2_ROW_FLOATS = AT2_V256 + 4 * 2 * 32;
constexpr int AT2_OUT_BYTES = 4 * 2 * 4096;             // epilogue staging: [4 warps][2 tiles][32 rows x 128 B], swizzled
constexpr int AT2_SMEM = 4 * AT2_KV_BYTES + 2 * AT2_TILE_BYTES + AT2_OUT_BYTES + 2 * AT2_Q256_BYTES +
                         2 * AT2_STATS_FLOATS * 4 + AT2_ROW_FLOATS * 4 + 512 + 1024;
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

// q_r . k_256 for one row of a swizzled [128 x 64] bf16 Q tile; k_256 comes as 64 fp32 (broadcast reads)
__device__ __forceinline__ float at2_row_dot(uint32_t q_row, int sw, uint32_t kf_addr) {
    f32x2_t acc0 = f2_pack(0.f, 0.f), acc1 = acc0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 qv = lds_u4(q_row + ((c ^ sw) << 4));
        const float4 k0 = lds_f4(kf_addr + c * 32), k1 = lds_f4(kf_addr + c * 32 + 16);
        acc0 = f2_fma(f2_pack(at2_lo(qv.x), at2_hi(qv.x)), f2_pack(k0.x, k0.y), acc0);
        acc1 = f2_fma(f2_pack(at2_lo(qv.y), at2_hi(qv.y)), f2_pack(k0.z, k0.w), acc1);
        acc0 = f2_fma(f2_pack(at2_lo(qv.z), at2_hi(qv.z)), f2_pack(k1.x, k1.y), acc0);
        acc1 = f2_fma(f2_pack(at2_lo(qv.w), at2_hi(qv.w)), f2_pack(k1.z, k1.w), acc1);
    }
    float x0, x1;
    f2_unpack(f2_add(acc0, acc1), x0, x1);
    return x0 + x1;
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
                     const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, int n_items, int heads,
                     float scale_log2, int reverse) {
    extern __shared__ uint8_t smem_raw_at2[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at2) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                                   // [2 stages][34816]
    uint8_t* sV = sK + 2 * AT2_KV_BYTES;                  // [2 stages][34816]
    uint8_t* sQ = sV + 2 * AT2_KV_BYTES;                  // [2 softmax warpgroups][16384]
    uint8_t* sOut = sQ + 2 * AT2_TILE_BYTES;              // [4 epilogue warps][2 tiles][4096]
    uint8_t* sQ256 = sOut + AT2_OUT_BYTES;                // [2 stages][2048]
    float* sStats = reinterpret_cast<float*>(sQ256 + 2 * AT2_Q256_BYTES);       // [2 warpgroups][128][8]
    float* sRow = sStats + 2 * AT2_STATS_FLOATS;          // query 256 scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRow + AT2_ROW_FLOATS);
    uint64_t* kv_full = bars;            // [2 stages]   TMA bytes of K, V and the query-256 box
    uint64_t* kv_empty = bars + 2;       // [2 stages]   2 MMA commits (the item's last P V) + 256 softmax threads (key 256) + 128 epilogue threads (query 256, v_256) + warp 3 (fp32 copies)
    uint64_t* q_full = bars + 4;         // [2 warpgroups]
    uint64_t* q_empty = bars + 6;        // [2]          1 MMA commit (both S MMAs done) + 128 softmax threads (key-256 dot)
    uint64_t* s_full = bars + 8;         // [4 units]    S of the unit is in TMEM
    uint64_t* p_full = bars + 12;        // [4 units]    128 softmax threads: P is in TMEM
    uint64_t* o_full = bars + 16;        // [4 units]    O of the unit is in TMEM
    uint64_t* u_free = bars + 20;        // [4 units]    128 epilogue threads: O has been read, the unit's columns are free
    uint64_t* st_full = bars + 24;       // [2 warpgroups]  128 softmax threads: the tile's row statistics are in shared memory
    uint64_t* st_empty = bars + 26;      // [2]          128 epilogue threads
    uint64_t* cvt_ready = bars + 28;     // [2 stages]   warp 3: the fp32 copies of k_256 and q_256 are in shared memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = heads * 64;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2 + 256 + 128 + 1);
            mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 129);
            mbar_init(&st_full[i], 128); mbar_init(&st_empty[i], 128);
            mbar_init(&cvt_ready[i], 1);
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

    // register budget (64 K registers per SM): 128 x 56 (producer / MMA issuers / query-256 warp) + 256 x 168 (softmax: a
    // full 128-score row per thread) + 128 x 120 (epilogue: O_a whole + O_b by halves) = 65,536.  setmaxnreg sits inside
    // each role's branch so that the compiler sees that role's budget only.
    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        setmaxnreg_dec<56>();
        if (lane == 0) {
            for (int it = 0; it < my_items; ++it) {
                const int item = reverse ? n_items - 1 - (blockIdx.x + it * gridDim.x) : blockIdx.x + it * gridDim.x;
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
        setmaxnreg_dec<56>();
        const int w = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
        const uint32_t t_a = tmem_base + (2 * w) * 128, t_b = t_a + 128;       // the warpgroup's two unit buffers
        const uint64_t dq = umma_desc_k128(smem_u32(sQ + w * AT2_TILE_BYTES));

        // S of one key half of the warpgroup's tile of item `it` (inputs of the item already waited for)
        auto issue_s_half = [&](int it, int half) {
            const uint64_t dk = umma_desc_k128(smem_u32(sK + (it & 1) * AT2_KV_BYTES));
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(half ? t_b : t_a, dq + 2 * kk, dk + half * (AT2_TILE_BYTES >> 4) + 2 * kk, idesc_s, kk != 0);
                umma_commit(&s_full[2 * w + half]);
                if (half) umma_commit(&q_empty[w]);
            }
            __syncwarp();
        };
        auto issue_pv_half = [&](int it, int half) {
            const uint32_t v_base = smem_u32(sV + (it & 1) * AT2_KV_BYTES);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t t_u = half ? t_b : t_a;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                    // 8 x 16 keys: P columns 8 kk .., V rows 128 half + 16 kk ..
                    const uint64_t dv = umma_desc_k128(v_base + (half * 128 + kk * 16) * 128);
                    umma_bf16_ts(t_u + 64, t_u + 8 * kk, dv, idesc_pv, kk != 0);
                }
                umma_commit(&o_full[2 * w + half]);
                if (half) umma_commit(&kv_empty[it & 1]);
            }
            __syncwarp();
        };
        auto wait_inputs = [&](int it) {
            mbar_wait(&kv_full[it & 1], (it >> 1) & 1);
            mbar_wait(&q_full[w], it & 1);
        };
        // lane 0 polls, the warp follows its answer (a completed phase stays completed for the other lanes)
        auto ready = [&](uint64_t* bar, uint32_t parity) {
            return __shfl_sync(0xffffffffu, mbar_try_wait(bar, parity), 0) != 0;
        };

        if (my_items > 0) { wait_inputs(0); issue_s_half(0, 0); issue_s_half(0, 1); }
        for (int it = 0; it < my_items; ++it) {
            const bool more = it + 1 < my_items;
            mbar_wait(&p_full[2 * w], it & 1);
            issue_pv_half(it, 0);
            // The next item's half-A score tile goes out as soon as the epilogue has drained O_a of this item — normally
            // while the softmax warpgroup is still in half B — so the warpgroup finds it ready when it comes back to unit A.
            // Whichever of (P of half B, unit A free) comes first is served first.
            bool s_a = more, pv_b = true;
            const long long t0 = clock64();
            while (pv_b) {
                if (ready(&p_full[2 * w + 1], it & 1)) { issue_pv_half(it, 1); pv_b = false; }
                if (s_a && ready(&u_free[2 * w], it & 1) && ready(&kv_full[(it + 1) & 1], ((it + 1) >> 1) & 1) &&
                    ready(&q_full[w], (it + 1) & 1)) { issue_s_half(it + 1, 0); s_a = false; }
                if (clock64() - t0 > 4000000000LL) __trap();
            }
            if (more) {
                wait_inputs(it + 1);
                if (s_a) { mbar_wait(&u_free[2 * w], it & 1); issue_s_half(it + 1, 0); }
                mbar_wait(&u_free[2 * w + 1], it & 1);
                issue_s_half(it + 1, 1);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------------------------------ fp32 copies of k_256, q_256
        setmaxnreg_dec<56>();
        for (int it = 0; it < my_items; ++it) {
            const int st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);           // K and the query-256 box have landed (rows 256: unswizzled)
            const uint32_t kw = lds_u1(smem_u32(sK + st * AT2_KV_BYTES + 2 * AT2_TILE_BYTES) + lane * 4);
            const uint32_t qw = lds_u1(smem_u32(sQ256 + st * AT2_Q256_BYTES) + lane * 4);
            const uint32_t dst = smem_u32(sRow + AT2_CVT + st * 128) + lane * 8;
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(at2_lo(kw)), "f"(at2_hi(kw)) : "memory");
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst + 256), "f"(at2_lo(qw)), "f"(at2_hi(qw)) : "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&cvt_ready[st]); mbar_arrive(&kv_empty[st]); }
        }
    } else if (warp < 12) {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        setmaxnreg_inc<168>();
        const int w = (warp - 4) >> 2;                        // warpgroup = query tile of every item
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // query row inside the tile = TMEM lane
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (2 * w) * 128;
        // The hand-off of the row statistics uses shared-space st / ld, and the reader releases the record with
        // mbar_arrive_after_reads: generic-address accesses (ST.E / LD.E) immediately followed by a bare arrive were observed
        // to be overtaken by it about once in 10^5 items (the last word of the record, p_256, arrived stale).
        const uint32_t my_stats = smem_u32(sStats + w * AT2_STATS_FLOATS + r * 8);
        for (int it = 0; it < my_items; ++it) {
            AT2_TRACE(w, it, 0);
            // ---- score of every row against key 256 (row 256 of K is row 0 of its own swizzle atom: unswizzled)
            const int st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);
            mbar_wait(&q_full[w], it & 1);
            mbar_wait(&cvt_ready[st], (it >> 1) & 1);
            const float s256 = at2_row_dot(smem_u32(sQ + w * AT2_TILE_BYTES) + r * 128, r & 7, smem_u32(sRow + AT2_CVT + st * 128));
#ifdef HB_EXP_AT2_DEBUG
            const float dbg_q0 = 0.f;
#endif
            mbar_arrive_after_reads(&q_empty[w]);             // the Q row and k_256 were only loaded so far: see hb_ptx.cuh
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
#ifdef HB_EXP_AT2_DEBUG
            { const int item = reverse ? n_items - 1 - (blockIdx.x + it * gridDim.x) : blockIdx.x + it * gridDim.x; if (item < 1536) { float* d = g_at2_dbg + (static_cast<size_t>(item) * 256 + w * 128 + r) * 4; d[0] = s256; d[1] = p256; d[3] = dbg_q0; } }
#endif
            mbar_arrive(&st_full[w]);
            AT2_TRACE(w, it, 6);
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue warpgroup
        setmaxnreg_dec<120>();                                // the kernel launches with 128 registers per thread
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                       // row of the tile this thread finishes
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        uint8_t* my_out = sOut + quad * 8192;                 // this warp's two staging buffers (tile 0 / tile 1)
        for (int it = 0; it < my_items; ++it) {
            const int item = reverse ? n_items - 1 - (blockIdx.x + it * gridDim.x) : blockIdx.x + it * gridDim.x;
            const int seq = item / heads, h = item - seq * heads;
            const int st = it & 1;
            mbar_wait(&kv_full[st], (it >> 1) & 1);           // acquire the TMA writes of K and V
            mbar_wait(&cvt_ready[st], (it >> 1) & 1);
            const uint32_t v256 = smem_u32(sRow + AT2_V256 + (quad * 2 + st) * 32);   // this warp's copy of v_256 for the merges
            {
                // ---- query 256 over this warp's keys [64 quad, 64 quad + 64) (+ key 256 in warp 3 of the group)
                const uint32_t k_addr = smem_u32(sK + st * AT2_KV_BYTES), v_addr = smem_u32(sV + st * AT2_KV_BYTES);
                const uint32_t kf = smem_u32(sRow + AT2_CVT + st * 128), qf = kf + 256;
                if (lane < 8) sts_u4(v256 + lane * 16, lds_u4(v_addr + 2 * AT2_TILE_BYTES + lane * 16));   // row 256: unswizzled
                const int ka = quad * 64 + lane, kb = ka + 32;    // (ka & 7) == (kb & 7) == (lane & 7)
                float s_a = at2_row_dot(k_addr + ka * 128, lane & 7, qf) * scale_log2;
                float s_b = at2_row_dot(k_addr + kb * 128, lane & 7, qf) * scale_log2;
                float s_c = -INFINITY;
                if (quad == 3) {
                    float d = __uint_as_float(lds_u1(kf + lane * 8)) * __uint_as_float(lds_u1(qf + lane * 8)) +
                              __uint_as_float(lds_u1(kf + lane * 8 + 4)) * __uint_as_float(lds_u1(qf + lane * 8 + 4));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (lane == 0) s_c = d * scale_log2;      // key 256 counted once
                }
                float mw = fmaxf(fmaxf(s_a, s_b), s_c);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
                s_a = at2_ex2(s_a - mw); s_b = at2_ex2(s_b - mw); s_c = at2_ex2(s_c - mw);
                float lw = s_a + s_b + s_c;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lw += __shfl_xor_sync(0xffffffffu, lw, o);
                const uint32_t pw = smem_u32(sRow + AT2_PW + quad * 68);
                sts_f1(pw + lane * 4, s_a);
                sts_f1(pw + (32 + lane) * 4, s_b);
                if (lane == 0) sts_f1(pw + 64 * 4, s_c);
                __syncwarp();
                const uint32_t lane_off = ((lane & 3) << 2);
                f32x2_t acc[4] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f), f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk) {
                    const float4 p4 = lds_f4(pw + kk * 16);
                    const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = quad * 64 + kk * 4 + u;
                        const uint32_t vw = lds_u1(v_addr + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | lane_off));
                        acc[u] = f2_fma(f2_pack(pv[u], pv[u]), f2_pack(at2_lo(vw), at2_hi(vw)), acc[u]);
                    }
                }
                if (quad == 3) {
                    const float p = __uint_as_float(lds_u1(pw + 64 * 4));
                    const uint32_t vw = lds_u1(v_addr + 2 * AT2_TILE_BYTES + lane * 4);
                    acc[0] = f2_fma(f2_pack(p, p), f2_pack(at2_lo(vw), at2_hi(vw)), acc[0]);
                }
                float x0, x1;
                f2_unpack(f2_add(f2_add(acc[0], acc[1]), f2_add(acc[2], acc[3])), x0, x1);
                const uint32_t part = smem_u32(sRow + AT2_PART + (it & 1) * 4 * AT2_PART_W);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(part + (quad * AT2_PART_W + 2 * lane) * 4), "f"(x0), "f"(x1) : "memory");
                if (lane == 0)
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(part + (quad * AT2_PART_W + 64) * 4), "f"(mw), "f"(lw) : "memory");
                mbar_arrive_after_reads(&kv_empty[st]);       // this thread is done with K and V of the item
                named_bar_sync(4, 128);                       // also orders the v_256 copy before the merges' reads
                if (quad == 0) {
                    float mq[4], lq[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(mq[q]), "=f"(lq[q]) : "r"(part + (q * AT2_PART_W + 64) * 4) : "memory");
                    const float m = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
                    float l = 0.f, y0 = 0.f, y1 = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float e = at2_ex2(mq[q] - m);
                        float a0, a1;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0), "=f"(a1) : "r"(part + (q * AT2_PART_W + 2 * lane) * 4) : "memory");
                        l = fmaf(lq[q], e, l); y0 = fmaf(a0, e, y0); y1 = fmaf(a1, e, y1);
                    }
                    const float inv = 1.0f / l;
                    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(seq) * AT2_S + 256) * D + h * 64);
                    dst[lane] = pack_bf16x2(y0 * inv, y1 * inv);
                }
            }
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
                mbar_arrive_after_reads(&st_empty[w]);
#ifdef HB_EXP_AT2_DEBUG
                if (item < 1536) g_at2_dbg[(static_cast<size_t>(item) * 256 + w * 128 + r) * 4 + 2] = p256;
#endif
                const float m = fmaxf(s4.x, s4.z);
                const float e_a = at2_ex2((s4.x - m) * scale_log2), e_b = at2_ex2((s4.z - m) * scale_log2);
                const float inv = 1.0f / fmaf(s4.y, e_a, s4.w * e_b);
                const float w_a = e_a * inv, w_b = e_b * inv, w_256 = p256 * w_b;
                const f32x2_t wa2 = f2_pack(w_a, w_a), wb2 = f2_pack(w_b, w_b), w2 = f2_pack(w_256, w_256);
                // ---- O_b by halves: out = O_a w_a + O_b w_b + p_256 w_b v_256 into this warp's swizzled staging rows
                mbar_wait(&o_full[2 * w + 1], it & 1);
                tc_fence_after();
                if (lane == 0) tma_store_wait_read<1>();      // the store that read this buffer one item ago is done with it
                __syncwarp();
                const uint32_t srow = smem_u32(my_out + w * 4096) + lane * 128;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t ob[32];
                    at2_ld32(t_row + 128 + 64 + hf * 32, ob);
                    tmem_ld_wait();
                    if (hf == 1) { tc_fence_before(); mbar_arrive(&u_free[2 * w + 1]); }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 vv = lds_u4(v256 + (hf * 4 + q) * 16);
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
                        sts_u4(srow + (((hf * 4 + q) ^ (lane & 7)) << 4), make_uint4(o4[0], o4[1], o4[2], o4[3]));
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_out, my_out + w * 4096, h * 64, seq * AT2_S + w * 128 + quad * 32);
                    tma_store_commit();
                }
                AT2_TRACE(2, it, 2 * w + 1);
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
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
    CUtensorMap map_out;
    if (encode_tmap_2d(&map_out, TMAP_BF16, out_bf16, rows, D, static_cast<uint64_t>(D) * 2, 32, 64)) return -1;
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(attention_tc2_kernel), AT2_SMEM)) return -1;
    const int n_items = n_seq * heads;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    // Items are walked from the LAST sequence down: the qkv GEMM before this kernel wrote its tiles in ascending row order, so the
    // rows it wrote last are the ones still in L2, and the proj GEMM after it reads `out` in ascending order — starting with
    // the rows this kernel writes last (HB_ATT_REVERSE=0 restores the ascending walk for comparison).
    static int rev = -1;
    if (rev < 0) { const char* e = getenv("HB_ATT_REVERSE"); rev = (e && e[0] == '0') ? 0 : 1; }
    attention_tc2_kernel<<<grid, AT2_THREADS, AT2_SMEM, stream>>>(map128, map16, map_out, static_cast<__nv_bfloat16*>(out_bf16),
                                                                  n_items, heads, scale * 1.4426950408889634f, rev);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
