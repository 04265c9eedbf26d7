// bf16 tensor-core tier of the dense contraction:  C[M,N] = A[M,K] * B[N,K]^T + epilogue   (fp32 accumulate/out)
//
// Blackwell-native (sm_100a) kernel, hand-written PTX:
//   * operands: bf16, staged by TMA (cp.async.bulk.tensor.2d, 128B swizzle) into a shared-memory ring.  Each operand is
//     either K-major (rows of K contiguous elements: activations x, weights W of y = x W^T) or MN-major (rows of M / N
//     contiguous elements, one row per k: the operands of every weight-gradient contraction dW = dY^T X and of every
//     dX = dY W contraction).  The UMMA shared-memory / instruction descriptors carry the major-ness, so NO operand is
//     ever transposed in memory;
//   * math: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16 (BN = 256 / 128 / 64), issued by ONE elected
//     thread, accumulators in TMEM (2 stages x BN fp32 columns: the epilogue of tile i overlaps the MMAs of tile i+1);
//   * sync: mbarrier full/empty ring (TMA <-> MMA), tcgen05.commit -> mbarrier (MMA -> TMA slot release and
//     MMA -> epilogue), tmem_empty barrier (epilogue -> MMA);
//   * warp roles (384 threads): warps 0-2 TMA producers (k-blocks round-robin), warp 3 TMEM allocator + MMA issuer,
//     warps 4-11 epilogue
//     (tcgen05.ld 32x32b.x32 -> registers -> padded smem transpose -> coalesced 128-bit loads/stores with the fused
//     epilogue: bias1 + bias2 + add1 + add2, row mask, beta, optional bf16 copy of the result);
//   * persistent: grid = min(#work units, #SMs), static round-robin schedule;
//   * small problems (the per-time-step contractions with M = batch): narrower BN and deterministic split-K
//     (work unit = tile x K-slice, raw partial tiles to workspace, a second kernel sums the slices in fixed order
//     and applies the epilogue) so that one wave still covers the 148 SMs.
#include "common.cuh"
#include <stdlib.h>
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <cmath>
#include <cstdlib>

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
#ifndef ICD_GEMM_EPI_DEBUG
#define ICD_GEMM_EPI_DEBUG 0            // 1 / 2: diagnostic builds that skip the epilogue's global stores / its staging too (tools/gemm_bench.py --fc)
#endif
constexpr int EPI_LD = 32;                           // row of the per-warp 32x32 transpose buffer; the eight float4 groups of a
                                                     // row are XOR-swizzled with (row & 7) so that both the 128-bit row writes
                                                     // (lane = row) and the 128-bit tile reads (8 lanes = one row) are
                                                     // bank-conflict free without padding
constexpr int NUM_EPI_WARPS = 8;
// warp roles: warps 0-2 TMA producers (k-blocks round-robin: one UTMALDG costs ~100 issue clocks, a single producer thread
// caps the ring at ~400 clocks per k-block — twice the MMA time of a 128 x 64 tile), warp 3 TMEM allocator + MMA issuer,
// warps 4-11 epilogue.  12 warps = 3 per scheduler: 168 registers per thread (13 warps would put 4 on one scheduler: 128)
constexpr int NUM_PROD_WARPS = 3;
constexpr int MMA_WARP = NUM_PROD_WARPS;
constexpr int EPI_WARP0 = NUM_PROD_WARPS + 1;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);        // 384
constexpr int MN_BLOCK_BYTES = TC_MN_BLOCK_BYTES;    // one 64-element MN block of an MN-major tile: BK rows x 128 B

struct EpiArgs {
    float* C; long long ldc;
    int M, N, K;
    const float* bias1; const float* bias2;
    const float* add1; long long ld1;
    const float* add2; long long ld2;
    const unsigned char* row_mask;
    float beta;
    __nv_bfloat16* C16; long long ldc16;      // optional bf16 copy of the result (C may then be NULL)
    int vec;                                  // 4 / 2 / 1: widest vector (in floats) every row start of C / add / C16 allows
};

// Fused LSTMCell epilogue (models/attention.py:277-278 after the K4 contraction): the B operand is the GATE-PERMUTED copy of
// W_ih[:, E:] (row ug*32 + g*8 + j = natural row g*D + ug*8 + j, icd_convert_bf16_gateperm), so every 32-column chunk of the
// accumulator holds the four gates (i, f, g, o) of 8 hidden units and the epilogue thread that owns a row can finish the cell
// for them: pre = acc + xg + z_hh, c = f*c_prev + i*g, h = o*tanh(c), dropout(h) — no gates_pre round trip, no second kernel.
struct LstmEpi {
    int on, D;
    const float* xg; long long ld_xg;         // hoisted input term (+ both biases), natural gate layout [rows][4D]
    const float* zhh; long long ld_zhh;       // h W_hh^T columns of z, natural gate layout
    const float* c_prev;                      // [rows][D]
    float* gates_act;                         // [rows][4D] activated gates (saved for the backward)
    float* c_new; float* h_new;               // [rows][D]
    float* hdrop; long long hdrop_stride;     // dropout(h) rows hdrop_stride apart, may be NULL
    const unsigned char* mask; float scale;   // keep mask [rows][D] or NULL
    __nv_bfloat16* h16; __nv_bfloat16* hdrop16;
};

struct KArgs {
    EpiArgs e;
    LstmEpi lstm;
    int a_mn, b_mn;                           // 1: operand is MN-major in memory
    int splits, kb_per_split;                 // split-K: work unit = (tile, slice)
    float* partial;                           // [splits][M][N] raw accumulators when splits > 1
    int cm, cn;                               // thread-block cluster of cm x cn CTAs (cm along M, cn along N; 1 x 1 = none): the cn
                                              // CTAs of a cluster row share their A tile and the cm CTAs of a cluster column their
                                              // B tile — every CTA fetches 1/cn of A and 1/cm of B and TMA-multicasts the slice into
                                              // all shared memories that need it, so each operand byte leaves L2 once per cluster
    int late_trigger;                         // 1: dependents are not released early (an attention-step launch follows, see api.cu)
    int raster_n;                             // 1: consecutive work units walk along N (a wave of CTAs writes whole output rows:
                                              // contiguous DRAM pages) instead of along M
    const int* m_live;                        // optional DEVICE row count: only rows < min(M, *m_live) are computed / written
                                              // (beam search: the live rows are compacted to the front, no host round trip)
    int x3_kb_a, x3_kb_b;                     // fp32-grade tier, THREE stored planes per operand ([t1 | t2 | t3] along K, segment = x3_kb
                                              // k-blocks) instead of the six-segment K-concatenation: the producer maps k-block kb of the
                                              // 6-segment loop to the plane its segment uses (x3_kcoord); 0 = operand stored as it is walked
};

// k coordinate (elements) of k-block kb.  Six-term product of 3-term splits, segment order (a1 b1, a1 b2, a2 b1, a1 b3, a3 b1, a2 b2):
// A walks planes 0,0,1,0,2,1 and B planes 0,1,0,2,0,1 (one nibble per segment in `pat`).
constexpr uint32_t X3_PAT_A = 0x120100u, X3_PAT_B = 0x102010u;
__device__ __forceinline__ int x3_kcoord(int kb, int seg_kb, uint32_t pat) {
    if (seg_kb == 0) return kb * BK;
    const int sg = kb / seg_kb, kk = kb - sg * seg_kb;
    return ((int)((pat >> (4 * sg)) & 3u) * seg_kb + kk) * BK;
}

template <int BN>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * 32 * EPI_LD * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;         // two accumulator stages (power of two: 128, 256 or 512)
    // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), a_major bit 15, b_major bit 16,
    // n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ int epi_off4(int r, int g) { return r * EPI_LD + ((g ^ (r & 7)) << 2); }        // float4 group g of row r
__device__ __forceinline__ int epi_off(int r, int c) { return epi_off4(r, c >> 2) + (c & 3); }             // element (r, c)
__device__ __forceinline__ float4 epi_ld4(const float* sE, int r, int g) { return *reinterpret_cast<const float4*>(sE + epi_off4(r, g)); }

// One 32x32 sub-tile of the accumulator, already transposed into sE (swizzled, see EPI_LD): apply the fused epilogue and
// write it out with coalesced accesses.  rowbits: bit r set <=> row mrow0 + r is inside M and not masked out (the row
// mask is read ONCE per sub-tile, one coalesced byte load + a warp ballot, never in the store loop).
//   vec4 : lane -> 4 consecutive columns, 8 lanes per row, 4 rows per instruction (128-bit accesses)
//   vec2 : lane -> 2 consecutive columns, 16 lanes per row, 2 rows per instruction (64-bit; row stride even, e.g. V = 9490)
//   else : lane -> 1 column, 1 row per instruction (still 128 B coalesced); also the N tail of the other two
// Row validity of a 32-row strip (bit r: row mrow0 + r inside M / not masked out): ONE coalesced byte load + two ballots per
// tile, issued before the epilogue warp waits for the accumulator so the load latency hides behind the contraction.
__device__ __forceinline__ void epi_row_bits(const EpiArgs& e, int lane, int mrow0, unsigned& rows_in, unsigned& rows_keep) {
    unsigned inb = (mrow0 + lane < e.M) ? 1u : 0u;
    unsigned keep = inb;
    if (e.row_mask && inb) keep = e.row_mask[mrow0 + lane] ? 1u : 0u;
    rows_in = __ballot_sync(0xffffffffu, inb != 0); rows_keep = __ballot_sync(0xffffffffu, keep != 0);
}
// bias1 / bias2 of the lane's four columns of chunk nb in the 128-bit store paths (vec 4: columns 4j..4j+3; vec 3: the shifted
// mapping of the 8-byte-aligned rows), loaded before the accumulator chunk is read and NOT touched until the store loop (the
// first use is where the load latency would be paid); zero where those paths do not apply.
struct Bias4 { float4 a, b; };
__device__ __forceinline__ Bias4 epi_bias4(const EpiArgs& e, int lane, int nb) {
    Bias4 r; r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a;
    if (nb + 32 > e.N || (!e.bias1 && !e.bias2)) return r;
    if (e.vec >= 4) {
        const int n = nb + (lane & 7) * 4;
        if (e.bias1) r.a = *reinterpret_cast<const float4*>(e.bias1 + n);
        if (e.bias2) r.b = *reinterpret_cast<const float4*>(e.bias2 + n);
    } else if (e.vec == 3) {
        const int j = lane & 7;
        const bool odd = ((lane >> 3) & 1) != 0;
        float t[4], u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = (4 * j + (odd ? 2 : 0) + q) & 31;        // odd rows: rotated by two (the last lane holds 30, 31, 0, 1)
            t[q] = e.bias1 ? __ldg(e.bias1 + nb + col) : 0.f;
            u[q] = e.bias2 ? __ldg(e.bias2 + nb + col) : 0.f;
        }
        r.a = make_float4(t[0], t[1], t[2], t[3]); r.b = make_float4(u[0], u[1], u[2], u[3]);
    }
    return r;
}

// Staging of one accumulator chunk: the thread that owns accumulator row `lane` writes its 32 columns as eight float4 groups
// (XOR-swizzled with the row, see epi_off4).  rot: the row is an ODD row of an output whose rows are only 8-byte aligned
// (vec 3): its columns are stored rotated left by two — group g holds columns 4g+2 .. 4g+5 (mod 32) — so that the store
// loop reads ONE aligned group per lane on every row and the 128-bit global store of an odd row is aligned again.
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void epi_stage(float* sE, int lane, const uint32_t (&v)[32], bool rot) {
    // rows are 128 B long and 128-B aligned: group g of row r lives at (row + ((r & 7) << 4)) ^ (g << 4)
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sE + lane * EPI_LD) + ((lane & 7) << 4);
    if (!rot) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
            sts128(a0 ^ (uint32_t)(g << 4), __uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                   __uint_as_float(v[4 * g + 3]));
    } else {
#pragma unroll
        for (int g = 0; g < 8; ++g)
            sts128(a0 ^ (uint32_t)(g << 4), __uint_as_float(v[(4 * g + 2) & 31]), __uint_as_float(v[(4 * g + 3) & 31]),
                   __uint_as_float(v[(4 * g + 4) & 31]), __uint_as_float(v[(4 * g + 5) & 31]));
    }
}

// Store loop of a full 32-column chunk with bias + row mask only (no add / beta operands): lane -> one float4 group of 4
// rows per pass, 8 passes.  Everything that does not depend on the pass is computed once per chunk (the two swizzled
// shared-memory addresses a lane alternates between, the output pointer, which then only advances by 4 rows), and a strip
// whose 32 rows are all inside M and all kept takes a loop without per-row predicates: one shared-memory read, four adds and
// one 128-bit store per row.  SHIFT (vec 3): rows only 8-byte aligned on odd m, odd rows rotated by epi_stage — the last
// lane of an odd row writes its group as two 64-bit halves (columns 30-31 and 0-1).
template <bool SHIFT>
__device__ __forceinline__ void epi_store_plain(const float* __restrict__ sE, int lane, int mrow0, int nb, const EpiArgs& e,
                                                unsigned rows_in, unsigned rows_keep, const float4 b) {
    const int l3 = lane >> 3, j = lane & 7;
    const bool odd = SHIFT && (l3 & 1) != 0;                      // mrow0 is even: parity of row 4i + l3 == parity of l3
    const bool edge = odd && j == 7;
    const int c0 = edge ? 30 : 4 * j + (odd ? 2 : 0);
    const uint32_t srow = (uint32_t)__cvta_generic_to_shared(sE + l3 * EPI_LD);
    const uint32_t s_even = srow + (((j ^ l3) & 7) << 4), s_odd = srow + (((j ^ (l3 + 4)) & 7) << 4);   // (row & 7) = l3 / l3 + 4
    float4 sv[8];                                                 // the lane's eight groups first: the loads do not wait for the stores
#pragma unroll
    for (int i = 0; i < 8; ++i) sv[i] = lds128(((i & 1) ? s_odd : s_even) + (uint32_t)(i * 4 * EPI_LD * 4));
    float* dst = e.C ? e.C + (long long)(mrow0 + l3) * e.ldc + nb + c0 : nullptr;
    __nv_bfloat16* dst16 = (!SHIFT && e.C16) ? e.C16 + (long long)(mrow0 + l3) * e.ldc16 + nb + c0 : nullptr;
    const long long step = 4 * e.ldc, step16 = 4 * e.ldc16;
    auto put = [&](const float4 x) {
        if (SHIFT) {
            if (!edge) *reinterpret_cast<float4*>(dst) = x;
            else {
                *reinterpret_cast<float2*>(dst) = make_float2(x.x, x.y);                 // columns 30, 31
                *reinterpret_cast<float2*>(dst - 30) = make_float2(x.z, x.w);            // columns 0, 1
            }
        } else {
            if (dst) *reinterpret_cast<float4*>(dst) = x;
            if (dst16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dst16) = pk;
            }
        }
    };
    if ((rows_in & rows_keep) == 0xffffffffu) {                    // warp-uniform: the whole strip is stored as computed
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            put(make_float4(sv[i].x + b.x, sv[i].y + b.y, sv[i].z + b.z, sv[i].w + b.w));
            if (dst) dst += step;
            if (dst16) dst16 += step16;
        }
    } else {
        const unsigned rin = rows_in >> l3, rkp = rows_keep >> l3;   // bit 4i: row 4i + l3
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool kp = ((rkp >> (4 * i)) & 1u) != 0;
            const float4 x = kp ? make_float4(sv[i].x + b.x, sv[i].y + b.y, sv[i].z + b.z, sv[i].w + b.w) : make_float4(0.f, 0.f, 0.f, 0.f);
            if ((rin >> (4 * i)) & 1u) put(x);
            if (dst) dst += step;
            if (dst16) dst16 += step16;
        }
    }
}

__device__ __forceinline__ void epilogue_store_chunk(const float* __restrict__ sE, int lane, int mrow0, int nb,
                                                     const EpiArgs& e, unsigned rows_in, unsigned rows_keep, const Bias4& bias) {
    const float4 bsum = make_float4(bias.a.x + bias.b.x, bias.a.y + bias.b.y, bias.a.z + bias.b.z, bias.a.w + bias.b.w);
    if (e.vec >= 4 && nb + 32 <= e.N) {
        const int cc = (lane & 7) * 4, n = nb + cc;
        if (!e.add1 && !e.add2 && e.beta == 0.f) {                   // plain store (bias + row mask only)
            epi_store_plain<false>(sE, lane, mrow0, nb, e, rows_in, rows_keep, bsum);
            return;
        }
        // general path: the add / beta operands of four rows are fetched before those rows are stored (two halves keep the
        // register footprint of the operand prefetch at 32 instead of 64)
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            float4 a1[4], a2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = (h * 4 + i) * 4 + (lane >> 3), m = mrow0 + rr;
                a1[i] = make_float4(0.f, 0.f, 0.f, 0.f); a2[i] = a1[i];
                if ((rows_keep >> rr) & 1u) {
                    if (e.add1) a1[i] = *reinterpret_cast<const float4*>(e.add1 + (long long)m * e.ld1 + n);
                    if (e.add2) a2[i] = *reinterpret_cast<const float4*>(e.add2 + (long long)m * e.ld2 + n);
                    if (e.beta != 0.f && e.C) {
                        const float4 o = *reinterpret_cast<const float4*>(e.C + (long long)m * e.ldc + n);
                        a1[i].x += e.beta * o.x; a1[i].y += e.beta * o.y; a1[i].z += e.beta * o.z; a1[i].w += e.beta * o.w;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = (h * 4 + i) * 4 + (lane >> 3), m = mrow0 + rr;
                if (!((rows_in >> rr) & 1u)) continue;
                const float4 sv = epi_ld4(sE, rr, lane & 7);
                float4 x = make_float4(sv.x + bsum.x + a1[i].x + a2[i].x, sv.y + bsum.y + a1[i].y + a2[i].y,
                                       sv.z + bsum.z + a1[i].z + a2[i].z, sv.w + bsum.w + a1[i].w + a2[i].w);
                if (!((rows_keep >> rr) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e.C) *reinterpret_cast<float4*>(e.C + (long long)m * e.ldc + n) = x;
                if (e.C16) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(e.C16 + (long long)m * e.ldc16 + n) = pk;
                }
            }
        }
    } else if (e.vec == 3 && nb + 32 <= e.N) {
        // fp32 rows that are only 8-byte aligned on odd m (ldc % 4 == 2, e.g. the (B*T, V = 9490) logits): odd rows were staged
        // rotated by two columns (epi_stage), so every lane still moves one aligned float4 group per row.  bias + row mask only.
        epi_store_plain<true>(sE, lane, mrow0, nb, e, rows_in, rows_keep, bsum);
    } else if (e.vec >= 2 && nb + 32 <= e.N) {
        const int cc = (lane & 15) * 2, n = nb + cc;
        float2 bsum = make_float2(0.f, 0.f);
        if (e.bias1) { const float2 b = *reinterpret_cast<const float2*>(e.bias1 + n); bsum.x += b.x; bsum.y += b.y; }
        if (e.bias2) { const float2 b = *reinterpret_cast<const float2*>(e.bias2 + n); bsum.x += b.x; bsum.y += b.y; }
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += 8) {
            float2 a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = (i0 + u) * 2 + (lane >> 4), m = mrow0 + rr;
                a[u] = make_float2(0.f, 0.f);
                if ((rows_keep >> rr) & 1u) {
                    if (e.add1) { const float2 t = *reinterpret_cast<const float2*>(e.add1 + (long long)m * e.ld1 + n); a[u].x += t.x; a[u].y += t.y; }
                    if (e.add2) { const float2 t = *reinterpret_cast<const float2*>(e.add2 + (long long)m * e.ld2 + n); a[u].x += t.x; a[u].y += t.y; }
                    if (e.beta != 0.f && e.C) { const float2 t = *reinterpret_cast<const float2*>(e.C + (long long)m * e.ldc + n); a[u].x += e.beta * t.x; a[u].y += e.beta * t.y; }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = (i0 + u) * 2 + (lane >> 4), m = mrow0 + rr;
                if (!((rows_in >> rr) & 1u)) continue;
                const float2 sv = *reinterpret_cast<const float2*>(sE + epi_off(rr, cc));
                float2 x = make_float2(sv.x + bsum.x + a[u].x, sv.y + bsum.y + a[u].y);
                if (!((rows_keep >> rr) & 1u)) x = make_float2(0.f, 0.f);
                if (e.C) *reinterpret_cast<float2*>(e.C + (long long)m * e.ldc + n) = x;
                if (e.C16) *reinterpret_cast<__nv_bfloat162*>(e.C16 + (long long)m * e.ldc16 + n) = __floats2bfloat162_rn(x.x, x.y);
            }
        }
    } else {
        const int n = nb + lane;
        const bool nok = n < e.N;
        float bsum = 0.f;
        if (nok) { if (e.bias1) bsum += e.bias1[n]; if (e.bias2) bsum += e.bias2[n]; }
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
            float a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r0 + u, m = mrow0 + rr;
                a[u] = 0.f;
                if (nok && ((rows_keep >> rr) & 1u)) {
                    if (e.add1) a[u] += e.add1[(long long)m * e.ld1 + n];
                    if (e.add2) a[u] += e.add2[(long long)m * e.ld2 + n];
                    if (e.beta != 0.f && e.C) a[u] += e.beta * e.C[(long long)m * e.ldc + n];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r0 + u, m = mrow0 + rr;
                if (!nok || !((rows_in >> rr) & 1u)) continue;
                float x = sE[epi_off(rr, lane)] + bsum + a[u];
                if (!((rows_keep >> rr) & 1u)) x = 0.f;
                if (e.C) e.C[(long long)m * e.ldc + n] = x;
                if (e.C16) e.C16[(long long)m * e.ldc16 + n] = __float2bfloat16_rn(x);
            }
        }
    }
}

__device__ __forceinline__ void ld8(const float* p, float (&x)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&x)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&x)[8]) {
    uint4 pk;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(x[0], x[1]); pk.x = *reinterpret_cast<const uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[2], x[3]); pk.y = *reinterpret_cast<const uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[4], x[5]); pk.z = *reinterpret_cast<const uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[6], x[7]); pk.w = *reinterpret_cast<const uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p) = pk;
}
// One accumulator row of a 32-column chunk = gates (i, f, g, o) x 8 hidden units [8*ug, 8*ug + 8) of row m (see LstmEpi).
// The row's operands are fetched BEFORE the epilogue warp waits for the accumulator (the loads — lane-strided 32-byte
// segments, slow to issue — then hide behind the contraction); same arithmetic, in the same order, as the unfused pair
// (contraction epilogue acc + add1 + add2, then lstm_pointwise_fwd).
struct LstmRowOps { float a1[4][8], a2[4][8], cp[8]; uint2 mk; };
__device__ __forceinline__ void lstm_prefetch_row(LstmRowOps& o, int m, int ug, const LstmEpi& L) {
    const int D = L.D, u0 = ug * 8;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        ld8(L.xg + (long long)m * L.ld_xg + g * D + u0, o.a1[g]);
        ld8(L.zhh + (long long)m * L.ld_zhh + g * D + u0, o.a2[g]);
    }
    ld8(L.c_prev + (long long)m * D + u0, o.cp);
    o.mk = make_uint2(0x01010101u, 0x01010101u);
    if (L.hdrop && L.mask) o.mk = *reinterpret_cast<const uint2*>(L.mask + (long long)m * D + u0);
}
__device__ __forceinline__ void lstm_epilogue_row(const uint32_t (&v)[32], const LstmRowOps& o, int m, int ug, const LstmEpi& L) {
    const int D = L.D, u0 = ug * 8;
    float gi[8], gf[8], gg[8], go[8], c[8], h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gi[j] = sigmoidf_(__uint_as_float(v[j]) + o.a1[0][j] + o.a2[0][j]);
        gf[j] = sigmoidf_(__uint_as_float(v[8 + j]) + o.a1[1][j] + o.a2[1][j]);
        gg[j] = tanhf(__uint_as_float(v[16 + j]) + o.a1[2][j] + o.a2[2][j]);
        go[j] = sigmoidf_(__uint_as_float(v[24 + j]) + o.a1[3][j] + o.a2[3][j]);
        c[j] = gf[j] * o.cp[j] + gi[j] * gg[j];
        h[j] = go[j] * tanhf(c[j]);
    }
    if (L.gates_act) {
        float* ga = L.gates_act + (long long)m * 4 * D + u0;
        st8(ga, gi); st8(ga + D, gf); st8(ga + 2 * D, gg); st8(ga + 3 * D, go);
    }
    st8(L.c_new + (long long)m * D + u0, c);
    st8(L.h_new + (long long)m * D + u0, h);
    if (L.h16) st8_bf16(L.h16 + (long long)m * D + u0, h);
    if (L.hdrop) {
        float hd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool keep = (((j < 4 ? o.mk.x : o.mk.y) >> (8 * (j & 3))) & 0xffu) != 0;
            hd[j] = L.mask ? (keep ? h[j] * L.scale : 0.f) : h[j];
        }
        st8(L.hdrop + (long long)m * L.hdrop_stride + u0, hd);
        if (L.hdrop16) st8_bf16(L.hdrop16 + (long long)m * L.hdrop_stride + u0, hd);
    }
}

#ifdef ICD_GEMM_TRACE
__device__ long long g_trace[6][512];      // [0] producer: empty-wait done, [1] producer: TMA issued, [2] MMA: full-wait done, [3] MMA: issued
#define TRACE(slot, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_trace[slot][idx] = clock64(); } while (0)
#else
#define TRACE(slot, idx) do {} while (0)
#endif

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KArgs p) {
    using C_ = Cfg<BN>;
    constexpr int STAGES = C_::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + STAGES * A_BYTES;
    float* sEpi = reinterpret_cast<float*>(gbase + STAGES * C_::STAGE_BYTES);
    const uint32_t bars = base + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES + 16 * STAGES + 32);

    const EpiArgs& e = p.e;
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    // cluster cm x cn: the cluster works on a "super tile" of cm x cn tiles (m-tiles cm*mp + rm, n-tiles cn*np + rn); all of
    // its CTAs walk the same unit sequence, so their pipelines run in lock step through the shared empty barriers
    const int CM = p.cm, CN = p.cn, CL = CM * CN;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    const int rm = (int)crank % CM, rn = (int)crank / CM;               // position inside the cluster
    uint16_t mask_row = 0, mask_col = 0;                                // CTAs sharing this CTA's A tile (same rm) / B tile (same rn)
    for (int j = 0; j < CN; ++j) mask_row |= (uint16_t)(1u << (rm + CM * j));
    for (int i = 0; i < CM; ++i) mask_col |= (uint16_t)(1u << (rn * CM + i));
    const int tiles_n = (e.N + BN - 1) / BN;
    const int nkb = (e.K + BK - 1) / BK;
    const int unit0 = blockIdx.x / CL, unit_stride = gridDim.x / CL;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
        // a slot is free once every CTA this one multicasts into (its cluster row and column) has consumed it
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, (uint32_t)(CM + CN - 1)); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C_::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // the peer's barriers must be initialised before any multicast can target them
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
    // from here on operands / epilogue inputs produced upstream are read and C is written
    if (!p.late_trigger) pdl_trigger();
    pdl_wait();
    const int M_eff = p.m_live ? min(e.M, max((int)uniform_u32((uint32_t)*p.m_live), 0)) : e.M;     // device-side row count (read after the dependency wait)
    const int tiles_m = (M_eff + BM - 1) / BM;
    const int tiles_mc = (tiles_m + CM - 1) / CM, tiles_nc = (tiles_n + CN - 1) / CN;
    const int num_tiles = tiles_mc * tiles_nc;                           // super tiles
    const int num_units = num_tiles * p.splits;

    if (warp < NUM_PROD_WARPS) {
        // =================================== TMA producers ===================================
        // producer warp w fetches every NUM_PROD_WARPS-th k-block (both operands); ONE elected thread runs the whole loop (see
        // the MMA issuer).  A UTMALDG costs ~100 issue clocks: one producer alone caps the ring at ~400 clocks per k-block.
        // In a cluster this CTA fetches slice rn of its A tile for its cluster row and slice rm of its B tile for its cluster
        // column (rows of a K-major tile, k-rows of every 64-wide block of an MN-major one) and multicasts them.
        const int a_mn = p.a_mn, b_mn = p.b_mn;
        const uint32_t a_off = a_mn ? (uint32_t)(rn * (MN_BLOCK_BYTES / CN)) : (uint32_t)(rn * (A_BYTES / CN));
        const uint32_t b_off = b_mn ? (uint32_t)(rm * (MN_BLOCK_BYTES / CM)) : (uint32_t)(rm * (C_::B_BYTES / CM));
        const int a_row = a_mn ? 0 : rn * (BM / CN), a_k = a_mn ? rn * (BK / CN) : 0;
        const int b_row = b_mn ? 0 : rm * (BN / CM), b_k = b_mn ? rm * (BK / CM) : 0;
        if (elect_one()) {
            int stage = warp; uint32_t phase = 0;
            int it = 0;                                                   // k-blocks seen so far (all units)
            for (int unit = unit0; unit < num_units; unit += unit_stride) {
                const int tile = unit % num_tiles, slice = unit / num_tiles;
                const int tm = p.raster_n ? tile / tiles_nc : tile % tiles_mc, tn = p.raster_n ? tile % tiles_nc : tile / tiles_mc;
                const int m0 = (tm * CM + rm) * BM, n0 = (tn * CN + rn) * BN;
                const int kb0 = slice * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    if ((it % NUM_PROD_WARPS) != warp) continue;
                    mbar_wait_backoff(empty0 + 8 * stage, phase ^ 1);   // every CTA of the cluster row / column has consumed this slot
                    TRACE(0, kb);
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_arrive_expect_tx(fb, C_::STAGE_BYTES);      // all slices of both tiles, whoever sends them
                    const uint32_t a_dst = sA + stage * A_BYTES + a_off, b_dst = sB + stage * C_::B_BYTES + b_off;
                    const int ka = x3_kcoord(kb, p.x3_kb_a, X3_PAT_A), kbn = x3_kcoord(kb, p.x3_kb_b, X3_PAT_B);
                    if (!a_mn) {
                        if (CN == 1) tma_load_2d(a_dst, &tmA, ka, m0 + a_row, fb);
                        else tma_load_2d_mc(a_dst, &tmA, ka, m0 + a_row, fb, mask_row);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) {
                            if (CN == 1) tma_load_2d(a_dst + j * MN_BLOCK_BYTES, &tmA, m0 + 64 * j, ka, fb);
                            else tma_load_2d_mc(a_dst + j * MN_BLOCK_BYTES, &tmA, m0 + 64 * j, ka + a_k, fb, mask_row);
                        }
                    }
                    if (!b_mn) {
                        if (CM == 1) tma_load_2d(b_dst, &tmB, kbn, n0 + b_row, fb);
                        else tma_load_2d_mc(b_dst, &tmB, kbn, n0 + b_row, fb, mask_col);
                    } else {
#pragma unroll
                        for (int j = 0; j < (BN + 63) / 64; ++j) {
                            if (CM == 1) tma_load_2d(b_dst + j * MN_BLOCK_BYTES, &tmB, n0 + 64 * j, kbn, fb);
                            else tma_load_2d_mc(b_dst + j * MN_BLOCK_BYTES, &tmB, n0 + 64 * j, kbn + b_k, fb, mask_col);
                        }
                    }
                    TRACE(1, kb);
                    stage += NUM_PROD_WARPS;
                    if (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // =================================== MMA issuer ===================================
        // ONE elected thread runs the whole loop: leaving an elect region after every k-block costs ~100 clocks of
        // reconvergence behind the UTCHMMAs, and a UTCHMMA issues at best every 48 clocks whatever its N — the issuing thread
        // is the bottleneck of the narrow tiles, so nothing else may sit on its critical path: the barrier of the NEXT ring
        // slot is probed before the MMAs of the current one are issued (tools/micro/mma_loop.cu)
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint32_t idesc = C_::IDESC | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16);
            const int a_mn = p.a_mn, b_mn = p.b_mn;
            const uint64_t a_kstep = a_mn ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);   // descriptor units of 16 B
            const uint64_t b_kstep = b_mn ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);
            const uint16_t mc_all = (uint16_t)(mask_row | mask_col);
            uint32_t ready = 0;                                           // the current slot's barrier was already seen complete
            for (int unit = unit0; unit < num_units; unit += unit_stride) {
                const int slice = unit / num_tiles;
                const int kb0 = slice * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);              // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    TRACE(4, kb);
                    if (!ready) mbar_wait(full0 + 8 * stage, phase);      // TMA bytes have landed
                    TRACE(5, kb);
                    tc_fence_after();
                    const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
                    const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1) : phase;
                    ready = mbar_test(full0 + 8 * nstage, nphase);        // consumed in the next iteration
                    TRACE(2, kb);
                    const uint64_t adesc = make_smem_desc(sA + stage * A_BYTES, a_mn);
                    const uint64_t bdesc = make_smem_desc(sB + stage * C_::B_BYTES, b_mn);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        tc_mma_f16(d_tmem, adesc + (uint64_t)k * a_kstep, bdesc + (uint64_t)k * b_kstep, idesc,
                                   (kb > kb0 || k > 0) ? 1u : 0u);
                    if (CL == 1) tc_commit(empty0 + 8 * stage);           // smem slot reusable once these MMAs retire
                    else tc_commit_mc(empty0 + 8 * stage, mc_all);        // ... signalled to every producer that fills it
                    if (kb == kb1 - 1) tc_commit(tfull0 + 8 * acc);       // accumulator complete -> epilogue
                    TRACE(3, kb);
                    stage = nstage; phase = nphase;
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // =================================== epilogue warps 4..11 ===================================
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        const int cg = (warp - EPI_WARP0) >> 2;                       // column group: chunks cg, cg+2, ...
        float* sE = sEpi + (warp - EPI_WARP0) * 32 * EPI_LD;
        int acc = 0; uint32_t acc_phase = 0;
        constexpr int NCHUNK = BN / 32;
        for (int unit = unit0; unit < num_units; unit += unit_stride) {
            const int tile = unit % num_tiles, slice = unit / num_tiles;
            const int tm = p.raster_n ? tile / tiles_nc : tile % tiles_mc, tn = p.raster_n ? tile % tiles_nc : tile / tiles_mc;
                const int m0 = (tm * CM + rm) * BM, n0 = (tn * CN + rn) * BN;
            EpiArgs ee = e;
            ee.M = M_eff;
            if (p.splits > 1) {                                       // raw partial tile, epilogue deferred
                ee.C = p.partial + (long long)slice * e.M * e.N; ee.ldc = e.N;
                ee.bias1 = ee.bias2 = ee.add1 = ee.add2 = nullptr; ee.row_mask = nullptr; ee.beta = 0.f; ee.C16 = nullptr;
                ee.vec = (e.N % 4 == 0) ? 4 : ((e.N % 2 == 0) ? 2 : 1);
            }
            const int mrow0 = m0 + q * 32;
            LstmRowOps lops;
            const bool lstm_row = p.lstm.on && (n0 + cg * 32 < e.N) && (mrow0 + lane < M_eff);
            if (lstm_row) lstm_prefetch_row(lops, mrow0 + lane, (n0 + cg * 32) >> 5, p.lstm);   // hides behind the contraction
            unsigned rows_in, rows_keep;                                 // ... and so does the row mask
            epi_row_bits(ee, lane, mrow0, rows_in, rows_keep);
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            bool released = false;
#pragma unroll 1
            for (int c = cg; c < NCHUNK; c += 2) {
                const int nb = n0 + c * 32;
                const bool last = (c + 2 >= NCHUNK);
                if (nb < e.N && mrow0 < M_eff) {                        // warp-uniform
                    const Bias4 bias_c = epi_bias4(ee, lane, nb);    // in flight while the accumulator chunk is read and staged
                    uint32_t v[32];
                    tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                    tc_wait_ld();
                    if (last) {                                       // all of this warp's TMEM reads are done
                        tc_fence_before();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                        released = true;
                    }
                    if (p.lstm.on) {                                  // fused LSTMCell: the thread finishes its row's 8 units
                        if (lstm_row) lstm_epilogue_row(v, lops, mrow0 + lane, nb >> 5, p.lstm);   // BN = 64: chunk c == cg
                        continue;
                    }
#if ICD_GEMM_EPI_DEBUG >= 2
                    if (e.K < 0)                                     // diagnostic build: accumulator read, nothing staged / stored
#endif
                    epi_stage(sE, lane, v, ee.vec == 3 && nb + 32 <= ee.N && (lane & 1));
                    __syncwarp();
#if ICD_GEMM_EPI_DEBUG >= 1
                    if (e.K < 0)                                     // diagnostic build: staged, never stored
#endif
                    epilogue_store_chunk(sE, lane, mrow0, nb, ee, rows_in, rows_keep, bias_c);
                    __syncwarp();
                }
            }
            if (!released) {
                tc_fence_before();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // the peer may still multicast into this CTA's shared memory / barriers
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)C_::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant: tcgen05.mma.cta_group::2.  Two CTAs of a thread-block cluster (same TPC) compute ONE 256 x BN tile:
// each CTA stages its own 128 rows of A and only HALF of the B tile (BN/2 rows of N); the MMA, issued by one thread of
// the leader CTA (cluster rank 0), reads A and B from both shared memories and writes rows 0-127 of the accumulator into
// the leader's TMEM and rows 128-255 into the peer's.  Per SM and k-block this moves 16 KB (A) + BN/2 * 128 B (B half)
// instead of 16 KB + BN * 128 B — the single-CTA kernel saturates the SM's L2 ingress at ~65 % tensor-pipe utilisation.
//   * TMA loads of BOTH CTAs complete on the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2, barrier
//     address mapped into the leader with mapa); the leader's producer arms it with the bytes of both CTAs;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster releases the smem slot in both CTAs (each producer waits on its
//     own empty barrier) and announces the finished accumulator on both CTAs' tmem-full barriers;
//   * every epilogue warp of both CTAs arrives on the LEADER's tmem-empty barrier (remote mbarrier.arrive), which the MMA
//     thread waits on before reusing an accumulator stage.
// Everything else (roles, epilogue, split-K, operand majors, PDL) is as in gemm_tc_kernel.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <int BN>
struct Cfg2 {
    static constexpr int BH_BYTES = (BN / 2) * BK * 2;                  // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;              // per CTA
    static constexpr int STAGES = BN == 256 ? 6 : 8;
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * 32 * EPI_LD * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    // UMMA M = 256 (m_dim = 16), N = BN
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KArgs p) {
    using C_ = Cfg2<BN>;
    constexpr int STAGES = C_::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + STAGES * A_BYTES;
    float* sEpi = reinterpret_cast<float*>(gbase + STAGES * C_::STAGE_BYTES);
    const uint32_t bars = base + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES + 16 * STAGES + 32);

    const EpiArgs& e = p.e;
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();                           // 0 = leader
    const int tiles_n = (e.N + BN - 1) / BN;
    const int nkb = (e.K + BK - 1) / BK;
    const int unit0 = blockIdx.x / 2, unit_stride = gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C_::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    if (!p.late_trigger) pdl_trigger();
    pdl_wait();
    const int M_eff = p.m_live ? min(e.M, max((int)uniform_u32((uint32_t)*p.m_live), 0)) : e.M;
    const int tiles_m = (M_eff + BM - 1) / BM;
    const int tiles_mc = (tiles_m + 1) / 2;
    const int num_tiles = tiles_mc * tiles_n;                            // 256 x BN super tiles
    const int num_units = num_tiles * p.splits;

    if (warp < NUM_PROD_WARPS) {
        // =================================== TMA producers (both CTAs) ===================================
        // producer warp w fetches every NUM_PROD_WARPS-th k-block: this CTA's 128 rows of A and its half of B; the bytes of both
        // CTAs complete on the LEADER's full barrier, armed by the leader's producer of that k-block
        const int a_mn = p.a_mn, b_mn = p.b_mn;
        if (elect_one()) {
            int stage = warp; uint32_t phase = 0;
            int it = 0;
            for (int unit = unit0; unit < num_units; unit += unit_stride) {
                const int tile = unit % num_tiles, slice = unit / num_tiles;
                const int tm = p.raster_n ? tile / tiles_n : tile % tiles_mc, tn = p.raster_n ? tile % tiles_n : tile / tiles_mc;
                const int m0 = (tm * 2 + (int)crank) * BM, n0 = tn * BN + (int)crank * (BN / 2);
                const int kb0 = slice * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    if ((it % NUM_PROD_WARPS) != warp) continue;
                    mbar_wait_backoff(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb_leader = mapa_u32(full0 + 8 * stage, 0);
                    if (crank == 0) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * C_::STAGE_BYTES);   // bytes of BOTH CTAs
                    const uint32_t a_dst = sA + stage * A_BYTES, b_dst = sB + stage * C_::BH_BYTES;
                    const int ka = x3_kcoord(kb, p.x3_kb_a, X3_PAT_A), kbn = x3_kcoord(kb, p.x3_kb_b, X3_PAT_B);
                    if (!a_mn) tma_load_2d_2sm(a_dst, &tmA, ka, m0, fb_leader);
                    else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(a_dst + j * MN_BLOCK_BYTES, &tmA, m0 + 64 * j, ka, fb_leader);
                    }
                    if (!b_mn) tma_load_2d_2sm(b_dst, &tmB, kbn, n0, fb_leader);
                    else {
#pragma unroll
                        for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(b_dst + j * MN_BLOCK_BYTES, &tmB, n0 + 64 * j, kbn, fb_leader);
                    }
                    stage += NUM_PROD_WARPS;
                    if (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // =================================== MMA issuer (leader CTA only) ===================================
        // one elected thread runs the whole loop; the next ring slot's barrier is probed before the current MMAs are issued
        if (crank == 0 && elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int a_mn = p.a_mn, b_mn = p.b_mn;
            const uint32_t idesc = C_::IDESC | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
            const uint64_t a_kstep = a_mn ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);
            const uint64_t b_kstep = b_mn ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);
            uint32_t ready = 0;
            for (int unit = unit0; unit < num_units; unit += unit_stride) {
                const int slice = unit / num_tiles;
                const int kb0 = slice * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);          // the epilogues of BOTH CTAs have drained this stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!ready) mbar_wait(full0 + 8 * stage, phase);  // both CTAs' tiles have landed
                    tc_fence_after();
                    const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
                    const uint32_t nphase = (stage + 1 == STAGES) ? (phase ^ 1) : phase;
                    ready = mbar_test(full0 + 8 * nstage, nphase);
                    const uint64_t adesc = make_smem_desc(sA + stage * A_BYTES, a_mn);
                    const uint64_t bdesc = make_smem_desc(sB + stage * C_::BH_BYTES, b_mn);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        tc_mma_f16_2sm(d_tmem, adesc + (uint64_t)k * a_kstep, bdesc + (uint64_t)k * b_kstep, idesc,
                                       (kb > kb0 || k > 0) ? 1u : 0u);
                    tc_commit_2sm(empty0 + 8 * stage, (uint16_t)0x3);                 // slot free in both CTAs
                    if (kb == kb1 - 1) tc_commit_2sm(tfull0 + 8 * acc, (uint16_t)0x3);   // accumulator ready in both
                    stage = nstage; phase = nphase;
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // =================================== epilogue warps 4..11 (both CTAs, own TMEM half) ===================================
        const int q = warp & 3;
        const int cg = (warp - EPI_WARP0) >> 2;
        float* sE = sEpi + (warp - EPI_WARP0) * 32 * EPI_LD;
        int acc = 0; uint32_t acc_phase = 0;
        constexpr int NCHUNK = BN / 32;
        for (int unit = unit0; unit < num_units; unit += unit_stride) {
            const int tile = unit % num_tiles, slice = unit / num_tiles;
            const int tm = p.raster_n ? tile / tiles_n : tile % tiles_mc, tn = p.raster_n ? tile % tiles_n : tile / tiles_mc;
            const int m0 = (tm * 2 + (int)crank) * BM, n0 = tn * BN;
            EpiArgs ee = e;
            ee.M = M_eff;
            if (p.splits > 1) {
                ee.C = p.partial + (long long)slice * e.M * e.N; ee.ldc = e.N;
                ee.bias1 = ee.bias2 = ee.add1 = ee.add2 = nullptr; ee.row_mask = nullptr; ee.beta = 0.f; ee.C16 = nullptr;
                ee.vec = (e.N % 4 == 0) ? 4 : ((e.N % 2 == 0) ? 2 : 1);
            }
            const int mrow0 = m0 + q * 32;
            unsigned rows_in, rows_keep;                                 // row mask fetched while the contraction runs
            epi_row_bits(ee, lane, mrow0, rows_in, rows_keep);
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t tempty_leader = mapa_u32(tempty0 + 8 * acc, 0);
            bool released = false;
#pragma unroll 1
            for (int c = cg; c < NCHUNK; c += 2) {
                const int nb = n0 + c * 32;
                const bool last = (c + 2 >= NCHUNK);
                if (nb < e.N && mrow0 < M_eff) {
                    const Bias4 bias_c = epi_bias4(ee, lane, nb);
                    uint32_t v[32];
                    tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                    tc_wait_ld();
                    if (last) {
                        tc_fence_before();
                        if (lane == 0) mbar_arrive_remote(tempty_leader);
                        released = true;
                    }
#if ICD_GEMM_EPI_DEBUG >= 2
                    if (e.K < 0)                                     // diagnostic build: accumulator read, nothing staged / stored
#endif
                    epi_stage(sE, lane, v, ee.vec == 3 && nb + 32 <= ee.N && (lane & 1));
                    __syncwarp();
#if ICD_GEMM_EPI_DEBUG >= 1
                    if (e.K < 0)                                     // diagnostic build: staged, never stored
#endif
                    epilogue_store_chunk(sE, lane, mrow0, nb, ee, rows_in, rows_keep, bias_c);
                    __syncwarp();
                }
            }
            if (!released) {
                tc_fence_before();
                if (lane == 0) mbar_arrive_remote(tempty_leader);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)C_::TMEM_COLS) : "memory");
    }
}

// split-K second pass: C = epilogue(sum_s partial[s]) with the slices summed in fixed order (deterministic).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, const EpiArgs e) {
    pdl_trigger();
    pdl_wait();
    const long long MN = (long long)e.M * e.N;
    if (e.vec >= 4 && e.N % 4 == 0) {
        const long long total = MN / 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long o = i * 4;
            const int m = (int)(o / e.N), n = (int)(o % e.N);
            float4 x = *reinterpret_cast<const float4*>(partial + o);
            for (int s = 1; s < splits; ++s) {
                const float4 y = *reinterpret_cast<const float4*>(partial + (long long)s * MN + o);
                x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            }
            if (e.bias1) { const float4 b = *reinterpret_cast<const float4*>(e.bias1 + n); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (e.bias2) { const float4 b = *reinterpret_cast<const float4*>(e.bias2 + n); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (e.add1) { const float4 b = *reinterpret_cast<const float4*>(e.add1 + (long long)m * e.ld1 + n); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (e.add2) { const float4 b = *reinterpret_cast<const float4*>(e.add2 + (long long)m * e.ld2 + n); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (e.row_mask && !e.row_mask[m]) x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.C) {
                float* c = e.C + (long long)m * e.ldc + n;
                if (e.beta != 0.f) { const float4 o4 = *reinterpret_cast<const float4*>(c); x.x += e.beta * o4.x; x.y += e.beta * o4.y; x.z += e.beta * o4.z; x.w += e.beta * o4.w; }
                *reinterpret_cast<float4*>(c) = x;
            }
            if (e.C16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(e.C16 + (long long)m * e.ldc16 + n) = pk;
            }
        }
    } else {
        for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < MN; o += (long long)gridDim.x * blockDim.x) {
            const int m = (int)(o / e.N), n = (int)(o % e.N);
            float x = partial[o];
            for (int s = 1; s < splits; ++s) x += partial[(long long)s * MN + o];
            if (e.bias1) x += e.bias1[n];
            if (e.bias2) x += e.bias2[n];
            if (e.add1) x += e.add1[(long long)m * e.ld1 + n];
            if (e.add2) x += e.add2[(long long)m * e.ld2 + n];
            if (e.row_mask && !e.row_mask[m]) x = 0.f;
            if (e.C) { float* c = e.C + (long long)m * e.ldc + n; if (e.beta != 0.f) x += e.beta * (*c); *c = x; }
            if (e.C16) e.C16[(long long)m * e.ldc16 + n] = __float2bfloat16_rn(x);
        }
    }
}

// ---------------------------------------------------------------------------------------------- fp32 -> bf16 staging
// dst[r*ldd + c] = bf16(src[r*s_r + c]), r < rows, c < cols (source rows contiguous).
__global__ void convert_rows_kernel(const float* __restrict__ src, long long s_r, int rows, int cols,
                                    __nv_bfloat16* __restrict__ dst, long long ldd, int vec) {
    pdl_trigger();
    pdl_wait();
    if (vec) {                                          // cols % 4 == 0, rows 16 B aligned: 128-bit loads, 64-bit stores
        const long long n4 = cols >> 2, total = (long long)rows * n4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / n4; const int c = (int)(i % n4) * 4;
            const float4 x = ld_stream_f4(src + r * s_r + c);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dst + r * ldd + c) = pk;
        }
        return;
    }
    const long long n2 = ((long long)cols + 1) / 2;
    const long long total = (long long)rows * n2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n2; const int c = (int)(i % n2) * 2;
        const float* s = src + r * s_r + c;
        const float x = s[0], y = (c + 1 < cols) ? s[1] : 0.f;
        if (c + 1 < cols || (ldd & 1) == 0)          // ldd is even (multiple of 8): the pair store stays inside the row
            *reinterpret_cast<__nv_bfloat162*>(dst + r * ldd + c) = __floats2bfloat162_rn(x, y);
        else dst[r * ldd + c] = __float2bfloat16_rn(x);
    }
}

// up to ICD_CVT_MAX_SEGS conversions of the 128-bit kind in one grid: work item i (one float4) belongs to the segment whose
// prefix range holds it
struct CvtBatch {
    const float* src[ICD_CVT_MAX_SEGS]; __nv_bfloat16* dst[ICD_CVT_MAX_SEGS];
    long long s_r[ICD_CVT_MAX_SEGS], ldd[ICD_CVT_MAX_SEGS], end[ICD_CVT_MAX_SEGS];      // end: exclusive prefix sum of float4 items
    int n4[ICD_CVT_MAX_SEGS];                                                           // float4 items per row
    int n;
};
__global__ void __launch_bounds__(256) convert_rows_batch_kernel(const CvtBatch b) {
    pdl_trigger();
    pdl_wait();
    const long long total = b.end[b.n - 1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int g = 0;
        while (i >= b.end[g]) ++g;
        const long long j = i - (g ? b.end[g - 1] : 0);
        const long long r = j / b.n4[g]; const int c = (int)(j % b.n4[g]) * 4;
        const float4 x = ld_stream_f4(b.src[g] + r * b.s_r[g] + c);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(b.dst[g] + r * b.ldd[g] + c) = pk;
    }
}

// gate-permuting variant for an LSTM weight block [4D][cols] (gate-major rows i|f|g|o): destination row
// ug*32 + g*8 + j = source row g*D + ug*8 + j, so that 32 consecutive rows hold the four gates of 8 hidden units (LstmEpi).
__global__ void convert_rows_gateperm_kernel(const float* __restrict__ src, long long s_r, int D, int cols,
                                             __nv_bfloat16* __restrict__ dst, long long ldd) {
    pdl_trigger();
    pdl_wait();
    const long long n4 = cols >> 2, total = 4LL * D * n4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int rp = (int)(i / n4), c = (int)(i % n4) * 4;          // destination row
        const int ug = rp >> 5, g = (rp >> 3) & 3, j = rp & 7;
        const float4 x = ld_stream_f4(src + (long long)(g * D + ug * 8 + j) * s_r + c);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dst + (long long)rp * ldd + c) = pk;
    }
}

// transposing variant: src is contiguous along r (s_r == 1, s_c = source leading dimension).  32x32 smem tiles.
__global__ void convert_transpose_kernel(const float* __restrict__ src, long long s_c, int rows, int cols,
                                         __nv_bfloat16* __restrict__ dst, long long ldd) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // read: lanes along r (contiguous in src)
        const int c = c0 + j, r = r0 + threadIdx.x;
        t[j][threadIdx.x] = (c < cols && r < rows) ? src[(long long)c * s_c + r] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // write: lanes along c (contiguous in dst)
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)r * ldd + c] = __float2bfloat16_rn(t[threadIdx.x][j]);
    }
}

// ---------------------------------------------------------------------------------------------- fp32 -> 3 x bf16 split
// x = b1 + b2 + b3 with b1 = bf16(x), b2 = bf16(x - b1), b3 = bf16(x - b1 - b2)  (24 mantissa bits kept).
// The six significant cross terms of a product are laid out along K so that ONE bf16 contraction over K' = 6 * seg sums
//   A': [a1 | a1 | a2 | a1 | a3 | a2]      B': [b1 | b2 | b1 | b3 | b1 | b2]
// which: 0 = A pattern, 1 = B pattern.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& b1, __nv_bfloat16& b2, __nv_bfloat16& b3) {
    b1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b2);
    b3 = __float2bfloat16_rn(r2);
}
// K-major source rows: dst[r*ldd + s*seg + c] = term_s(src[r*s_r + c]), c < cols; pad columns [cols, seg) zeroed.
// vec8 = 1 (cols % 8 == 0, 16-byte aligned rows): a thread splits 8 consecutive columns — two 128-bit loads, six 128-bit stores.
__global__ void split3_rows_kernel(const float* __restrict__ src, long long s_r, int rows, int cols, int seg,
                                   __nv_bfloat16* __restrict__ dst, long long ldd, int which, const int* __restrict__ m_live,
                                   int vec8) {
    if (m_live) rows = min(rows, max(*m_live, 0));              // device-side row count (beam search)
    if (vec8) {
        const int seg8 = seg >> 3;
        const long long total = (long long)rows * seg8;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / seg8; const int c = (int)(i % seg8) * 8;
            const float4 x0 = *reinterpret_cast<const float4*>(src + r * s_r + c);
            const float4 x1 = *reinterpret_cast<const float4*>(src + r * s_r + c + 4);
            const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            uint32_t t[3][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                __nv_bfloat16 a[3], b[3];
                split3(xs[2 * q], a[0], a[1], a[2]);
                split3(xs[2 * q + 1], b[0], b[1], b[2]);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    t[k][q] = (uint32_t)__bfloat16_as_ushort(a[k]) | ((uint32_t)__bfloat16_as_ushort(b[k]) << 16);
            }
            const int pa[6] = {0, 0, 1, 0, 2, 1}, pb[6] = {0, 1, 0, 2, 0, 1};
            __nv_bfloat16* d = dst + r * ldd + c;
            if (which == 2) {                                     // three stored planes [t1 | t2 | t3]
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    *reinterpret_cast<uint4*>(d + (long long)k * seg) = make_uint4(t[k][0], t[k][1], t[k][2], t[k][3]);
                continue;
            }
#pragma unroll
            for (int sgm = 0; sgm < 6; ++sgm) {
                const int k = which == 0 ? pa[sgm] : pb[sgm];
                *reinterpret_cast<uint4*>(d + (long long)sgm * seg) = make_uint4(t[k][0], t[k][1], t[k][2], t[k][3]);
            }
        }
        return;
    }
    const long long total = (long long)rows * seg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / seg; const int c = (int)(i % seg);
        __nv_bfloat16 t[3];
        if (c < cols) split3(src[r * s_r + c], t[0], t[1], t[2]);
        else t[0] = t[1] = t[2] = __float2bfloat16_rn(0.f);
        __nv_bfloat16* d = dst + r * ldd + c;
        if (which == 2)      { d[0] = t[0]; d[seg] = t[1]; d[2 * seg] = t[2]; }
        else if (which == 0) { d[0] = t[0]; d[seg] = t[0]; d[2 * seg] = t[1]; d[3 * seg] = t[0]; d[4 * seg] = t[2]; d[5 * seg] = t[1]; }
        else                 { d[0] = t[0]; d[seg] = t[1]; d[2 * seg] = t[0]; d[3 * seg] = t[2]; d[4 * seg] = t[0]; d[5 * seg] = t[1]; }
    }
}
// MN-major source ([K][MN] rows): dst[(s*seg + k)*ldd + m] = term_s(src[k*s_k + m])   (seg >= K: rows per segment)
__global__ void split3_mn_kernel(const float* __restrict__ src, long long s_k, int K, int mn, int seg,
                                 __nv_bfloat16* __restrict__ dst, long long ldd, int which) {
    const long long total = (long long)K * mn;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / mn; const int m = (int)(i % mn);
        __nv_bfloat16 t[3];
        split3(src[k * s_k + m], t[0], t[1], t[2]);
        const int pa[6] = {0, 0, 1, 0, 2, 1}, pb[6] = {0, 1, 0, 2, 0, 1};
        if (which == 2) {
#pragma unroll
            for (int sgm = 0; sgm < 3; ++sgm) dst[((long long)sgm * seg + k) * ldd + m] = t[sgm];
            continue;
        }
#pragma unroll
        for (int sgm = 0; sgm < 6; ++sgm) dst[((long long)sgm * seg + k) * ldd + m] = t[which == 0 ? pa[sgm] : pb[sgm]];
    }
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KArgs& k, int units, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        ICD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
        attr_set = true;
    }
    if (k.cm * k.cn > 1) {                     // units = super tiles x slices; one cluster per unit slot
        const int cl = k.cm * k.cn;
        static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // co-resident clusters of this size (queried once)
        if (max_clusters[cl] == 0) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cl * (ICD_NUM_SMS / cl)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = Cfg<BN>::SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = ICD_NUM_SMS / cl / 2; }
            max_clusters[cl] = n < ICD_NUM_SMS / cl ? n : ICD_NUM_SMS / cl;
            if (getenv("ICD_GEMM_VERBOSE")) fprintf(stderr, "[icd] gemm_tc<%d>: %d co-resident clusters of %d CTAs\n", BN, max_clusters[cl], cl);
        }
        const int clusters = units < max_clusters[cl] ? units : max_clusters[cl];
        ICD_CUDA(icd_launch_pdl_cluster(k.late_trigger ? ICD_PDL_GEMM_LATE : ICD_PDL_GEMM, gemm_tc_kernel<BN>, dim3(cl * clusters), dim3(NUM_THREADS), (size_t)Cfg<BN>::SMEM, s,
                                        (unsigned)cl, tmA, tmB, k));
    } else {
        const int grid = units < ICD_NUM_SMS ? units : ICD_NUM_SMS;
        ICD_CUDA(icd_launch_pdl(k.late_trigger ? ICD_PDL_GEMM_LATE : ICD_PDL_GEMM, gemm_tc_kernel<BN>, dim3(grid), dim3(NUM_THREADS), (size_t)Cfg<BN>::SMEM, s, tmA, tmB, k));
    }
    ICD_LAUNCH_CHECK();
    return 0;
}

template <int BN>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const KArgs& k, int units, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        ICD_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM));
        attr_set = true;
    }
    const int pairs = units < ICD_NUM_SMS / 2 ? units : ICD_NUM_SMS / 2;
    ICD_CUDA(icd_launch_pdl_cluster(k.late_trigger ? ICD_PDL_GEMM_LATE : ICD_PDL_GEMM, gemm_tc2_kernel<BN>, dim3(2 * pairs), dim3(NUM_THREADS), (size_t)Cfg2<BN>::SMEM, s, 2u,
                                    tmA, tmB, k));
    ICD_LAUNCH_CHECK();
    return 0;
}

// Pairing mode of the contraction (ICD_GEMM_PAIR / icd_gemm_set_pair_mode):
//   0 = single-CTA tiles only; 1 = CTA pairs sharing the B tile by TMA multicast (cta_group::1 MMAs);
//   2 = CTA pairs computing one 256 x BN tile with tcgen05.mma.cta_group::2 (default where the shape allows it)
int g_pair_mode = -1;
bool g_pair_forced = false;               // set through icd_gemm_set_pair_mode / ICD_GEMM_PAIR: apply the mode to every eligible shape
int pair_mode() {
    if (g_pair_mode < 0) {
        const char* e = getenv("ICD_GEMM_PAIR");
        if (e && e[0] >= '0' && e[0] <= '2') { g_pair_mode = e[0] - '0'; g_pair_forced = true; } else g_pair_mode = 2;
    }
    return g_pair_mode;
}

// tile width / split-K plan from a small cost model (SM clocks): one persistent wave of work units should cover the
// 148 SMs; narrower tiles pay more shared-memory / L2 traffic per flop, split-K pays a second (reduce) kernel.
struct Plan { int bn, splits, kb_per_split; double cost; };

Plan make_plan(int M, int N, int K, bool allow_split, bool deferred = false) {
    // Cost model in SM clocks, fitted to CUDA-graph replays of the in-loop shapes (tools/gemm_bench.py --plans / --kslope --graph,
    // profiles/r02_gemm_plans.txt) AFTER the issue path was fixed (one elected thread per role, four producer warps):
    //   one k-block costs 295 / 362 / 567 clk at BN = 64 / 128 / 256 — a UTCHMMA issues at best every 48 clk whatever its N, so the
    //   narrow tiles are bound by the issuing thread, the 256-wide one by the tensor pipe (4 x 128 clk);
    //   launch + prologue + epilogue 7400 / 8500 / 12000; a split-K reduce pass 3900 + 500 clk per MB of partial planes, or — when
    //   the consumer sums the planes itself (deferred) — only the extra plane traffic, 300 clk per MB.
    const int tm = (M + BM - 1) / BM, nkb = (K + BK - 1) / BK;
    const int cand[3] = {256, 128, 64};
    const double kb_clk[3] = {567.0, 362.0, 295.0}, fixed_clk[3] = {12000.0, 8500.0, 7400.0};
    Plan best; best.bn = 64; best.splits = 1; best.kb_per_split = nkb; best.cost = 1e300;
    double best_cost = 1e300;
    const double plane_mb = (double)M * N * 4.0 / 1048576.0;
    for (int i = 0; i < 3; ++i) {
        const int bn = cand[i];
        if (i < 2 && N <= cand[i + 1]) continue;                 // a narrower tile already covers N
        const int tiles = tm * ((N + bn - 1) / bn);
        int smax = 1;
        if (allow_split && nkb >= 8) {
            smax = ICD_NUM_SMS / tiles;
            if (smax > nkb / 4) smax = nkb / 4;
            if (smax > 32) smax = 32;
            if (smax < 1) smax = 1;
        }
        for (int sp = 1; sp <= smax; ++sp) {
            const int kbs = (nkb + sp - 1) / sp;
            if ((nkb + kbs - 1) / kbs != sp) continue;           // this slice length leaves an empty slice
            const double units = (double)tiles * sp;
            double waves = units / ICD_NUM_SMS;                  // static round-robin: the busiest CTA does ceil() units,
            if (waves < 1.0) waves = 1.0;                        // but short tiles overlap their epilogues: blend
            else waves = 0.5 * (waves + std::ceil(waves));
            const double split_cost = sp > 1 ? (deferred ? 300.0 * sp * plane_mb : 3900.0 + 500.0 * sp * plane_mb) : 0.0;
            const double cost = waves * kbs * kb_clk[i] + fixed_clk[i] + split_cost;
            if (cost < best_cost) { best_cost = cost; best.bn = bn; best.splits = sp; best.kb_per_split = kbs; best.cost = cost; }
        }
    }
    return best;
}

// A last m-tile with only a few rows can cost a whole extra wave (d fc.weight: M = 9490 = 74 * 128 + 18 -> 150 tiles on 148
// SMs).  Returns the number of tail rows to run as a second, split-K contraction when the model says that is cheaper.
int tail_rows_to_split(int M, int N, int K, bool allow_split) {
    const int rem = M % BM;
    if (!allow_split || M <= BM || rem == 0 || rem > 32) return 0;
    const double c_full = make_plan(M, N, K, true).cost;
    const double c_two = make_plan(M - rem, N, K, true).cost + make_plan(rem, N, K, true).cost;
    return c_two < 0.9 * c_full ? rem : 0;
}

// Cluster shape (cm x cn) of the multicast kernel.  The in-loop contractions with M = batch are bound by the chip-wide
// L2 -> shared-memory rate (~6.3 KB/clk): with 128 x 64 tiles every k-block pulls 24 KB per CTA for 1 MFLOP.  In a cm x cn
// cluster each A byte leaves L2 once per cluster row and each B byte once per cluster column (per CTA and k-block:
// 16/cn + BN/8/cm KB).  ICD_GEMM_CLUSTER="cm,cn" overrides (diagnostic).
void cluster_shape(int M, int N, int K, int bn, int splits, bool dev_rows, int* cm, int* cn) {
    *cm = 1; *cn = 1;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + bn - 1) / bn;
    int fm = 0, fn = 0;
    if (const char* f = getenv("ICD_GEMM_CLUSTER")) {
        if (sscanf(f, "%d,%d", &fm, &fn) == 2 && fm >= 1 && fn >= 1 && fm * fn <= 8 && (fm & (fm - 1)) == 0 && (fn & (fn - 1)) == 0) {
            if (tiles_m % fm == 0 && tiles_n % fn == 0 && bn / fm >= 8) { *cm = fm; *cn = fn; }
            return;
        }
    }
    (void)K; (void)splits; (void)dev_rows;
}

}  // namespace

extern "C" int icd_has_tensor_core_gemm(void) { return 1; }
#ifdef ICD_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int icd_gemm_trace_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 6 * 512);
}
#endif
extern "C" int icd_gemm_set_pair_mode(int mode) {
    const int old = g_pair_forced ? pair_mode() : -1;
    pair_mode();
    if (mode >= 0 && mode <= 2) { g_pair_mode = mode; g_pair_forced = true; }
    else { g_pair_mode = 2; g_pair_forced = false; }          // any other value: back to the built-in policy
    return old;
}

int icd_convert_bf16(const float* src, int64_t s_r, int64_t s_c, int rows, int cols, void* dst, int64_t ldd,
                     cudaStream_t s) {
    if (rows == 0 || cols == 0) return 0;
    ICD_CHECK_ARG(s_r == 1 || s_c == 1, "convert_bf16: source needs a unit stride");
    ICD_CHECK_ARG(ldd % 8 == 0 && ldd >= cols, "convert_bf16: ldd=%lld must be a multiple of 8 and >= cols", (long long)ldd);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    if (s_c == 1) {
        const int vec = (cols % 4 == 0) && (s_r % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const long long total = vec ? (long long)rows * (cols / 4) : (long long)rows * ((cols + 1) / 2);
        long long blocks = (total + 255) / 256;
        if (blocks > ICD_NUM_SMS * 16) blocks = ICD_NUM_SMS * 16;
        ICD_CUDA(icd_launch_pdl(ICD_PDL_POINTWISE, convert_rows_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, s, src, (long long)s_r, rows, cols,
                                d, (long long)ldd, vec));
    } else {
        dim3 grid((rows + 31) / 32, (cols + 31) / 32);
        ICD_CHECK_ARG(grid.y <= 65535, "convert_bf16: too many columns for the transposing path");
        convert_transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(src, s_c, rows, cols, d, ldd);
    }
    ICD_LAUNCH_CHECK();
    return 0;
}

int icd_convert_bf16_batch(const IcdCvtSeg* segs, int n, cudaStream_t s) {
    CvtBatch b = {};
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        const IcdCvtSeg& g = segs[i];
        if (g.rows == 0 || g.cols == 0) continue;
        const bool vec = (g.cols % 4 == 0) && (g.s_r % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.src) & 15) == 0) && g.ldd % 8 == 0 &&
                         g.ldd >= g.cols && ((reinterpret_cast<uintptr_t>(g.dst) & 7) == 0);
        if (!vec || b.n == ICD_CVT_MAX_SEGS) { ICD_TRY(icd_convert_bf16(g.src, g.s_r, 1, g.rows, g.cols, g.dst, g.ldd, s)); continue; }
        total += (long long)g.rows * (g.cols / 4);
        b.src[b.n] = g.src; b.dst[b.n] = reinterpret_cast<__nv_bfloat16*>(g.dst); b.s_r[b.n] = g.s_r; b.ldd[b.n] = g.ldd;
        b.n4[b.n] = g.cols / 4; b.end[b.n] = total; ++b.n;
    }
    if (b.n == 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > ICD_NUM_SMS * 16) blocks = ICD_NUM_SMS * 16;
    ICD_CUDA(icd_launch_pdl(ICD_PDL_POINTWISE, convert_rows_batch_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, s, b));
    ICD_LAUNCH_CHECK();
    return 0;
}

int icd_convert_bf16_gateperm(const float* src, int64_t s_r, int D, int cols, void* dst, int64_t ldd, cudaStream_t s) {
    ICD_CHECK_ARG(D % 8 == 0 && cols % 4 == 0 && s_r % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && ldd % 8 == 0,
                  "convert_bf16_gateperm: D %% 8, cols %% 4 and 16-byte aligned rows required (D=%d cols=%d)", D, cols);
    const long long total = 4LL * D * (cols / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > ICD_NUM_SMS * 16) blocks = ICD_NUM_SMS * 16;
    ICD_CUDA(icd_launch_pdl(ICD_PDL_POINTWISE, convert_rows_gateperm_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, s, src, (long long)s_r, D, cols,
                            reinterpret_cast<__nv_bfloat16*>(dst), (long long)ldd));
    ICD_LAUNCH_CHECK();
    return 0;
}

// gates contraction + LSTMCell in one kernel (see LstmEpi): A16 = gated [rows][K] bf16, Wp16 = gate-permuted W_ih[:, E:] [4D][K]
int icd_gemm_bf16_lstm_cell(const void* A16, int64_t lda, const void* Wp16, int64_t ldb, int rows, int D, int K,
                            const float* xg, int64_t ld_xg, const float* zhh, int64_t ld_zhh, const float* c_prev,
                            float* gates_act, float* c_new, float* h_new, float* hdrop, int64_t hdrop_stride,
                            const uint8_t* mask, float scale, void* h16, void* hdrop16, cudaStream_t s) {
    if (rows == 0) return 0;
    ICD_CHECK_ARG(D % 16 == 0 && K > 0, "gemm_lstm_cell: D must be a multiple of 16 (D=%d)", D);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    ICD_CHECK_ARG(al16(xg) && al16(zhh) && al16(c_prev) && al16(gates_act) && al16(c_new) && al16(h_new) && al16(hdrop) && al16(h16) &&
                  al16(hdrop16) && (!mask || (reinterpret_cast<uintptr_t>(mask) & 7) == 0) && ld_xg % 4 == 0 && ld_zhh % 4 == 0 &&
                  hdrop_stride % 8 == 0, "gemm_lstm_cell: operands must be 16-byte aligned");
    const int N = 4 * D;
    CUtensorMap tmA, tmB;
    ICD_TRY(make_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(A16), lda, rows, K, BM, 0));
    ICD_TRY(make_tmap(&tmB, reinterpret_cast<const __nv_bfloat16*>(Wp16), ldb, N, K, 64, 0));
    KArgs k = {};
    EpiArgs& e = k.e;
    e.C = nullptr; e.ldc = 0; e.M = rows; e.N = N; e.K = K; e.vec = 4; e.beta = 0.f;
    k.a_mn = 0; k.b_mn = 0; k.splits = 1; k.kb_per_split = (K + BK - 1) / BK; k.partial = nullptr; k.cm = 1; k.cn = 1; k.m_live = nullptr;
    LstmEpi& L = k.lstm;
    L.on = 1; L.D = D; L.xg = xg; L.ld_xg = ld_xg; L.zhh = zhh; L.ld_zhh = ld_zhh; L.c_prev = c_prev; L.gates_act = gates_act;
    L.c_new = c_new; L.h_new = h_new; L.hdrop = hdrop; L.hdrop_stride = hdrop_stride; L.mask = mask; L.scale = scale;
    L.h16 = reinterpret_cast<__nv_bfloat16*>(h16); L.hdrop16 = reinterpret_cast<__nv_bfloat16*>(hdrop16);
    const int units = ((rows + BM - 1) / BM) * (N / 64);
    return launch<64>(tmA, tmB, k, units, s);
}

int64_t icd_gemm_bf16_splitk_floats(int M, int N, int K) {
    auto need = [&](int m) {                                   // either flavour of the plan (reduce pass / deferred to the consumer)
        const Plan p = make_plan(m, N, K, true, false), q = make_plan(m, N, K, true, true);
        const int sp = p.splits > q.splits ? p.splits : q.splits;
        return sp > 1 ? (int64_t)sp * m * N : (int64_t)0;
    };
    const int rem = tail_rows_to_split(M, N, K, true);
    return rem ? std::max(need(M - rem), need(rem)) : need(M);
}

static thread_local int g_x3_planes[2] = {0, 0};
void icd_gemm_x3_planes(int seg_kb_a, int seg_kb_b) { g_x3_planes[0] = seg_kb_a; g_x3_planes[1] = seg_kb_b; }

int icd_gemm_bf16_ex(const void* A16, int64_t lda, int a_mn, const void* B16, int64_t ldb, int b_mn,
                     float* C, int64_t ldc, int M, int N, int K,
                     const float* bias1, const float* bias2, const float* add1, int64_t ld1,
                     const float* add2, int64_t ld2, const uint8_t* row_mask, float beta, cudaStream_t s,
                     void* C16, int64_t ldc16, float* splitk_ws, int64_t splitk_ws_floats, int* deferred_splits,
                     const int* m_live) {
    if (deferred_splits) *deferred_splits = 0;
    if (M == 0 || N == 0) return 0;
    ICD_CHECK_ARG(K > 0, "gemm_tc: K must be positive");
    ICD_CHECK_ARG(C != nullptr || C16 != nullptr, "gemm_tc: no output");
    if (!m_live && !deferred_splits) {
        // few rows in the last m-tile and an extra wave because of them: run them as a second (split-K) contraction
        const int rem = tail_rows_to_split(M, N, K, splitk_ws != nullptr);
        if (rem && icd_gemm_bf16_splitk_floats(M, N, K) <= splitk_ws_floats) {
            const int Mm = M - rem;
            auto offA = [&](int m) { return reinterpret_cast<const char*>(A16) + 2 * (a_mn ? (int64_t)m : (int64_t)m * lda); };
            for (int part = 0; part < 2; ++part) {
                const int m0 = part ? Mm : 0, mm = part ? rem : Mm;
                ICD_TRY(icd_gemm_bf16_ex(offA(m0), lda, a_mn, B16, ldb, b_mn, C ? C + (int64_t)m0 * ldc : nullptr, ldc, mm, N, K,
                                         bias1, bias2, add1 ? add1 + (int64_t)m0 * ld1 : nullptr, ld1,
                                         add2 ? add2 + (int64_t)m0 * ld2 : nullptr, ld2, row_mask ? row_mask + m0 : nullptr, beta, s,
                                         C16 ? reinterpret_cast<char*>(C16) + 2 * (int64_t)m0 * ldc16 : nullptr, ldc16,
                                         splitk_ws, splitk_ws_floats, nullptr, nullptr));
            }
            return 0;
        }
    }
    Plan pl = make_plan(M, N, K, splitk_ws != nullptr && m_live == nullptr, deferred_splits != nullptr);   // a device-side row count excludes split-K
    if (const char* f = getenv("ICD_GEMM_FORCE_PLAN")) {           // diagnostic: "bn,splits" (tools/gemm_bench.py)
        int bn = 0, sp = 0;
        if (sscanf(f, "%d,%d", &bn, &sp) == 2 && (bn == 64 || bn == 128 || bn == 256) && sp >= 1 && (sp == 1 || (splitk_ws && !m_live))) {
            const int nkb = (K + BK - 1) / BK;
            pl.bn = bn; pl.kb_per_split = (nkb + sp - 1) / sp; pl.splits = (nkb + pl.kb_per_split - 1) / pl.kb_per_split;
        }
    }
    if (pl.splits > 1 && (int64_t)pl.splits * M * N > splitk_ws_floats) {
        pl.splits = 1; pl.kb_per_split = (K + BK - 1) / BK;
    }
    const int tiles_m = (M + BM - 1) / BM;
    // pairing pays on long-K contractions (deeper ring: 6 stages of 32 KB instead of 4 of 48 KB, half the B traffic per SM);
    // the K = 512 shapes are epilogue / store bound and run slightly better on independent CTAs
    int mode = (pl.bn >= 128 && tiles_m >= 2) ? pair_mode() : 0;
    if (mode == 2 && !g_pair_forced && (K + BK - 1) / BK < 16) mode = 0;
    // cluster shape of the multicast kernel (mode 1: pairs along M sharing B; small-M contractions: cm x cn, see cluster_shape)
    int cm = mode ? 2 : 1, cn = 1;
    if (mode != 2 || !g_pair_forced) {
        int qm = 1, qn = 1;
        cluster_shape(M, N, K, pl.bn, pl.splits, m_live != nullptr, &qm, &qn);
        if (qm * qn > 1) { mode = 1; cm = qm; cn = qn; }
    }
    // fp32-grade tier with three stored planes per operand (icd_gemm_x3_planes): K = 6 segments are walked, 3 are stored
    const int x3a = g_x3_planes[0], x3b = g_x3_planes[1];
    ICD_CHECK_ARG((!x3a || (int64_t)x3a * 6 * BK == K) && (!x3b || (int64_t)x3b * 6 * BK == K),
                  "gemm_tc: three-plane operands need K = 6 * 64 * segment k-blocks (K=%d, %d / %d)", K, x3a, x3b);
    const int Ka = x3a ? K / 2 : K, Kb = x3b ? K / 2 : K;
    CUtensorMap tmA, tmB;
    if (mode == 2) {
        ICD_TRY(make_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(A16), lda, M, Ka, BM, a_mn));
        ICD_TRY(make_tmap(&tmB, reinterpret_cast<const __nv_bfloat16*>(B16), ldb, N, Kb, pl.bn / 2, b_mn));
    } else {
        // K-major: the box is the CTA's row slice of the tile; MN-major: all 64-wide MN blocks, the CTA's k-row slice
        ICD_TRY(make_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(A16), lda, M, Ka, BM / cn, a_mn, BK / cn));
        ICD_TRY(make_tmap(&tmB, reinterpret_cast<const __nv_bfloat16*>(B16), ldb, N, Kb, pl.bn / cm, b_mn, BK / cm));
    }
    const int cluster = mode == 2 ? 2 : cm;                    // CTAs along M that form one super tile
    static const bool verbose = getenv("ICD_GEMM_VERBOSE") != nullptr;
    if (verbose) fprintf(stderr, "[icd] gemm %d x %d x %d (a_mn %d b_mn %d): bn %d, splits %d x %d k-blocks, %s, cluster %d x %d\n", M, N, K,
                         a_mn, b_mn, pl.bn, pl.splits, pl.kb_per_split, mode == 2 ? "cta_group::2 pairs" : "cta_group::1", cm, cn);
    KArgs k = {};
    EpiArgs& e = k.e;
    e.C = C; e.ldc = ldc; e.M = M; e.N = N; e.K = K; e.bias1 = bias1; e.bias2 = bias2;
    e.add1 = add1; e.ld1 = ld1; e.add2 = add2; e.ld2 = ld2; e.row_mask = row_mask; e.beta = beta;
    e.C16 = reinterpret_cast<__nv_bfloat16*>(C16); e.ldc16 = ldc16;
    auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    auto fits = [&](int v) {     // v floats per access: fp32 rows aligned to 4v bytes, bf16 rows to 2v bytes
        return (!C || (al(C, 4 * v) && ldc % v == 0)) && (!add1 || (al(add1, 4 * v) && ld1 % v == 0)) &&
               (!add2 || (al(add2, 4 * v) && ld2 % v == 0)) && (!C16 || (al(C16, 2 * v) && ldc16 % v == 0)) &&
               (!bias1 || al(bias1, 4 * v)) && (!bias2 || al(bias2, 4 * v));
    };
    const bool shifted = C && !C16 && !add1 && !add2 && beta == 0.f && al(C, 16) && (ldc & 3) == 2;   // see epilogue_store_chunk
    e.vec = fits(4) ? 4 : (shifted ? 3 : (fits(2) ? 2 : 1));
    k.a_mn = a_mn ? 1 : 0; k.b_mn = b_mn ? 1 : 0;
    k.splits = pl.splits; k.kb_per_split = pl.kb_per_split; k.partial = splitk_ws;
    k.cm = mode == 2 ? 2 : cm; k.cn = mode == 2 ? 1 : cn;
    k.m_live = m_live;
    k.x3_kb_a = x3a; k.x3_kb_b = x3b;
    k.late_trigger = icd_gemm_take_late_hint();
    {   // rasterisation of the work units (ICD_GEMM_RASTER=0|1 overrides): see KArgs::raster_n
        static const int raster_env = [] { const char* f = getenv("ICD_GEMM_RASTER"); return f ? atoi(f) : -1; }();
        // default: along N when the operand with the long dimension is A (>= 8 row tiles, several column tiles): the column
        // tiles of one row block then run in the same wave, so the A block is fetched from HBM once (K1: 411 MB of features read
        // once instead of once per column tile, 178 -> 157 us) and a wave writes whole output rows
        k.raster_n = raster_env >= 0 ? (raster_env != 0) : (tiles_m >= 8 && N > pl.bn && m_live == nullptr);
    }
    const int tiles_n = (N + pl.bn - 1) / pl.bn;
    const int tiles = ((tiles_m + cluster - 1) / cluster) * ((tiles_n + k.cn - 1) / k.cn);  // super tiles when clustered
    const int units = tiles * pl.splits;
    if (mode == 2) {
        if (pl.bn == 256) ICD_TRY(launch2<256>(tmA, tmB, k, units, s)); else ICD_TRY(launch2<128>(tmA, tmB, k, units, s));
    } else {
        switch (pl.bn) {
            case 256: ICD_TRY(launch<256>(tmA, tmB, k, units, s)); break;
            case 128: ICD_TRY(launch<128>(tmA, tmB, k, units, s)); break;
            default:  ICD_TRY(launch<64>(tmA, tmB, k, units, s)); break;
        }
    }
    if (pl.splits > 1 && deferred_splits) {
        *deferred_splits = pl.splits;          // the consumer sums the [splits][M][N] planes left in splitk_ws
    } else if (pl.splits > 1) {
        const long long work = ((long long)M * N + 3) / 4;
        long long blocks = (work + 255) / 256;
        if (blocks > ICD_NUM_SMS * 8) blocks = ICD_NUM_SMS * 8;
        ICD_CUDA(icd_launch_pdl(ICD_PDL_REDUCE, splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, s,
                                (const float*)splitk_ws, pl.splits, e));
        ICD_LAUNCH_CHECK();
    }
    return 0;
}

// C[M,N] = sum of the `splits` planes a deferred contraction left in splitk_ws (fixed order); optional bf16 copy
int icd_splitk_finish(const float* splitk_ws, int splits, int M, int N, float* C, int64_t ldc, void* C16, int64_t ldc16,
                      cudaStream_t s) {
    if (M == 0 || N == 0 || splits <= 0) return 0;
    EpiArgs e;
    e.C = C; e.ldc = ldc; e.M = M; e.N = N; e.K = 0; e.bias1 = e.bias2 = e.add1 = e.add2 = nullptr; e.ld1 = e.ld2 = 0;
    e.row_mask = nullptr; e.beta = 0.f; e.C16 = reinterpret_cast<__nv_bfloat16*>(C16); e.ldc16 = ldc16;
    auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    e.vec = ((!C || (al(C, 16) && ldc % 4 == 0)) && (!C16 || (al(C16, 8) && ldc16 % 4 == 0)) && N % 4 == 0) ? 4 : 1;
    const long long work = ((long long)M * N + 3) / 4;
    long long blocks = (work + 255) / 256;
    if (blocks > ICD_NUM_SMS * 8) blocks = ICD_NUM_SMS * 8;
    ICD_CUDA(icd_launch_pdl(ICD_PDL_REDUCE, splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, s, splitk_ws, splits, e));
    ICD_LAUNCH_CHECK();
    return 0;
}

int icd_gemm_bf16(const void* A16, int64_t lda, const void* B16, int64_t ldb, float* C, int64_t ldc,
                  int M, int N, int K, const float* bias1, const float* bias2, const float* add1, int64_t ld1,
                  const float* add2, int64_t ld2, const uint8_t* row_mask, float beta, cudaStream_t s,
                  void* C16, int64_t ldc16) {
    return icd_gemm_bf16_ex(A16, lda, 0, B16, ldb, 0, C, ldc, M, N, K, bias1, bias2, add1, ld1, add2, ld2, row_mask,
                            beta, s, C16, ldc16, nullptr, 0, nullptr);
}

namespace {
inline int64_t up8(int64_t x) { return (x + 7) / 8 * 8; }
inline int64_t up256(int64_t x) { return (x + 255) / 256 * 256; }
}

// workspace of the public fp32-in entry point: bf16 copies of both operands (in their own orientation) + split-K slices
int64_t icd_gemm_tc_ws_bytes(int M, int N, int K) {
    const int64_t a = up256(std::max((int64_t)M * up8(K), (int64_t)K * up8(M)) * 2);
    const int64_t b = up256(std::max((int64_t)N * up8(K), (int64_t)K * up8(N)) * 2);
    return a + b + up256(icd_gemm_bf16_splitk_floats(M, N, K) * 4);
}

int icd_gemm_tc_launch(const icd_gemm_desc_t* d, cudaStream_t s) {
    if (d->M == 0 || d->N == 0) return 0;
    ICD_CHECK_ARG(d->sak == 1 || d->sam == 1, "gemm: A needs a unit stride");
    ICD_CHECK_ARG(d->sbk == 1 || d->sbn == 1, "gemm: B needs a unit stride");
    const int64_t need = icd_gemm_tc_ws_bytes(d->M, d->N, d->K);
    ICD_CHECK_ARG(d->ws && d->ws_bytes >= need, "gemm: ICD_PREC_BF16 needs %lld bytes of workspace (icd_gemm_ws_bytes), got %lld",
                  (long long)need, (long long)d->ws_bytes);
    const int M = d->M, N = d->N, K = d->K;
    const int a_mn = (d->sak != 1), b_mn = (d->sbk != 1);          // row-contiguous source => MN-major operand, no transpose
    const int64_t a_bytes = up256(std::max((int64_t)M * up8(K), (int64_t)K * up8(M)) * 2);
    const int64_t b_bytes = up256(std::max((int64_t)N * up8(K), (int64_t)K * up8(N)) * 2);
    char* a16 = reinterpret_cast<char*>(d->ws);
    char* b16 = a16 + a_bytes;
    float* sk = reinterpret_cast<float*>(b16 + b_bytes);
    const int64_t lda = a_mn ? up8(M) : up8(K), ldb = b_mn ? up8(N) : up8(K);
    if (!a_mn) ICD_TRY(icd_convert_bf16(d->A, d->sam, 1, M, K, a16, lda, s));
    else       ICD_TRY(icd_convert_bf16(d->A, d->sak, 1, K, M, a16, lda, s));
    if (!b_mn) ICD_TRY(icd_convert_bf16(d->B, d->sbn, 1, N, K, b16, ldb, s));
    else       ICD_TRY(icd_convert_bf16(d->B, d->sbk, 1, K, N, b16, ldb, s));
    return icd_gemm_bf16_ex(a16, lda, a_mn, b16, ldb, b_mn, d->C, d->ldc, M, N, K, d->bias1, d->bias2, d->add1, d->ld1,
                            d->add2, d->ld2, d->row_mask, d->beta, s, nullptr, 0, sk,
                            icd_gemm_bf16_splitk_floats(M, N, K));
}

// ---------------------------------------------------------------------------------------------- fp32-grade (3-term) tier
// K-major operand [rows][K] -> bf16 [rows][6*seg], seg = up8(K); returns the leading dimension.
int icd_split3_bf16(const float* src, int64_t s_r, int rows, int cols, void* dst, int which, cudaStream_t s, const int* m_live) {
    if (rows == 0 || cols == 0) return 0;
    const int seg = (int)up8(cols);
    const int vec8 = (cols % 8 == 0) && (s_r % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const long long total = vec8 ? (long long)rows * (seg / 8) : (long long)rows * seg;
    long long blocks = (total + 255) / 256;
    if (blocks > ICD_NUM_SMS * 16) blocks = ICD_NUM_SMS * 16;
    split3_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, s_r, rows, cols, seg, reinterpret_cast<__nv_bfloat16*>(dst),
                                                        (which == 2 ? 3LL : 6LL) * seg, which, m_live, vec8);
    ICD_LAUNCH_CHECK();
    return 0;
}

int64_t icd_gemm_x3_ws_bytes(int M, int N, int K) {
    const int64_t a = up256((int64_t)up8(M) * 6 * up8(K) * 2);
    const int64_t b = up256((int64_t)up8(N) * 6 * up8(K) * 2);
    return a + b + up256(std::max(icd_gemm_bf16_splitk_floats(M, N, 6 * (int)up8(K)), icd_gemm_bf16_splitk_floats(M, N, 6 * K)) * 4);
}

int icd_gemm_x3_launch(const icd_gemm_desc_t* d, cudaStream_t s) {
    if (d->M == 0 || d->N == 0) return 0;
    ICD_CHECK_ARG(d->sak == 1 || d->sam == 1, "gemm: A needs a unit stride");
    ICD_CHECK_ARG(d->sbk == 1 || d->sbn == 1, "gemm: B needs a unit stride");
    const int M = d->M, N = d->N, K = d->K;
    const int64_t need = icd_gemm_x3_ws_bytes(M, N, K);
    ICD_CHECK_ARG(d->ws && d->ws_bytes >= need, "gemm: ICD_PREC_FP32X3 needs %lld bytes of workspace (icd_gemm_ws_bytes), got %lld",
                  (long long)need, (long long)d->ws_bytes);
    const int a_mn = (d->sak != 1), b_mn = (d->sbk != 1);
    const int64_t a_bytes = up256((int64_t)up8(M) * 6 * up8(K) * 2);
    const int64_t b_bytes = up256((int64_t)up8(N) * 6 * up8(K) * 2);
    char* a16 = reinterpret_cast<char*>(d->ws);
    char* b16 = a16 + a_bytes;
    float* sk = reinterpret_cast<float*>(b16 + b_bytes);
    const int seg = (int)up8(K);
    // both operands must agree on the K' layout: K-major segments are padded to seg = up8(K) columns; an MN-major operand
    // uses K rows per segment when its partner is MN-major too, and seg rows (pad rows zeroed) when the majors differ
    const bool mixed = (a_mn != b_mn);
    const int mn_seg = mixed ? seg : K;
    const int Kx = (a_mn && b_mn) ? 6 * K : 6 * seg;
    // Three stored planes per operand instead of six K-segments whenever a segment is a whole number of k-blocks: the split
    // passes write half the bytes, the contraction is the same sequence of MMAs on the same values (bit-identical results).
    const int seg_len = (a_mn && b_mn) ? K : seg;                 // elements of K per segment, the same for both operands
    const bool p3 = (seg_len % BK) == 0 && getenv("ICD_X3_PLANES6") == nullptr;
    const int nseg = p3 ? 3 : 6;
    int64_t lda, ldb;
    auto split_mn = [&](const float* src, int64_t s_k, int mn, void* dst, int64_t ldd, int which) -> int {
        if (p3) which = 2;
        if (mn_seg != K) ICD_CUDA(cudaMemsetAsync(dst, 0, (size_t)nseg * mn_seg * ldd * 2, s));
        const long long total = (long long)K * mn;
        long long blocks = (total + 255) / 256;
        if (blocks > ICD_NUM_SMS * 16) blocks = ICD_NUM_SMS * 16;
        split3_mn_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, s_k, K, mn, mn_seg, reinterpret_cast<__nv_bfloat16*>(dst), ldd, which);
        ICD_LAUNCH_CHECK();
        return 0;
    };
    // an operand that stays constant during the enclosing entry-point call (a marked weight matrix) is split once and reused
    bool a_fresh = true, b_fresh = true;
    if (void* c = icd_x3_cache_lookup(d->A, a_mn ? d->sak : d->sam, M, K, a_mn, a_mn ? mn_seg : seg, p3 ? 2 : 0, a_bytes, &a_fresh)) a16 = reinterpret_cast<char*>(c);
    if (void* c = icd_x3_cache_lookup(d->B, b_mn ? d->sbk : d->sbn, N, K, b_mn, b_mn ? mn_seg : seg, p3 ? 2 : 1, b_bytes, &b_fresh)) b16 = reinterpret_cast<char*>(c);
    if (!a_mn) { lda = (int64_t)nseg * seg; if (a_fresh) ICD_TRY(icd_split3_bf16(d->A, d->sam, M, K, a16, p3 ? 2 : 0, s)); }
    else { lda = up8(M); if (a_fresh) ICD_TRY(split_mn(d->A, d->sak, M, a16, lda, 0)); }
    if (!b_mn) { ldb = (int64_t)nseg * seg; if (b_fresh) ICD_TRY(icd_split3_bf16(d->B, d->sbn, N, K, b16, p3 ? 2 : 1, s)); }
    else { ldb = up8(N); if (b_fresh) ICD_TRY(split_mn(d->B, d->sbk, N, b16, ldb, 1)); }
    if (p3) icd_gemm_x3_planes(seg_len / BK, seg_len / BK);
    const int rc = icd_gemm_bf16_ex(a16, lda, a_mn, b16, ldb, b_mn, d->C, d->ldc, M, N, Kx, d->bias1, d->bias2, d->add1, d->ld1,
                                    d->add2, d->ld2, d->row_mask, d->beta, s, nullptr, 0, sk,
                                    icd_gemm_bf16_splitk_floats(M, N, Kx), nullptr);
    icd_gemm_x3_planes(0, 0);
    return rc;
}
