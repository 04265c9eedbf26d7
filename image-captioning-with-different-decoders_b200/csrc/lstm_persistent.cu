// K8 — the recurrent part of the baseline LSTM decoder (nn.LSTM, models/baseline.py:106) as ONE persistent kernel per
// direction, driven by a grid barrier instead of a launch chain (SURVEY.md 2.3 row `baseline.py:106`, 7.1 step 6).
//
// Forward, step t:   gates = h_{t-1} W_hh^T + xg_t   (xg = x W_ih^T + b_ih + b_hh, hoisted over all steps by the caller)
//                    i,f,g,o -> c_t = f c_{t-1} + i g,  h_t = o tanh(c_t)          (gate order i,f,g,o)
// Backward, step t:  dh_t = dg_{t+1} W_hh + d_hout_t ;  LSTM cell adjoint -> dg_t, dc_{t-1}
//
// Decomposition (weight-stationary — it pays here because the per-step activation is only B x H):
//   * the grid is H/4 CTAs (forward: 4 hidden units = 16 gate columns each) or H/16 CTAs (backward: 16 hidden units each);
//     a CTA keeps ITS slice of W_hh (16 rows of the B operand x K, bf16, 128B-swizzled K-major UMMA tiles) in shared memory
//     for all L steps — converted from the fp32 weight once, by the kernel itself;
//   * per step the whole CTA grid streams the same small A operand (h_{t-1} or dg_{t+1}, bf16, B x K) through a TMA ring and
//     issues tcgen05.mma (M = 128 rows per tile, N = 16, K = 16 per instruction) into a TMEM accumulator;
//   * the epilogue warps read the accumulator (tcgen05.ld), apply the LSTM cell math for their row and write the new state —
//     including the bf16 copy that is the NEXT step's A operand — straight to global memory;
//   * steps are separated by a grid-wide barrier (one atomic counter; cooperative launch guarantees co-residency), with the
//     generic->async proxy fences the TMA reads of freshly written rows require.
// No per-step launches, no gates_pre round trip, no separate gate-math kernel.
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace {

constexpr int LP_THREADS = 192;           // warp 0: TMA producer + grid barrier, warp 1: TMEM alloc + MMA issue, warps 2-5: epilogue
constexpr int LP_BM = 128;
constexpr int LP_N = 16;                  // accumulator columns per CTA
constexpr int LP_A_BYTES = LP_BM * TC_BK * 2;       // 16 KB per stage
constexpr int LP_W_KB_BYTES = LP_N * TC_BK * 2;     // 2 KB of the resident weight slice per k-block
constexpr int LP_TMEM_COLS = 64;          // up to 4 row tiles of 16 columns
constexpr int LP_MAX_MT = LP_TMEM_COLS / LP_N;
// cute::UMMA::InstrDescriptor: c F32, a/b BF16, both K-major, N = 16, M = 128
constexpr uint32_t LP_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LP_N >> 3) << 17) | ((uint32_t)(LP_BM >> 4) << 24);

struct LstmSeqArgs {
    int B, L, H, K, nkb, stages;
    const float* w_hh;                    // (4H, H) fp32
    // forward
    const float* xg;                      // (L,B,4H)
    float* gates_act;                     // (L,B,4H)
    float* c_all;                         // (L+1,B,H)
    float* h_all;                         // (L+1,B,H)
    float* hout;                          // (B,L,H)
    __nv_bfloat16* h16;                   // ((L+1)*B, H): row block t = h_t, the A operand of step t
    __nv_bfloat16* hout16;                // (B*L, H)
    // backward
    const float* d_hout;                  // (B,L,H)
    float* dc;                            // (B,H) running cell-state gradient (zero on entry)
    float* dg;                            // (L,B,4H)
    __nv_bfloat16* dg16;                  // (L*B, 4H): row block t = dg_t, the A operand of step t-1
    unsigned int* bar;                    // grid barrier counter (zero on entry)
};

__device__ __forceinline__ void tc_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of element (n, k) of the resident weight slice: k-blocks of 16 rows x 128 B, 8-row swizzle atoms of 1024 B,
// 16-byte chunk index XOR row-in-atom (the SWIZZLE_128B pattern the UMMA descriptor expects of a K-major tile)
__device__ __forceinline__ uint32_t w_slice_off(int n, int k) {
    const int kb = k >> 6, kk = k & 63;
    return (uint32_t)(kb * LP_W_KB_BYTES + (n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2);
}

__device__ __forceinline__ void grid_arrive(unsigned int* bar) {
    __threadfence();
    atomicAdd(bar, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned int* bar, unsigned int target) {
    unsigned int v;
    unsigned long long spins = 0;
    while (true) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v >= target) break;
        if (++spins > (1ull << 26)) __trap();          // a lost CTA must fault, never hang the GPU
    }
    __threadfence();
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    return pk;
}

// BWD = 0: forward recurrence, CTA c owns hidden units [4c, 4c+4)   (accumulator column n = gate * 4 + unit)
// BWD = 1: backward recurrence, CTA c owns hidden units [16c, 16c+16) (accumulator column n = unit)
template <int BWD>
__global__ void __launch_bounds__(LP_THREADS, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmA, const LstmSeqArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int S = p.stages;
    const uint32_t sA = base;                                          // S stages x 16 KB
    const uint32_t sW = base + (uint32_t)S * LP_A_BYTES;               // nkb x 2 KB, resident
    uint8_t* gW = gbase + (size_t)S * LP_A_BYTES;
    const uint32_t bars = sW + (uint32_t)p.nkb * LP_W_KB_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * S, tfull = bars + 16 * S, tempty = tfull + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gW + (size_t)p.nkb * LP_W_KB_BYTES + 16 * S + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = p.B, L = p.L, H = p.H, K = p.K;
    const int MT = (B + LP_BM - 1) / LP_BM;
    const int unit0 = blockIdx.x * (BWD ? 16 : 4);
    const unsigned int NC = gridDim.x;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)LP_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // resident weight slice, fp32 -> bf16, written in the swizzled UMMA layout (zero beyond K)
    {
        const int Kpad = p.nkb * TC_BK;
        for (int idx = threadIdx.x; idx < LP_N * Kpad; idx += LP_THREADS) {
            int n, k;
            float v = 0.f;
            if (!BWD) {                                    // B(n, k) = W_hh[gate * H + unit0 + u][k],  n = gate * 4 + u
                n = idx / Kpad; k = idx % Kpad;
                if (k < K) v = p.w_hh[(size_t)((n >> 2) * H + unit0 + (n & 3)) * H + k];
            } else {                                       // B(n, k) = W_hh[k][unit0 + n]   (dh = dg W_hh)
                k = idx / LP_N; n = idx % LP_N;
                if (k < K) v = p.w_hh[(size_t)k * H + unit0 + n];
            }
            *reinterpret_cast<__nv_bfloat16*>(gW + w_slice_off(n, k)) = __float2bfloat16_rn(v);
        }
    }
    fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int stage = 0; uint32_t phase = 0;                     // A ring (producer and MMA warps keep their own copies)
    uint32_t acc_phase = 0;                                // accumulator hand-off (MMA and epilogue warps)

    for (int it = 0; it < L; ++it) {
        const int t = BWD ? (L - 1 - it) : it;
        const bool has_gemm = it > 0;                      // forward: h_0 = 0; backward: no dh flows in from beyond the last step
        const int a_row0 = (BWD ? (t + 1) : t) * B;        // first row of this step's A operand (h_t resp. dg_{t+1})
        if (warp == 0) {
            // =================================== TMA producer + grid barrier ===================================
            if (lane == 0 && has_gemm) {
                grid_wait(p.bar, NC * (unsigned int)it);   // every CTA has published its rows of iteration it-1
                fence_proxy_async_all();                   // ... and the async proxy (TMA) may now read them
                for (int mt = 0; mt < MT; ++mt)
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        mbar_arrive_expect_tx(full0 + 8 * stage, LP_A_BYTES);
                        tma_load_2d(sA + stage * LP_A_BYTES, &tmA, kb * TC_BK, a_row0 + mt * LP_BM, full0 + 8 * stage);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
            }
        } else if (warp == 1) {
            // =================================== MMA issuer ===================================
            if (has_gemm) {
                mbar_wait(tempty, acc_phase ^ 1);          // the epilogue has drained the previous step's accumulator
                tc_fence_after();
                for (int mt = 0; mt < MT; ++mt)
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t adesc = make_smem_desc(sA + stage * LP_A_BYTES, 0);
                            const uint64_t bdesc = make_smem_desc(sW + kb * LP_W_KB_BYTES, 0);
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k)
                                tc_mma_f16(tmem_base + (uint32_t)(mt * LP_N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                           LP_IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                            tc_commit(empty0 + 8 * stage);
                            if (mt == MT - 1 && kb == p.nkb - 1) tc_commit(tfull);
                        }
                        __syncwarp();
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                acc_phase ^= 1;
            }
        } else {
            // =================================== epilogue: LSTM cell math, warps 2..5 ===================================
            const int q = warp & 3;                        // TMEM lane quarter this warp may access
            if (has_gemm) {
                mbar_wait(tfull, acc_phase);
                tc_fence_after();
            }
            for (int mt = 0; mt < MT; ++mt) {
                const int row = mt * LP_BM + q * 32 + lane;
                uint32_t v[16];
                if (has_gemm) {
                    tc_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * LP_N), v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (has_gemm && mt == MT - 1) {            // all TMEM reads of this warp are done: hand the accumulator back
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(tempty);
                }
                if (row >= B) continue;
                if (!BWD) {
                    const size_t g_off = ((size_t)t * B + row) * 4 * H + unit0;
                    const size_t s_off = (size_t)row * H + unit0;
                    const size_t BH = (size_t)B * H;
                    const float4 xi = *reinterpret_cast<const float4*>(p.xg + g_off);
                    const float4 xf = *reinterpret_cast<const float4*>(p.xg + g_off + H);
                    const float4 xc = *reinterpret_cast<const float4*>(p.xg + g_off + 2 * H);
                    const float4 xo = *reinterpret_cast<const float4*>(p.xg + g_off + 3 * H);
                    const float4 cp = *reinterpret_cast<const float4*>(p.c_all + (size_t)t * BH + s_off);
                    const float xi_[4] = {xi.x, xi.y, xi.z, xi.w}, xf_[4] = {xf.x, xf.y, xf.z, xf.w};
                    const float xc_[4] = {xc.x, xc.y, xc.z, xc.w}, xo_[4] = {xo.x, xo.y, xo.z, xo.w};
                    const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
                    float gi[4], gf[4], gg[4], go[4], cn[4], hn[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        gi[u] = sigmoidf_(__uint_as_float(v[u]) + xi_[u]);
                        gf[u] = sigmoidf_(__uint_as_float(v[4 + u]) + xf_[u]);
                        gg[u] = tanhf(__uint_as_float(v[8 + u]) + xc_[u]);
                        go[u] = sigmoidf_(__uint_as_float(v[12 + u]) + xo_[u]);
                        cn[u] = gf[u] * cp_[u] + gi[u] * gg[u];
                        hn[u] = go[u] * tanhf(cn[u]);
                    }
                    *reinterpret_cast<float4*>(p.gates_act + g_off) = make_float4(gi[0], gi[1], gi[2], gi[3]);
                    *reinterpret_cast<float4*>(p.gates_act + g_off + H) = make_float4(gf[0], gf[1], gf[2], gf[3]);
                    *reinterpret_cast<float4*>(p.gates_act + g_off + 2 * H) = make_float4(gg[0], gg[1], gg[2], gg[3]);
                    *reinterpret_cast<float4*>(p.gates_act + g_off + 3 * H) = make_float4(go[0], go[1], go[2], go[3]);
                    *reinterpret_cast<float4*>(p.c_all + (size_t)(t + 1) * BH + s_off) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    const float4 h4 = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(p.h_all + (size_t)(t + 1) * BH + s_off) = h4;
                    const uint2 hb = pack4_bf16(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<uint2*>(p.h16 + (size_t)(t + 1) * BH + s_off) = hb;            // next step's A operand
                    const size_t o_off = ((size_t)row * L + t) * H + unit0;
                    *reinterpret_cast<float4*>(p.hout + o_off) = h4;
                    if (p.hout16) *reinterpret_cast<uint2*>(p.hout16 + o_off) = hb;
                } else {
                    const size_t BH = (size_t)B * H;
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const int u0 = unit0 + 4 * g4;
                        const size_t g_off = ((size_t)t * B + row) * 4 * H + u0;
                        const size_t s_off = (size_t)row * H + u0;
                        const float4 dho = *reinterpret_cast<const float4*>(p.d_hout + ((size_t)row * L + t) * H + u0);
                        const float4 ai = *reinterpret_cast<const float4*>(p.gates_act + g_off);
                        const float4 af = *reinterpret_cast<const float4*>(p.gates_act + g_off + H);
                        const float4 ag = *reinterpret_cast<const float4*>(p.gates_act + g_off + 2 * H);
                        const float4 ao = *reinterpret_cast<const float4*>(p.gates_act + g_off + 3 * H);
                        const float4 c0 = *reinterpret_cast<const float4*>(p.c_all + (size_t)t * BH + s_off);
                        const float4 c1 = *reinterpret_cast<const float4*>(p.c_all + (size_t)(t + 1) * BH + s_off);
                        const float4 dcv = *reinterpret_cast<const float4*>(p.dc + s_off);
                        const float dh_[4] = {__uint_as_float(v[4 * g4]) + dho.x, __uint_as_float(v[4 * g4 + 1]) + dho.y,
                                              __uint_as_float(v[4 * g4 + 2]) + dho.z, __uint_as_float(v[4 * g4 + 3]) + dho.w};
                        const float i_[4] = {ai.x, ai.y, ai.z, ai.w}, f_[4] = {af.x, af.y, af.z, af.w};
                        const float g_[4] = {ag.x, ag.y, ag.z, ag.w}, o_[4] = {ao.x, ao.y, ao.z, ao.w};
                        const float cp_[4] = {c0.x, c0.y, c0.z, c0.w}, cn_[4] = {c1.x, c1.y, c1.z, c1.w};
                        const float dc_[4] = {dcv.x, dcv.y, dcv.z, dcv.w};
                        float pi[4], pf[4], pg[4], po[4], dcn[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float tc = tanhf(cn_[u]);
                            const float d_o = dh_[u] * tc;
                            const float dc = dc_[u] + dh_[u] * o_[u] * (1.f - tc * tc);
                            const float d_i = dc * g_[u], d_g = dc * i_[u], d_f = dc * cp_[u];
                            dcn[u] = dc * f_[u];
                            pi[u] = d_i * i_[u] * (1.f - i_[u]); pf[u] = d_f * f_[u] * (1.f - f_[u]);
                            pg[u] = d_g * (1.f - g_[u] * g_[u]); po[u] = d_o * o_[u] * (1.f - o_[u]);
                        }
                        *reinterpret_cast<float4*>(p.dc + s_off) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
                        *reinterpret_cast<float4*>(p.dg + g_off) = make_float4(pi[0], pi[1], pi[2], pi[3]);
                        *reinterpret_cast<float4*>(p.dg + g_off + H) = make_float4(pf[0], pf[1], pf[2], pf[3]);
                        *reinterpret_cast<float4*>(p.dg + g_off + 2 * H) = make_float4(pg[0], pg[1], pg[2], pg[3]);
                        *reinterpret_cast<float4*>(p.dg + g_off + 3 * H) = make_float4(po[0], po[1], po[2], po[3]);
                        *reinterpret_cast<uint2*>(p.dg16 + g_off) = pack4_bf16(pi[0], pi[1], pi[2], pi[3]);
                        *reinterpret_cast<uint2*>(p.dg16 + g_off + H) = pack4_bf16(pf[0], pf[1], pf[2], pf[3]);
                        *reinterpret_cast<uint2*>(p.dg16 + g_off + 2 * H) = pack4_bf16(pg[0], pg[1], pg[2], pg[3]);
                        *reinterpret_cast<uint2*>(p.dg16 + g_off + 3 * H) = pack4_bf16(po[0], po[1], po[2], po[3]);
                    }
                }
            }
            fence_proxy_async_all();                       // the rows just written are read by other CTAs' TMA next step
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + 1 < L) grid_arrive(p.bar);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)LP_TMEM_COLS) : "memory");
    }
}

int plan_smem(int nkb, int* stages) {
    // resident weight slice + barriers + as many 16 KB A stages as fit (2..8)
    const int fixed = 1024 + nkb * LP_W_KB_BYTES + 256;
    int s = (200 * 1024 - fixed) / LP_A_BYTES;
    if (s > 8) s = 8;
    *stages = s;
    return fixed + s * LP_A_BYTES;
}

bool persistent_enabled() {
    static const bool on = [] { const char* e = getenv("ICD_LSTM_PERSISTENT"); return !(e && e[0] == '0'); }();
    return on;
}

template <int BWD>
int launch_seq(const LstmSeqArgs& a0, const __nv_bfloat16* A16, int a_rows, cudaStream_t s) {
    LstmSeqArgs a = a0;
    a.nkb = (a.K + TC_BK - 1) / TC_BK;
    const int smem = plan_smem(a.nkb, &a.stages);
    ICD_CHECK_ARG(a.stages >= 2, "lstm_seq: hidden size %d too large for the resident weight slice", a.H);
    CUtensorMap tmA;
    ICD_TRY(make_tmap(&tmA, A16, a.K, a_rows, a.K, LP_BM, 0));
    static int smem_set = 0;
    if (smem_set < smem) {
        ICD_CUDA(cudaFuncSetAttribute(lstm_seq_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    ICD_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned int), s));
    const int grid = a.H / (BWD ? 16 : 4);
    void* args[] = {(void*)&tmA, (void*)&a};
    ICD_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_kernel<BWD>, dim3(grid), dim3(LP_THREADS), args, (size_t)smem, s));
    ICD_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// 1 if the persistent kernels cover this shape (else the caller keeps its launch chain): every CTA must be co-resident
// (grid <= #SMs), the accumulator of all row tiles must fit the TMEM allocation, H must split into whole unit groups.
int icd_lstm_seq_persistent_ok(int B, int L, int H) {
    if (!persistent_enabled()) return 0;
    if (B < 1 || L < 1 || H < 16 || H % 16 != 0) return 0;
    if (H / 4 > ICD_NUM_SMS) return 0;
    if ((B + LP_BM - 1) / LP_BM > LP_MAX_MT) return 0;
    int st;
    plan_smem((4 * H + TC_BK - 1) / TC_BK, &st);
    return st >= 2 ? 1 : 0;
}

int icd_lstm_seq_fwd_persistent(int B, int L, int H, const float* w_hh, const float* xg, float* gates_act, float* c_all,
                                float* h_all, float* hout, void* h16, void* hout16, unsigned int* bar, cudaStream_t s) {
    ICD_CHECK_ARG(icd_lstm_seq_persistent_ok(B, L, H), "lstm_seq_fwd: shape B=%d L=%d H=%d not covered by the persistent kernel", B, L, H);
    LstmSeqArgs a = {};
    a.B = B; a.L = L; a.H = H; a.K = H; a.w_hh = w_hh; a.xg = xg; a.gates_act = gates_act; a.c_all = c_all; a.h_all = h_all;
    a.hout = hout; a.h16 = reinterpret_cast<__nv_bfloat16*>(h16); a.hout16 = reinterpret_cast<__nv_bfloat16*>(hout16); a.bar = bar;
    return launch_seq<0>(a, a.h16, (L + 1) * B, s);
}

int icd_lstm_seq_bwd_persistent(int B, int L, int H, const float* w_hh, const float* d_hout, const float* gates_act,
                                const float* c_all, float* dc, float* dg, void* dg16, unsigned int* bar, cudaStream_t s) {
    ICD_CHECK_ARG(icd_lstm_seq_persistent_ok(B, L, H), "lstm_seq_bwd: shape B=%d L=%d H=%d not covered by the persistent kernel", B, L, H);
    LstmSeqArgs a = {};
    a.B = B; a.L = L; a.H = H; a.K = 4 * H; a.w_hh = w_hh; a.d_hout = d_hout; a.gates_act = const_cast<float*>(gates_act);
    a.c_all = const_cast<float*>(c_all); a.dc = dc; a.dg = dg; a.dg16 = reinterpret_cast<__nv_bfloat16*>(dg16); a.bar = bar;
    return launch_seq<1>(a, a.dg16, L * B, s);
}

// ---- C ABI (include/icd_b200.h): the LSTM recurrence on its own, for callers that hoist the input contraction themselves ----
extern "C" int icd_lstm_seq_supported(int B, int L, int H) { return icd_lstm_seq_persistent_ok(B, L, H); }

extern "C" int icd_lstm_seq_fwd(int B, int L, int H, const float* w_hh, const float* xg, float* gates_act, float* c_all,
                                float* h_all, float* hout, void* h16, void* hout16, void* barrier_ws, void* stream) {
    cudaStream_t s = icd_stream(stream);
    const size_t BH = (size_t)B * H;
    ICD_CUDA(cudaMemsetAsync(c_all, 0, sizeof(float) * BH, s));                    // zero (h0, c0): nn.LSTM default state
    ICD_CUDA(cudaMemsetAsync(h_all, 0, sizeof(float) * BH, s));
    ICD_CUDA(cudaMemsetAsync(h16, 0, 2 * BH, s));
    return icd_lstm_seq_fwd_persistent(B, L, H, w_hh, xg, gates_act, c_all, h_all, hout, h16, hout16,
                                       reinterpret_cast<unsigned int*>(barrier_ws), s);
}

extern "C" int icd_lstm_seq_bwd(int B, int L, int H, const float* w_hh, const float* d_hout, const float* gates_act,
                                const float* c_all, float* dc_ws, float* dg, void* dg16, void* barrier_ws, void* stream) {
    cudaStream_t s = icd_stream(stream);
    ICD_CUDA(cudaMemsetAsync(dc_ws, 0, sizeof(float) * (size_t)B * H, s));
    return icd_lstm_seq_bwd_persistent(B, L, H, w_hh, d_hout, gates_act, c_all, dc_ws, dg, dg16,
                                       reinterpret_cast<unsigned int*>(barrier_ws), s);
}
