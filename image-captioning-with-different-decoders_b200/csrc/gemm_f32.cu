// fp32 (parity-tier) dense contraction  C[M,N] = A[M,K] * B[N,K]^T + epilogue.
//
// CUDA-core FMA kernel: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered
// shared memory with register prefetch, 128-bit global and shared accesses, optional split-K
// (atomic accumulation) for the weight-gradient contractions whose M x N is small and K = B*T or B*196.
// This is the precision reference tier: the reference runs its Linear / LSTMCell matmuls in plain fp32
// (TF32 off by default) and identical greedy / beam captions need fp32-grade logits.
// The bf16 tcgen05/TMA tier lives in gemm_tc.cu.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS = BM + 4;

struct GemmArgs {
    const float* A; long long sam, sak;
    const float* B; long long sbn, sbk;
    float* C; long long ldc;
    int M, N, K;
    int kper;                 // K range per z-slice (multiple of BK)
    int vecA, vecB, vecC;
    const float* bias1; const float* bias2;
    const float* add1; long long ld1;
    const float* add2; long long ld2;
    const unsigned char* row_mask;
    float beta;
};

// Load one BKxB{M,N} operand tile into registers (2 float4 per thread).
//  KC = true : operand is K-contiguous  (X(r,k) = X[r*sr + k]);   f -> row = f>>2, kvec = f&3
//  KC = false: operand is row-contiguous (X(r,k) = X[k*sk + r]);  f -> k = f>>5,  rvec = f&31
template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ X, long long sr, long long sk,
                                          int r0, int R, int k0, int kend, int vec, float4 (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int f = threadIdx.x + i * 256;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KC) {
            const int r = r0 + (f >> 2), k = k0 + (f & 3) * 4;
            if (r < R) {
                const float* p = X + (long long)r * sr + k;
                if (vec && k + 3 < kend) t = *reinterpret_cast<const float4*>(p);
                else {
                    if (k + 0 < kend) t.x = p[0];
                    if (k + 1 < kend) t.y = p[1];
                    if (k + 2 < kend) t.z = p[2];
                    if (k + 3 < kend) t.w = p[3];
                }
            }
        } else {
            const int k = k0 + (f >> 5), r = r0 + (f & 31) * 4;
            if (k < kend) {
                const float* p = X + (long long)k * sk + r;
                if (vec && r + 3 < R) t = *reinterpret_cast<const float4*>(p);
                else {
                    if (r + 0 < R) t.x = p[0];
                    if (r + 1 < R) t.y = p[1];
                    if (r + 2 < R) t.z = p[2];
                    if (r + 3 < R) t.w = p[3];
                }
            }
        }
        v[i] = t;
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[LDS], const float4 (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int f = threadIdx.x + i * 256;
        if (KC) {
            const int r = f >> 2, k = (f & 3) * 4;
            S[k + 0][r] = v[i].x; S[k + 1][r] = v[i].y; S[k + 2][r] = v[i].z; S[k + 3][r] = v[i].w;
        } else {
            const int k = f >> 5, r = (f & 31) * 4;
            *reinterpret_cast<float4*>(&S[k][r]) = v[i];
        }
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2) gemm_f32_kernel(const GemmArgs p) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];

    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.kper;
    const int kend = min(p.K, kbeg + p.kper);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    const long long a_sr = A_KC ? p.sam : 1, a_sk = A_KC ? 1 : p.sak;
    const long long b_sr = B_KC ? p.sbn : 1, b_sk = B_KC ? 1 : p.sbk;

    int buf = 0;
    if (kbeg < kend) {
        load_tile<A_KC>(p.A, a_sr, a_sk, m0, p.M, kbeg, kend, p.vecA, ra);
        load_tile<B_KC>(p.B, b_sr, b_sk, n0, p.N, kbeg, kend, p.vecB, rb);
        store_tile<A_KC>(As[0], ra);
        store_tile<B_KC>(Bs[0], rb);
    }
    __syncthreads();

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool has_next = (k0 + BK) < kend;
        if (has_next) {
            load_tile<A_KC>(p.A, a_sr, a_sk, m0, p.M, k0 + BK, kend, p.vecA, ra);
            load_tile<B_KC>(p.B, b_sr, b_sk, n0, p.N, k0 + BK, kend, p.vecB, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            store_tile<A_KC>(As[buf ^ 1], ra);
            store_tile<B_KC>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue ----
    const bool split = gridDim.z > 1;
    const bool lead = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        const bool masked = p.row_mask && !p.row_mask[m];
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jh * 4 + j];
                const int nn = n + j;
                if (nn < p.N && lead) {
                    if (p.bias1) x += p.bias1[nn];
                    if (p.bias2) x += p.bias2[nn];
                    if (p.add1) x += p.add1[(long long)m * p.ld1 + nn];
                    if (p.add2) x += p.add2[(long long)m * p.ld2 + nn];
                }
                v[j] = masked ? 0.f : x;
            }
            float* c = p.C + (long long)m * p.ldc + n;
            if (split) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < p.N) atomicAdd(c + j, v[j]);
            } else if (p.vecC && n + 3 < p.N) {
                float4 o = make_float4(v[0], v[1], v[2], v[3]);
                if (p.beta != 0.f) {
                    const float4 old = *reinterpret_cast<const float4*>(c);
                    o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                }
                *reinterpret_cast<float4*>(c) = o;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) c[j] = (p.beta != 0.f) ? v[j] + p.beta * c[j] : v[j];
            }
        }
    }
}

__global__ void zero_rows_kernel(float* C, long long ldc, int M, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)M * N) C[(i / N) * ldc + (i % N)] = 0.f;
}

}  // namespace

int icd_gemm_f32_launch(const icd_gemm_desc_t* d, cudaStream_t s) {
    ICD_CHECK_ARG(d->M >= 0 && d->N >= 0 && d->K >= 0, "gemm: negative dimension");
    if (d->M == 0 || d->N == 0) return 0;
    ICD_CHECK_ARG(d->sak == 1 || d->sam == 1, "gemm: A needs a unit stride (sam=%lld sak=%lld)",
                  (long long)d->sam, (long long)d->sak);
    ICD_CHECK_ARG(d->sbk == 1 || d->sbn == 1, "gemm: B needs a unit stride (sbn=%lld sbk=%lld)",
                  (long long)d->sbn, (long long)d->sbk);
    GemmArgs p;
    p.A = d->A; p.sam = d->sam; p.sak = d->sak;
    p.B = d->B; p.sbn = d->sbn; p.sbk = d->sbk;
    p.C = d->C; p.ldc = d->ldc; p.M = d->M; p.N = d->N; p.K = d->K;
    p.bias1 = d->bias1; p.bias2 = d->bias2; p.add1 = d->add1; p.ld1 = d->ld1; p.add2 = d->add2; p.ld2 = d->ld2;
    p.row_mask = d->row_mask; p.beta = d->beta;
    const bool a_kc = (d->sak == 1), b_kc = (d->sbk == 1);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vecA = al16(d->A) && ((a_kc ? d->sam : d->sak) % 4 == 0);
    p.vecB = al16(d->B) && ((b_kc ? d->sbn : d->sbk) % 4 == 0);
    p.vecC = al16(d->C) && (d->ldc % 4 == 0);

    const int gm = (d->M + BM - 1) / BM, gn = (d->N + BN - 1) / BN;
    int splitk = 1;
    const long long tiles = (long long)gm * gn;
    if ((d->flags & ICD_GEMM_ALLOW_SPLITK) && d->beta == 0.f && d->K >= 2048 && tiles < ICD_NUM_SMS) {
        splitk = (int)((2 * ICD_NUM_SMS + tiles - 1) / tiles);
        const int maxs = d->K / 512;
        if (splitk > maxs) splitk = maxs;
        if (splitk < 1) splitk = 1;
    }
    int kper = (d->K + splitk - 1) / splitk;
    kper = ((kper + BK - 1) / BK) * BK;
    if (kper == 0) kper = BK;
    splitk = d->K > 0 ? (d->K + kper - 1) / kper : 1;
    p.kper = kper;
    if (splitk > 1) {   // atomics accumulate into C: clear it first
        const long long n = (long long)d->M * d->N;
        zero_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d->C, d->ldc, d->M, d->N);
        ICD_LAUNCH_CHECK();
    }
    dim3 grid(gn, gm, splitk);
    if (a_kc && b_kc)        gemm_f32_kernel<true, true><<<grid, 256, 0, s>>>(p);
    else if (a_kc && !b_kc)  gemm_f32_kernel<true, false><<<grid, 256, 0, s>>>(p);
    else if (!a_kc && b_kc)  gemm_f32_kernel<false, true><<<grid, 256, 0, s>>>(p);
    else                     gemm_f32_kernel<false, false><<<grid, 256, 0, s>>>(p);
    ICD_LAUNCH_CHECK();
    return 0;
}
