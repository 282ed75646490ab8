// Range and Doppler FFT stages.
//
//   rs_range_fft   : cube[F][A][C][S] * table[S] -> FFT over S -> zero bin 0 (DC removal) ->
//                    range fftshift -> mid[F][S][A][C]           (dechirp.py:139,108,120,205,208,211)
//   rs_doppler_fft : mid[F][S][A][C] -> FFT over C -> Doppler fftshift -> rds[F][S][A][C]  (same row order)
//
// Both are shared-memory Stockham autosort kernels: a CTA stages a block of rows, runs the radix
// passes out of a host-built fp64-accurate twiddle table, and stores with the transpose fused so
// that the next stage reads contiguous memory.  Any length whose prime factors are <= 13 is
// supported (the reference's defaults are S=400, C=64; BASELINE configs are powers of two).
#include <algorithm>
#include <cstdlib>
#include "rs_common.cuh"
#include "rs_fft_pow2.cuh"

namespace {

struct FftPlan {
    int n;
    int npass;
    int radix[16];
};

static bool make_plan(int n, FftPlan* p) {
    p->n = n;
    p->npass = 0;
    int m = n;
    // powers of two: as many radix-8 as possible, the remainder as 4 or 2
    while (m % 8 == 0 && p->npass < 16) { p->radix[p->npass++] = 8; m /= 8; }
    while (m % 4 == 0 && p->npass < 16) { p->radix[p->npass++] = 4; m /= 4; }
    while (m % 2 == 0 && p->npass < 16) { p->radix[p->npass++] = 2; m /= 2; }
    const int odd[] = {3, 5, 7, 11, 13};
    for (int q : odd)
        while (m % q == 0 && p->npass < 16) { p->radix[p->npass++] = q; m /= q; }
    return m == 1;
}

template <int R>
__device__ __forceinline__ void dft_small(float2 (&v)[R], const float2* tw, int tw_n) {
    if constexpr (R == 2) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), t3 = cmul_mi(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    } else if constexpr (R == 8) {
        // two radix-4 on even / odd inputs, then the radix-2 combine with w8^k
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        dft_small<4>(e, tw, tw_n);
        dft_small<4>(o, tw, tw_n);
        const float h = 0.70710678118654752440f;
        o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * w8^1 = (1 - i)/sqrt2
        o[2] = cmul_mi(o[2]);                                                    // * w8^2 = -i
        o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * w8^3 = (-1 - i)/sqrt2
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = cadd(e[k], o[k]);
            v[k + 4] = csub(e[k], o[k]);
        }
    } else {
        // odd prime radix: direct O(R^2) DFT out of the length-n twiddle table (w_R^q = tw[q * n / R])
        float2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = v[r];
        const int step = tw_n / R;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            float2 acc = x[0];
#pragma unroll
            for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(x[r], tw[((r * q) % R) * step]));
            v[q] = acc;
        }
    }
}

// One Stockham pass over `nrows` rows held in shared memory.  Ns = product of the radices of the
// passes already done.  in/out are distinct buffers with row strides ldi/ldo.
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, int ldi, float2* __restrict__ out, int ldo,
                                              int nrows, int n, int Ns, const float2* __restrict__ tw) {
    const int per_row = n / R;
    const int tstep = n / (Ns * R);
    const int total = nrows * per_row;
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
        const int b = w / per_row;
        const int j = w - b * per_row;
        const int k = j % Ns;
        float2 v[R];
        const float2* src = in + (size_t)b * ldi + j;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float2 x = src[r * per_row];
            if (r > 0 && Ns > 1) x = cmul(x, tw[k * r * tstep]);
            v[r] = x;
        }
        dft_small<R>(v, tw, n);
        float2* dst = out + (size_t)b * ldo + (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * Ns] = v[r];
    }
}

// Runs every pass of `plan`; returns the buffer that holds the result.
__device__ float2* run_plan(const FftPlan& plan, float2* a, float2* b, int ld, int nrows, const float2* tw) {
    int Ns = 1;
    float2* in = a;
    float2* out = b;
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p];
        switch (R) {
            case 8: stockham_pass<8>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 4: stockham_pass<4>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 2: stockham_pass<2>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 3: stockham_pass<3>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 5: stockham_pass<5>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 7: stockham_pass<7>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            case 11: stockham_pass<11>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
            default: stockham_pass<13>(in, ld, out, ld, nrows, plan.n, Ns, tw); break;
        }
        Ns *= R;
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
    return in;
}

// ---------------------------------------------------------------------------------------------
// range stage: one CTA = CB consecutive chirps of one (frame, antenna)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
range_fft_kernel(const float2* __restrict__ cube, const float2* __restrict__ table, const float2* __restrict__ tw_g,
                 float2* __restrict__ mid, FftPlan plan, int A, int C_total, int chirp0, int C_used, int S, int CB,
                 int dc_removal) {
    extern __shared__ float2 smem[];
    const int ld = S + 1;
    float2* tw = smem;                 // [S]
    float2* bufA = tw + S;             // [CB][ld]
    float2* bufB = bufA + CB * ld;     // [CB][ld]

    const int blocks_per_fa = C_used / CB;
    const int fa = blockIdx.x / blocks_per_fa;                 // f * A + a
    const int c0 = (blockIdx.x - fa * blocks_per_fa) * CB;     // first chirp (within the subset)
    const int f = fa / A, a = fa - f * A;

    for (int i = threadIdx.x; i < S; i += blockDim.x) tw[i] = tw_g[i];
    const float2* src = cube + ((size_t)fa * C_total + chirp0 + c0) * S;
    for (int i = threadIdx.x; i < CB * S; i += blockDim.x) {
        const int row = i / S, s = i - row * S;
        bufA[row * ld + s] = cmul(src[i], __ldg(table + s));
    }
    __syncthreads();
    float2* res = run_plan(plan, bufA, bufB, ld, CB, tw);

    // transposed store with the range fftshift (np.fft.fftshift moves bin k to (k + S/2) mod S)
    const int half = S / 2;
    float2* dst = mid + (size_t)f * S * A * C_used + (size_t)a * C_used + c0;
    for (int i = threadIdx.x; i < CB * S; i += blockDim.x) {
        const int k = i / CB, row = i - k * CB;
        float2 v = res[row * ld + k];
        if (dc_removal && k == 0) v = make_float2(0.f, 0.f);
        int p = k + half;
        if (p >= S) p -= S;
        dst[(size_t)p * A * C_used + row] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Doppler stage: rows of mid are (f, s, a) flattened, each C contiguous; one CTA = NB rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
doppler_fft_kernel(const float2* __restrict__ mid, const float2* __restrict__ tw_g, float2* __restrict__ rds,
                   FftPlan plan, int C, int NB, long long nrows_total) {
    extern __shared__ float2 smem[];
    const int ld = C + 1;
    float2* tw = smem;
    float2* bufA = tw + C;
    float2* bufB = bufA + NB * ld;

    const long long row0 = (long long)blockIdx.x * NB;
    const int nb = (int)min((long long)NB, nrows_total - row0);
    for (int i = threadIdx.x; i < C; i += blockDim.x) tw[i] = tw_g[i];
    const float2* src = mid + row0 * C;
    for (int i = threadIdx.x; i < nb * C; i += blockDim.x) {
        const int b = i / C, c = i - b * C;
        bufA[b * ld + c] = src[i];
    }
    __syncthreads();
    float2* res = run_plan(plan, bufA, bufB, ld, nb, tw);

    const int half = C / 2;
    for (int i = threadIdx.x; i < nb * C; i += blockDim.x) {
        const int b = i / C, k = i - b * C;              // lanes along the Doppler axis: coalesced rows
        int p = k + half;
        if (p >= C) p -= C;
        rds[(row0 + b) * C + p] = res[b * ld + k];
    }
}

// ---------------------------------------------------------------------------------------------
// layout converters for the legacy adapters
// ---------------------------------------------------------------------------------------------
// rds [F][S][A][C] <-> reference layout [F][A][S][C]: a permutation of whole Doppler rows (C contiguous elements)
__global__ void permute_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, int n1, int n2, int C,
                                    long long rows_total) {
    // in rows are indexed (f, i1, i2) with i1 < n1, i2 < n2; out rows (f, i2, i1)
    const long long per_f = (long long)n1 * n2;
    for (long long row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const long long f = row / per_f;
        const long long rem = row - f * per_f;
        const int i1 = (int)(rem / n2), i2 = (int)(rem - (long long)i1 * n2);
        const float2* src = in + row * C;
        float2* dst = out + ((f * n2 + i2) * n1 + i1) * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = src[c];
    }
}

// ---- power-of-two fast paths (rs_fft_pow2.cuh) ---------------------------------------------------
template <int R1, int R2, int CB>
static int launch_range_pow2(const float2* cube, const float2* table, const float2* tw, float2* mid, int F, int A,
                             int C_total, int chirp0, int C_used, int dc, cudaStream_t st) {
    using G = pow2::Geo<R1, R2>;
    const size_t smem = (size_t)(2 * G::N + CB * G::ROWP_RANGE) * sizeof(float2);
    auto kern = pow2::range_fft_pow2_kernel<R1, R2, CB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long nblocks = (long long)F * A * (C_used / CB);
    if (nblocks >= (1ll << 31)) return 1;
    const int per_sm = (int)((size_t)rs_smem_optin_limit() / smem);
    const long long grid = std::min<long long>(nblocks, (long long)rs_sm_count() * std::max(1, std::min(per_sm, 6)));
    kern<<<(unsigned)grid, pow2::THREADS, smem, st>>>(cube, table, tw, mid, A, C_total, chirp0, C_used, dc, (int)nblocks);
    return 0;
}

template <int R1, int R2, int NB>
static int launch_doppler_pow2(const float2* mid, const float2* tw, float2* rds, long long nrows, cudaStream_t st) {
    using G = pow2::Geo<R1, R2>;
    const size_t smem = (size_t)(G::N + NB * G::ROWP_DOPP) * sizeof(float2);
    auto kern = pow2::doppler_fft_pow2_kernel<R1, R2, NB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long nblocks = (nrows + NB - 1) / NB;
    if (nblocks >= (1ll << 31)) return 1;
    const int per_sm = (int)((size_t)rs_smem_optin_limit() / smem);
    const long long grid = std::min<long long>(nblocks, (long long)rs_sm_count() * std::max(1, std::min(per_sm, 6)));
    kern<<<(unsigned)grid, pow2::THREADS, smem, st>>>(mid, tw, rds, nrows, (int)nblocks);
    return 0;
}

static int largest_divisor_le(int n, int cap) {
    int best = 1;
    for (int d = 1; d <= cap && d <= n; ++d)
        if (n % d == 0) best = d;
    return best;
}

}  // namespace

extern "C" int rs_range_fft(const void* cube, const void* table, const void* twiddle_s, void* mid, int F, int A,
                            int C_total, int chirp0, int C_used, int S, int dc_removal, void* stream) {
    RS_CHECK_ARG(cube && table && twiddle_s && mid, "rs_range_fft: null pointer");
    RS_CHECK_ARG(F > 0 && A > 0 && A <= RS_MAX_ANTENNAS && S > 0 && S <= RS_MAX_RANGE_BINS, "rs_range_fft: bad F/A/S");
    RS_CHECK_ARG(C_used > 0 && chirp0 >= 0 && chirp0 + C_used <= C_total && C_used <= RS_MAX_DOPPLER_BINS,
                 "rs_range_fft: bad chirp subset");
    if (C_used % 32 == 0 && (S == 64 || S == 128 || S == 256 || S == 512)) {
        const float2 *cu = (const float2*)cube, *tb = (const float2*)table, *tw = (const float2*)twiddle_s;
        cudaStream_t st = (cudaStream_t)stream;
        int rc = 1;
        const char* cb_env = getenv("RS_FFT_CB");
        // measured on B200 (1k frames 256x128x8): CB=16 0.94 ms, CB=32 1.02 ms -- occupancy beats longer store runs
        if (S == 256 && !(cb_env && atoi(cb_env) == 32)) rc = launch_range_pow2<16, 16, 16>(cu, tb, tw, (float2*)mid, F, A, C_total, chirp0, C_used, dc_removal, st);
        else if (S == 256) rc = launch_range_pow2<16, 16, 32>(cu, tb, tw, (float2*)mid, F, A, C_total, chirp0, C_used, dc_removal, st);
        else if (S == 512) rc = launch_range_pow2<32, 16, 16>(cu, tb, tw, (float2*)mid, F, A, C_total, chirp0, C_used, dc_removal, st);
        else if (S == 128) rc = launch_range_pow2<16, 8, 32>(cu, tb, tw, (float2*)mid, F, A, C_total, chirp0, C_used, dc_removal, st);
        else rc = launch_range_pow2<8, 8, 32>(cu, tb, tw, (float2*)mid, F, A, C_total, chirp0, C_used, dc_removal, st);
        if (rc == 0) {
            RS_CHECK_LAUNCH("rs_range_fft(pow2)");
            return RS_OK;
        }
    }
    FftPlan plan;
    RS_CHECK_ARG(make_plan(S, &plan), "rs_range_fft: S=%d has a prime factor > 13", S);
    const int limit = rs_smem_optin_limit();
    int CB = largest_divisor_le(C_used, 16);
    auto need = [&](int cb) { return (size_t)(S + 2 * cb * (S + 1)) * sizeof(float2); };
    while (CB > 1 && need(CB) > (size_t)limit / 2) CB = largest_divisor_le(C_used, CB - 1);
    if (need(CB) > (size_t)limit) {
        rs_set_error("rs_range_fft: S=%d needs %zu B of shared memory", S, need(CB));
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(range_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(CB));
    const long long blocks = (long long)F * A * (C_used / CB);
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_range_fft: too many blocks");
    range_fft_kernel<<<(unsigned)blocks, 256, need(CB), (cudaStream_t)stream>>>(
        (const float2*)cube, (const float2*)table, (const float2*)twiddle_s, (float2*)mid, plan, A, C_total, chirp0,
        C_used, S, CB, dc_removal);
    RS_CHECK_LAUNCH("rs_range_fft");
    return RS_OK;
}

extern "C" int rs_doppler_fft(const void* mid, const void* twiddle_c, void* rds, int F, int A, int C, int S,
                              void* stream) {
    RS_CHECK_ARG(mid && twiddle_c && rds, "rs_doppler_fft: null pointer");
    RS_CHECK_ARG(F > 0 && A > 0 && A <= RS_MAX_ANTENNAS && S > 0 && C > 0 && C <= RS_MAX_DOPPLER_BINS,
                 "rs_doppler_fft: bad dims");
    if (C == 64 || C == 128 || C == 256) {
        const long long nrows = (long long)F * S * A;
        cudaStream_t st = (cudaStream_t)stream;
        const float2 *m = (const float2*)mid, *tw = (const float2*)twiddle_c;
        int rc = 1;
        if (C == 256) rc = launch_doppler_pow2<16, 16, 32>(m, tw, (float2*)rds, nrows, st);
        else if (C == 128 && !(getenv("RS_FFT_NB") && atoi(getenv("RS_FFT_NB")) == 64)) rc = launch_doppler_pow2<16, 8, 32>(m, tw, (float2*)rds, nrows, st);
        else if (C == 128) rc = launch_doppler_pow2<16, 8, 64>(m, tw, (float2*)rds, nrows, st);
        else rc = launch_doppler_pow2<8, 8, 64>(m, tw, (float2*)rds, nrows, st);
        if (rc == 0) {
            RS_CHECK_LAUNCH("rs_doppler_fft(pow2)");
            return RS_OK;
        }
    }
    FftPlan plan;
    RS_CHECK_ARG(make_plan(C, &plan), "rs_doppler_fft: C=%d has a prime factor > 13", C);
    const int limit = rs_smem_optin_limit();
    int NB = (A <= 16) ? A * (16 / A) : largest_divisor_le(A, 16);
    auto need = [&](int nb) { return (size_t)(C + 2 * nb * (C + 1)) * sizeof(float2); };
    while (NB > 1 && need(NB) > (size_t)limit / 2) NB = (NB % 2 == 0) ? NB / 2 : NB - 1;
    if (need(NB) > (size_t)limit) {
        rs_set_error("rs_doppler_fft: C=%d needs %zu B of shared memory", C, need(NB));
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(doppler_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(NB));
    const long long nrows = (long long)F * S * A;
    const long long blocks = (nrows + NB - 1) / NB;
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_doppler_fft: too many blocks");
    doppler_fft_kernel<<<(unsigned)blocks, 256, need(NB), (cudaStream_t)stream>>>(
        (const float2*)mid, (const float2*)twiddle_c, (float2*)rds, plan, C, NB, nrows);
    RS_CHECK_LAUNCH("rs_doppler_fft");
    return RS_OK;
}

extern "C" int rs_rds_to_reference_layout(const void* rds, void* out, int F, int A, int C, int S, void* stream) {
    RS_CHECK_ARG(rds && out && F > 0 && A > 0 && C > 0 && S > 0, "rs_rds_to_reference_layout: bad args");
    const long long rows = (long long)F * S * A;
    const unsigned grid = (unsigned)std::min<long long>(rows, (long long)rs_sm_count() * 16);
    permute_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float2*)rds, (float2*)out, S, A, C, rows);
    RS_CHECK_LAUNCH("rs_rds_to_reference_layout");
    return RS_OK;
}

extern "C" int rs_rds_from_reference_layout(const void* rds_ref, void* out, int F, int A, int C, int S, void* stream) {
    RS_CHECK_ARG(rds_ref && out && F > 0 && A > 0 && C > 0 && S > 0, "rs_rds_from_reference_layout: bad args");
    const long long rows = (long long)F * S * A;
    const unsigned grid = (unsigned)std::min<long long>(rows, (long long)rs_sm_count() * 16);
    permute_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float2*)rds_ref, (float2*)out, A, S, C, rows);
    RS_CHECK_LAUNCH("rs_rds_from_reference_layout");
    return RS_OK;
}
