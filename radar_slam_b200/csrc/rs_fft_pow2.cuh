// Two-pass register FFT kernels for power-of-two lengths N = R1 * R2 (R1 >= R2, R1 in {4, 8, 16, 32}, R2 in {4, 8, 16}).
//
// Pass 1: a thread owns x[t + R2 j], j < R1 (lanes run along the FFT axis -> coalesced loads),
//         does the R1-point DFT in registers, applies the inter-pass twiddle w_N^{t k1} and writes
//         Y[row][k1][t] to shared memory (padded, conflict-free).
// Pass 2: a thread owns one (row, k1): reads Y[row][k1][0..R2), does the R2-point DFT and stores
//         X[k1 + R1 k2] straight to global memory.  The lane -> (row, k1) map of pass 2 is chosen per
//         kernel so that the TRANSPOSED store of each stage is coalesced:
//           range stage   lanes = 32 consecutive chirps            -> mid[f][s][a][c..c+31]   (256 B runs)
//           Doppler stage lanes = 16 consecutive k1 of a row       -> rds[f][s][a][k..k+15]   (128 B runs)
// One shared-memory round trip per element, no staging copies.
#pragma once
#include "rs_common.cuh"

namespace pow2 {

template <int R>
__device__ __forceinline__ void dft(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dft<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft<4>(float2 (&v)[4]) {
    const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    const float2 t2 = cadd(v[1], v[3]), t3 = cmul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void dft<8>(float2 (&v)[8]) {
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dft<4>(e);
    dft<4>(o);
    const float h = 0.70710678118654752440f;
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    o[2] = cmul_mi(o[2]);
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = cadd(e[k], o[k]);
        v[k + 4] = csub(e[k], o[k]);
    }
}
// 16 = 4 x 4: U[a][q] = sum_b v[a + 4b] w4^{bq};  X[q + 4r] = sum_a w4^{ar} (w16^{aq} U[a][q])
template <>
__device__ __forceinline__ void dft<16>(float2 (&v)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin (pi/8)
    const float h = 0.70710678118654752440f;
    float2 u[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float2 w[4] = {v[a], v[a + 4], v[a + 8], v[a + 12]};
        dft<4>(w);
#pragma unroll
        for (int q = 0; q < 4; ++q) u[a][q] = w[q];
    }
    // w16^m = exp(-2 pi i m / 16)
    auto mul_w = [&](float2 x, int m) -> float2 {
        switch (m) {
            case 0: return x;
            case 1: return make_float2(c1 * x.x + s1 * x.y, c1 * x.y - s1 * x.x);
            case 2: return make_float2(h * (x.x + x.y), h * (x.y - x.x));
            case 3: return make_float2(s1 * x.x + c1 * x.y, s1 * x.y - c1 * x.x);
            case 4: return cmul_mi(x);
            case 6: return make_float2(h * (x.y - x.x), -h * (x.x + x.y));
            default: /* 9 */ return make_float2(-c1 * x.x - s1 * x.y, s1 * x.x - c1 * x.y);
        }
    };
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 w[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) w[a] = mul_w(u[a][q], a * q);
        dft<4>(w);
#pragma unroll
        for (int r = 0; r < 4; ++r) v[q + 4 * r] = w[r];
    }
}

// 32 = 2 x 16: X[k] = E[k] + w32^k O[k], X[k + 16] = E[k] - w32^k O[k]
template <>
__device__ __forceinline__ void dft<32>(float2 (&v)[32]) {
    constexpr float WC[16] = {1.00000000000000000000f, 0.98078528040323043058f, 0.92387953251128673848f, 0.83146961230254523567f, 0.70710678118654757274f, 0.55557023301960228867f, 0.38268343236508983729f, 0.19509032201612833135f, 0.00000000000000006123f, -0.19509032201612819257f, -0.38268343236508972627f, -0.55557023301960195560f, -0.70710678118654746172f, -0.83146961230254534669f, -0.92387953251128673848f, -0.98078528040323043058f};
    constexpr float WS[16] = {-0.00000000000000000000f, -0.19509032201612824808f, -0.38268343236508978178f, -0.55557023301960217765f, -0.70710678118654746172f, -0.83146961230254523567f, -0.92387953251128673848f, -0.98078528040323043058f, -1.00000000000000000000f, -0.98078528040323043058f, -0.92387953251128673848f, -0.83146961230254545772f, -0.70710678118654757274f, -0.55557023301960217765f, -0.38268343236508989280f, -0.19509032201612860891f};
    float2 e[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    dft<16>(e);
    dft<16>(o);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 t = (k == 0) ? o[0] : make_float2(fmaf(o[k].x, WC[k], -o[k].y * WS[k]), fmaf(o[k].x, WS[k], o[k].y * WC[k]));
        v[k] = cadd(e[k], t);
        v[k + 16] = csub(e[k], t);
    }
}

// ---- packed variants: Blackwell f32x2 arithmetic (FADD2 / FMUL2 / FFMA2) ----------------------------------------------
// A complex value is an (re, im) register pair, so a complex add is ONE add.rn.f32x2 with the IEEE result of two scalar
// adds.  ptxas folds the half swap and the per-half sign of the packed operands into the instruction's operand
// modifiers (SASS: R.F32x2.LO_HI, .NP / .PN, and a scalar broadcast R.F32), so
//   a -+ i b         = add.f32x2(a, (b.y, -b.x)) / add.f32x2(a, (-b.y, b.x))               1 instruction
//   a * b (complex)  = fma.f32x2((a.y, a.x), (-b.y, b.y), mul.f32x2(a, (b.x, b.x)))        2 instructions, b stays (re, im)
// dft<16>: 154 -> 80 floating-point instructions, dft<8>: 64 -> 28, a twiddle multiply: 4 -> 2.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 up2(unsigned long long a) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}
__device__ __forceinline__ float2 padd(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return up2(r);
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return up2(r);
}
// a + (-i) b and a - (-i) b
__device__ __forceinline__ float2 add_mi(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.y, -b.x)));
    return up2(r);
}
__device__ __forceinline__ float2 sub_mi(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(-b.y, b.x)));
    return up2(r);
}
// complex product
__device__ __forceinline__ float2 cmulp(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.x)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r) : "l"(pk2(a.y, a.x)), "l"(pk2(-b.y, b.y)));
    return up2(r);
}

template <int R>
__device__ __forceinline__ void dftp(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dftp<4>(float2 (&v)[4]) {
    const float2 t0 = padd(v[0], v[2]), t1 = psub(v[0], v[2]);
    const float2 t2 = padd(v[1], v[3]), d = psub(v[1], v[3]);
    v[0] = padd(t0, t2);
    v[2] = psub(t0, t2);
    v[1] = add_mi(t1, d);
    v[3] = sub_mi(t1, d);
}
template <>
__device__ __forceinline__ void dftp<8>(float2 (&v)[8]) {
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dftp<4>(e);
    dftp<4>(o);
    const float h = 0.70710678118654752440f;
    o[1] = cmulp(o[1], make_float2(h, -h));             // w8^1
    o[3] = cmulp(o[3], make_float2(-h, -h));            // w8^3
    v[0] = padd(e[0], o[0]);
    v[4] = psub(e[0], o[0]);
    v[1] = padd(e[1], o[1]);
    v[5] = psub(e[1], o[1]);
    v[2] = add_mi(e[2], o[2]);                          // w8^2 = -i
    v[6] = sub_mi(e[2], o[2]);
    v[3] = padd(e[3], o[3]);
    v[7] = psub(e[3], o[3]);
}
template <>
__device__ __forceinline__ void dftp<16>(float2 (&v)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin (pi/8)
    const float h = 0.70710678118654752440f;
    float2 u[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float2 w[4] = {v[a], v[a + 4], v[a + 8], v[a + 12]};
        dftp<4>(w);
#pragma unroll
        for (int q = 0; q < 4; ++q) u[a][q] = w[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 w[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int m = a * q;                        // w16^m = exp(-2 pi i m / 16)
            const float2 x = u[a][q];
            w[a] = m == 0 ? x
                 : m == 1 ? cmulp(x, make_float2(c1, -s1))
                 : m == 2 ? cmulp(x, make_float2(h, -h))
                 : m == 3 ? cmulp(x, make_float2(s1, -c1))
                 : m == 4 ? make_float2(x.y, -x.x)
                 : m == 6 ? cmulp(x, make_float2(-h, -h))
                 : /* 9 */  cmulp(x, make_float2(-c1, s1));
        }
        dftp<4>(w);
#pragma unroll
        for (int r = 0; r < 4; ++r) v[q + 4 * r] = w[r];
    }
}

constexpr int THREADS = 256;

template <int R1, int R2>
struct Geo {
    static constexpr int N = R1 * R2;
    static constexpr int K1P = R2 + 1;            // pitch of one k1 line (complex); odd -> conflict-free pass-2 reads
    static constexpr int ROWP_RANGE = R1 * K1P + ((R1 * K1P) % 2 == 0 ? 1 : 0);    // odd: lanes along rows
    static constexpr int ROWP_DOPP = ROWP_RANGE;                                    // lanes (row pair, k1)
};

// ---- pass 1 for `nrows` rows whose element (row, n) is at src[row * row_stride + n] -------------
template <int R1, int R2, bool TABLE>
__device__ __forceinline__ void pass1(const float2* __restrict__ src, size_t row_stride, int nrows, int rowp,
                                      const float2* __restrict__ tabs, const float2* __restrict__ tw1,
                                      float2* __restrict__ Y) {
    using G = Geo<R1, R2>;
    for (int it = threadIdx.x; it < nrows * R2; it += THREADS) {
        const int row = it / R2, t = it - row * R2;
        const float2* x = src + (size_t)row * row_stride + t;
        float2 v[R1];
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = __ldcs(x + R2 * j);          // streamed once: evict-first
        if (TABLE) {
#pragma unroll
            for (int j = 0; j < R1; ++j) v[j] = cmul(v[j], tabs[t + R2 * j]);
        }
        dft<R1>(v);
        float2* y = Y + row * rowp + t;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) y[k1 * G::K1P] = (k1 == 0) ? v[0] : cmul(v[k1], tw1[k1 * R2 + t]);
    }
}

// ---------------------------------------------------------------------------------------------
// range stage: CTA = CB consecutive chirps of one (frame, antenna)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int CB>
__global__ void __launch_bounds__(THREADS)
range_fft_pow2_kernel(const float2* __restrict__ cube, const float2* __restrict__ table, const float2* __restrict__ tw_g,
                      float2* __restrict__ mid, int A, int C_total, int chirp0, int C_used, int dc_removal,
                      int nblocks) {
    using G = Geo<R1, R2>;
    constexpr int S = G::N, RP = G::ROWP_RANGE;
    extern __shared__ float2 sm[];
    float2* tabs = sm;               // [S]
    float2* tw1 = tabs + S;          // [R1][R2]
    float2* Y = tw1 + S;             // [CB][RP]
    for (int i = threadIdx.x; i < S; i += THREADS) {
        tabs[i] = table[i];
        const int k1 = i / R2, t = i - k1 * R2;
        tw1[i] = tw_g[(k1 * t) % S];
    }
    __syncthreads();
    const int per_fa = C_used / CB;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int fa = blk / per_fa;
        const int c0 = (blk - fa * per_fa) * CB;
        const int f = fa / A, a = fa - f * A;
        pass1<R1, R2, true>(cube + ((size_t)fa * C_total + chirp0 + c0) * S, S, CB, RP, tabs, tw1, Y);
        __syncthreads();
        float2* dst = mid + ((size_t)f * S * A + a) * C_used + c0;
        for (int it = threadIdx.x; it < CB * R1; it += THREADS) {
            const int row = it % CB, k1 = it / CB;          // lanes along chirps
            const float2* y = Y + row * RP + k1 * G::K1P;
            float2 u[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) u[n2] = y[n2];
            dft<R2>(u);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int k = k1 + R1 * k2;
                float2 val = u[k2];
                if (dc_removal && k == 0) val = make_float2(0.f, 0.f);
                const int p = (k + S / 2) & (S - 1);                    // range fftshift
                dst[(size_t)p * A * C_used + row] = val;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Doppler stage: rows of mid are (f, s, a) flattened, C contiguous; the RDS keeps that row order (rds[F][S][A][C]),
// so row g of mid becomes row g of the RDS.  CTA = NB rows.
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int NB>
__global__ void __launch_bounds__(THREADS)
doppler_fft_pow2_kernel(const float2* __restrict__ mid, const float2* __restrict__ tw_g, float2* __restrict__ rds,
                        long long nrows_total, int nblocks) {
    using G = Geo<R1, R2>;
    constexpr int C = G::N, RP = G::ROWP_DOPP;
    extern __shared__ float2 sm[];
    float2* tw1 = sm;                // [R1][R2]
    float2* Y = tw1 + C;             // [NB][RP]
    for (int i = threadIdx.x; i < C; i += THREADS) {
        const int k1 = i / R2, t = i - k1 * R2;
        tw1[i] = tw_g[(k1 * t) % C];
    }
    __syncthreads();
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long long row0 = (long long)blk * NB;
        const int nb = (int)min((long long)NB, nrows_total - row0);
        pass1<R1, R2, false>(mid + row0 * C, C, nb, RP, nullptr, tw1, Y);
        __syncthreads();
        // item = (row, k1): R1 consecutive lanes share a row, so every k2 stores R1 consecutive Doppler bins
        for (int it = threadIdx.x; it < nb * R1; it += THREADS) {
            const int b = it / R1, k1 = it - b * R1;
            const float2* y = Y + b * RP + k1 * G::K1P;
            float2 u[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) u[n2] = y[n2];
            dft<R2>(u);
            float2* dst = rds + (row0 + b) * C;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) dst[(k1 + R1 * k2 + C / 2) & (C - 1)] = u[k2];      // Doppler fftshift
        }
        __syncthreads();
    }
}

}  // namespace pow2
