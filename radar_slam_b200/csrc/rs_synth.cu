// Device-side frame synthesis: FMCWRadarSimulator.synthesize_frame
// (scripts/simulate_raw.py:147-221) for a batch of frames.
//
// The reference builds, per scatterer, antenna and chirp, delayed_chirp(t - tau) * conj(ref_chirp(t)) scaled by
// amplitude * exp(i (doppler_phase + antenna_phase)) (simulate_raw.py:102-145, 190-209) and adds complex Gaussian noise
// (:216-219).  Two facts shape the kernels:
//   * the chirp index never enters the scatterer term (chirp_start_time is computed and not used, :192), so the term
//     is one [A][S] plane per frame: kernel 1 evaluates it once in fp64.  The product of the two chirps is evaluated in
//     closed form, exp(2 pi i (-fc tau - k tau t + k tau^2 / 2)): the reference subtracts two phases of ~1e7 rad in
//     fp64, so the two agree to ~2e-9 -- far inside the complex64 input rounding of the pipeline.
//   * what remains is a pure streaming write of plane + noise: kernel 2 writes one float4 (two cells) per thread
//     and step, noise from a counter-based Philox4x32-10 stream keyed by (seed, absolute frame index), so frame k
//     is the same whatever batch or rank generates it.  numpy's Mersenne-Twister stream (np.random.randn, :216-218)
//     cannot be reproduced on a GPU; the noise matches in distribution, not sample by sample.
#include "rs_common.cuh"

namespace {

constexpr int SYN_THREADS = 256;

__global__ void __launch_bounds__(SYN_THREADS)
scatterer_plane_kernel(const double* __restrict__ scat, const int32_t* __restrict__ n_scat, int n_max, double fc,
                       double slope, double T, double lambda_c, const double* __restrict__ ant_pos, int A, int S,
                       float2* __restrict__ plane) {
    // one CTA per (frame, antenna); thread = fast-time sample
    const int f = blockIdx.x / A, a = blockIdx.x - f * A;
    const int n = n_scat ? min(n_scat[f], n_max) : n_max;
    const double c0 = 3e8;
    const double pos = ant_pos[a];
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        // np.linspace(0, T, S): arange(S) * (T / (S - 1)), last element forced to T
        const double t = S > 1 ? (s == S - 1 ? T : (double)s * (T / (double)(S - 1))) : 0.0;
        double re = 0, im = 0;
        for (int k = 0; k < n; ++k) {
            const double* q = scat + ((size_t)f * n_max + k) * 4;
            const double rng = q[0], az = q[1], rcs = q[2], vr = q[3];
            if (!(rng > 0) || !isfinite(rng) || !isfinite(az) || !isfinite(rcs) || !isfinite(vr)) continue;   // :177
            const double tau = 2.0 * rng / c0;                                       // :121
            const double td = t - tau;
            if (td < 0 || td > T) continue;                                          // :198
            const double amp = sqrt(pow(10.0, rcs / 10.0)) / (4.0 * 3.14159265358979323846 * rng * rng);   // :124-125
            // phases in cycles: doppler (:128) + antenna (:140) + beat of the two chirps
            const double cyc_ant = 2.0 * vr * fc / c0 + pos * sin(az) / lambda_c;
            const double cyc_beat = -fc * tau - slope * tau * t + 0.5 * slope * tau * tau;
            double cyc = cyc_ant + cyc_beat;
            cyc -= rint(cyc);
            double sn, cs;
            sincospi(2.0 * cyc, &sn, &cs);
            re += amp * cs;
            im += amp * sn;
        }
        plane[((size_t)f * A + a) * S + s] = make_float2((float)re, (float)im);
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// two standard-normal pairs from four 32-bit words (Box-Muller on 24-bit uniforms in (0, 1))
__device__ __forceinline__ float4 normal4(const uint32_t (&c)[4]) {
    const float k = 5.9604644775390625e-8f;      // 2^-24
    const float u0 = ((float)(c[0] >> 8) + 0.5f) * k, u1 = ((float)(c[1] >> 8) + 0.5f) * k;
    const float u2 = ((float)(c[2] >> 8) + 0.5f) * k, u3 = ((float)(c[3] >> 8) + 0.5f) * k;
    const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__global__ void __launch_bounds__(SYN_THREADS)
cube_stream_kernel(const float2* __restrict__ plane, float sigma, uint32_t seed_lo, uint32_t seed_hi,
                   long long first_frame, float4* __restrict__ cube, int A, int C, int S, long long pairs_total) {
    // thread = one float4 = two consecutive fast-time cells; S is even (checked by the entry point)
    const long long pairs_per_frame = (long long)A * C * (S / 2);
    const int half = S / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs_total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / pairs_per_frame;
        const long long in = i - f * pairs_per_frame;
        const int sp = (int)(in % half);
        const long long ac = in / half;
        const int a = (int)(ac / C);
        const float4 p = __ldg(reinterpret_cast<const float4*>(plane + ((size_t)f * A + a) * S) + sp);
        float4 out = p;
        if (sigma > 0.f) {
            const unsigned long long fa = (unsigned long long)(first_frame + f);
            uint32_t ctr[4] = {(uint32_t)in, (uint32_t)((unsigned long long)in >> 32), (uint32_t)fa, (uint32_t)(fa >> 32)};
            philox4x32_10(ctr, seed_lo, seed_hi);
            const float4 z = normal4(ctr);
            out.x = fmaf(sigma, z.x, p.x); out.y = fmaf(sigma, z.y, p.y);
            out.z = fmaf(sigma, z.z, p.z); out.w = fmaf(sigma, z.w, p.w);
        }
        __stcs(cube + i, out);
    }
}

// odd S: one cell per thread.  The noise of cell n of a frame is component (n & 1) of the Philox block of pair n >> 1,
// i.e. the same stream as the vector kernel draws for even S.
__global__ void __launch_bounds__(SYN_THREADS)
cube_stream_scalar_kernel(const float2* __restrict__ plane, float sigma, uint32_t seed_lo, uint32_t seed_hi,
                          long long first_frame, float2* __restrict__ cube, int A, int C, int S, long long cells_total) {
    const long long cells_per_frame = (long long)A * C * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells_total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / cells_per_frame;
        const long long in = i - f * cells_per_frame;
        const int s_ = (int)(in % S);
        const int a = (int)((in / S) / C);
        float2 out = __ldg(plane + ((size_t)f * A + a) * S + s_);
        if (sigma > 0.f) {
            const unsigned long long fa = (unsigned long long)(first_frame + f), pr = (unsigned long long)in >> 1;
            uint32_t ctr[4] = {(uint32_t)pr, (uint32_t)(pr >> 32), (uint32_t)fa, (uint32_t)(fa >> 32)};
            philox4x32_10(ctr, seed_lo, seed_hi);
            const float4 z = normal4(ctr);
            out.x = fmaf(sigma, (in & 1) ? z.z : z.x, out.x);
            out.y = fmaf(sigma, (in & 1) ? z.w : z.y, out.y);
        }
        __stcs(cube + i, out);
    }
}

}  // namespace

extern "C" int rs_synthesize_frames(const double* scatterers, const int32_t* n_scatterers, int n_max, double fc,
                                    double chirp_rate, double chirp_duration, double lambda_c,
                                    const double* antenna_pos, double noise_power, unsigned long long seed,
                                    long long first_frame, void* plane_ws, void* cube, int F, int A, int C, int S,
                                    void* stream) {
    RS_CHECK_ARG(scatterers && antenna_pos && plane_ws && cube, "rs_synthesize_frames: null pointer");
    RS_CHECK_ARG(F > 0 && A > 0 && C > 0 && S > 0 && n_max >= 0 && noise_power >= 0 && first_frame >= 0,
                 "rs_synthesize_frames: bad dims");
    RS_CHECK_ARG((long long)F * A < (1ll << 31), "rs_synthesize_frames: too many planes");
    cudaStream_t st = (cudaStream_t)stream;
    scatterer_plane_kernel<<<(unsigned)(F * A), SYN_THREADS, 0, st>>>(scatterers, n_scatterers, n_max, fc, chirp_rate,
                                                                     chirp_duration, lambda_c, antenna_pos, A, S,
                                                                     (float2*)plane_ws);
    RS_CHECK_LAUNCH("rs_synthesize_frames(plane)");
    if (S % 2 != 0) {
        const long long cells = (long long)F * A * C * S;
        const long long want1 = (cells + SYN_THREADS - 1) / SYN_THREADS;
        const long long cap1 = (long long)rs_sm_count() * 32;
        cube_stream_scalar_kernel<<<(unsigned)(want1 < cap1 ? want1 : cap1), SYN_THREADS, 0, st>>>(
            (const float2*)plane_ws, (float)sqrt(noise_power), (uint32_t)seed, (uint32_t)(seed >> 32), first_frame,
            (float2*)cube, A, C, S, cells);
        RS_CHECK_LAUNCH("rs_synthesize_frames(stream, odd S)");
        return RS_OK;
    }
    const long long pairs = (long long)F * A * C * (S / 2);
    const long long want = (pairs + SYN_THREADS - 1) / SYN_THREADS;
    const long long cap = (long long)rs_sm_count() * 32;
    cube_stream_kernel<<<(unsigned)(want < cap ? want : cap), SYN_THREADS, 0, st>>>(
        (const float2*)plane_ws, (float)sqrt(noise_power), (uint32_t)seed, (uint32_t)(seed >> 32), first_frame,
        (float4*)cube, A, C, S, pairs);
    RS_CHECK_LAUNCH("rs_synthesize_frames(stream)");
    return RS_OK;
}
