// The fp64 frame path behind the legacy class API (one frame per call, numpy complex128 in and out).
//
// The reference computes the range-Doppler spectrum, the peak mask and every pseudo-spectrum in float64
// (dechirp.py:193-278, angle_estimation.py:83-176).  The batched throughput path (rs_range_doppler_fft, rs_detect,
// rs_angles) works in fp32 and settles what fp32 cannot decide with an fp64 recheck from the raw cube -- but the legacy
// methods are handed a bare RDS array (possibly np.load-ed from a stage file), so there is no cube to go back to.
// They therefore run these kernels instead: the same operations as the reference, in fp64, on exactly the array the
// caller passed.  Cost is irrelevant here (one frame per call); none of this is on the batched path.
//
//   rs_range_doppler_f64   (x conj(ref)) w - mean  ->  FFT over fast time  ->  FFT over slow time  ->  both fftshifts,
//                          reference layout rds[a][r][d]                               (dechirp.py:139,108,120,193-211)
//   rs_detect_f64          10 log10(|X|^2 + 1e-12), 3x3 'reflect' maximum filter equality on the dB values, strict
//                          threshold, range gate, detections compacted IN REFERENCE ORDER (antenna, range, doppler)
//                          by row counts + one scan + an ordered per-row write                  (dechirp.py:235-263)
//   rs_signatures_c128     unit-energy snapshots of listed cells from a complex128 RDS      (angle_estimation.py:83-88)
#include <algorithm>
#include "rs_common.cuh"

namespace {

struct PlanF64 {
    int n, npass;
    int radix[20];
};

static bool make_plan64(int n, PlanF64* p) {
    p->n = n;
    p->npass = 0;
    int m = n;
    const int rad[] = {4, 2, 3, 5, 7, 11, 13};
    for (int q : rad)
        while (m % q == 0 && p->npass < 20) { p->radix[p->npass++] = q; m /= q; }
    return m == 1;
}

__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// All Stockham passes of one row of length n held in shared memory; returns the buffer with the result.
// Butterflies are direct O(R^2) sums out of the length-n twiddle table (w_R^q = tw[q n / R]): accuracy first.
__device__ double2* fft_row_f64(const PlanF64& plan, double2* a, double2* b, const double2* __restrict__ tw) {
    const int n = plan.n;
    int Ns = 1;
    double2* in = a;
    double2* out = b;
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p];
        const int per = n / R, tstep = n / (Ns * R), wstep = n / R;
        for (int j = threadIdx.x; j < per; j += blockDim.x) {
            const int k = j % Ns;
            double2 x[13];
            for (int r = 0; r < R; ++r) {
                double2 v = in[j + r * per];
                if (r > 0 && Ns > 1) v = zmul(v, tw[(k * r * tstep) % n]);
                x[r] = v;
            }
            double2* dst = out + (j - k) * R + k;
            for (int q = 0; q < R; ++q) {
                double2 acc = x[0];
                for (int r = 1; r < R; ++r) {
                    const double2 t = zmul(x[r], tw[((r * q) % R) * wstep]);
                    acc.x += t.x;
                    acc.y += t.y;
                }
                dst[q * Ns] = acc;
            }
        }
        Ns *= R;
        __syncthreads();
        double2* t = in; in = out; out = t;
    }
    return in;
}

// one CTA per (antenna, chirp): dechirp, window, mean removal, range FFT, range fftshift -> mid[a][p][c]
__global__ void __launch_bounds__(256)
range_f64_kernel(const double2* __restrict__ cube, const double2* __restrict__ ref, const double* __restrict__ win,
                 const double2* __restrict__ tw, double2* __restrict__ mid, PlanF64 plan, int C_total, int chirp0,
                 int C_used, int S, int dc) {
    extern __shared__ double2 sm64[];
    double2* bufA = sm64;
    double2* bufB = sm64 + S;
    __shared__ double red[2][8];
    const int a = blockIdx.x / C_used, c = blockIdx.x - a * C_used;
    const double2* x = cube + ((size_t)a * C_total + chirp0 + c) * S;
    double sr = 0, si = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const double2 v = x[s], r = ref[s];
        double2 b = make_double2(v.x * r.x + v.y * r.y, v.y * r.x - v.x * r.y);     // v * conj(r)
        b.x *= win[s];
        b.y *= win[s];
        bufA[s] = b;
        sr += b.x;
        si += b.y;
    }
    if (dc) {
        for (int off = 16; off; off >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, off);
            si += __shfl_xor_sync(0xffffffffu, si, off);
        }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = si; }
        __syncthreads();
        double mr = 0, mi = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { mr += red[0][w]; mi += red[1][w]; }
        mr /= S;
        mi /= S;
        for (int s = threadIdx.x; s < S; s += blockDim.x) { bufA[s].x -= mr; bufA[s].y -= mi; }
    }
    __syncthreads();
    const double2* res = fft_row_f64(plan, bufA, bufB, tw);
    const int half = S / 2;                                          // np.fft.fftshift: bin k -> (k + S/2) mod S
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
        int p = k + half;
        if (p >= S) p -= S;
        mid[((size_t)a * S + p) * C_used + c] = res[k];
    }
}

// one CTA per (antenna, range bin): Doppler FFT in place in the reference layout, Doppler fftshift
__global__ void __launch_bounds__(128)
doppler_f64_kernel(double2* __restrict__ rds, const double2* __restrict__ tw, PlanF64 plan, int C) {
    extern __shared__ double2 sm64[];
    double2* bufA = sm64;
    double2* bufB = sm64 + C;
    double2* row = rds + (size_t)blockIdx.x * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) bufA[c] = row[c];
    __syncthreads();
    const double2* res = fft_row_f64(plan, bufA, bufB, tw);
    const int half = C / 2;
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        int p = k + half;
        if (p >= C) p -= C;
        row[p] = res[k];
    }
}

// 10 log10(|X|^2 + 1e-12) with |X| = hypot(re, im) as np.abs evaluates it (dechirp.py:235-238)
__global__ void power_db_c128_kernel(const double2* __restrict__ rds, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 x = rds[i];
        const double m = hypot(x.x, x.y);
        out[i] = 10.0 * log10(m * m + 1e-12);
    }
}

// scipy.ndimage.maximum_filter(size=3, mode='reflect') == value  <=>  no in-range neighbour is larger (the reflected
// neighbour of an edge cell is the cell's own row / column); ties count (dechirp.py:250-251)
__device__ __forceinline__ bool is_peak_db(const double* __restrict__ plane, int r, int d, int R, int D, double thr) {
    const double v = plane[(size_t)r * D + d];
    if (!(v > thr)) return false;
    for (int dr = -1; dr <= 1; ++dr) {
        const int rr = r + dr;
        if (rr < 0 || rr >= R) continue;
        for (int dd = -1; dd <= 1; ++dd) {
            const int c = d + dd;
            if (c < 0 || c >= D) continue;
            if (plane[(size_t)rr * D + c] > v) return false;
        }
    }
    return true;
}

// one warp per (antenna, range) row; WRITE = false counts, WRITE = true emits keys at the row's offset in Doppler order
template <bool WRITE>
__global__ void detect_rows_f64_kernel(const double* __restrict__ db, const uint8_t* __restrict__ gate, double thr,
                                       int* __restrict__ row_count, const long long* __restrict__ row_off,
                                       uint32_t* __restrict__ keys, long long cap, int A, int R, int D) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)A * R) return;
    const int a = (int)(row / R), r = (int)(row - (long long)a * R);
    const double* plane = db + (size_t)a * R * D;
    const bool open = gate[r] != 0;
    int n = 0;
    long long base = WRITE ? row_off[row] : 0;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        const bool hit = open && d < D && is_peak_db(plane, r, d, R, D, thr);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (WRITE && hit) {
            const long long pos = base + n + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) keys[pos] = rs_make_key(a, r, d);
        }
        n += __popc(m);
    }
    if (!WRITE && lane == 0) row_count[row] = n;
}

// exclusive scan of the row counts (one CTA; rows <= RS_MAX_ANTENNAS * RS_MAX_RANGE_BINS), total in *total
__global__ void scan_rows_kernel(const int* __restrict__ row_count, long long* __restrict__ row_off, long long nrows,
                                 long long* __restrict__ total) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long i0 = 0; i0 < nrows; i0 += blockDim.x) {
        const long long i = i0 + threadIdx.x;
        const long long v = i < nrows ? row_count[i] : 0;
        long long inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            long long w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
            for (int off = 1; off < 32; off <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += t;
            }
            wsum[lane] = w;                                          // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (wid ? wsum[wid - 1] : 0) + inc - v;
        if (i < nrows) row_off[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// snapshot of cell (r, d) across the antennas, normalised to unit energy when its power is > 0
__global__ void signatures_c128_kernel(const double2* __restrict__ rds, const int* __restrict__ rb,
                                       const int* __restrict__ dbin, int n, double2* __restrict__ out, int A, int R, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t cell = (size_t)rb[i] * D + dbin[i];
    double p = 0;
    for (int a = 0; a < A; ++a) {
        const double2 x = rds[(size_t)a * R * D + cell];
        const double m = hypot(x.x, x.y);                            // np.abs(s) ** 2, angle_estimation.py:86
        p += m * m;
    }
    const double nrm = sqrt(p);
    for (int a = 0; a < A; ++a) {
        const double2 x = rds[(size_t)a * R * D + cell];
        out[(size_t)i * A + a] = p > 0 ? make_double2(x.x / nrm, x.y / nrm) : x;
    }
}

}  // namespace

extern "C" int rs_range_doppler_f64(const void* cube128, const void* ref128, const double* window, const void* twiddle_s128,
                                    const void* twiddle_c128, void* rds128, int A, int C_total, int chirp0, int C_used,
                                    int S, int dc_removal, void* stream) {
    RS_CHECK_ARG(cube128 && ref128 && window && twiddle_s128 && twiddle_c128 && rds128, "rs_range_doppler_f64: null pointer");
    RS_CHECK_ARG(A > 0 && A <= RS_MAX_ANTENNAS && S > 0 && S <= RS_MAX_RANGE_BINS && C_used > 0 &&
                     C_used <= RS_MAX_DOPPLER_BINS && chirp0 >= 0 && chirp0 + C_used <= C_total,
                 "rs_range_doppler_f64: bad dims");
    PlanF64 ps, pc;
    RS_CHECK_ARG(make_plan64(S, &ps), "rs_range_doppler_f64: S=%d has a prime factor > 13", S);
    RS_CHECK_ARG(make_plan64(C_used, &pc), "rs_range_doppler_f64: C=%d has a prime factor > 13", C_used);
    const size_t sm_s = 2 * (size_t)S * sizeof(double2), sm_c = 2 * (size_t)C_used * sizeof(double2);
    cudaFuncSetAttribute(range_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s);
    cudaFuncSetAttribute(doppler_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c);
    cudaStream_t st = (cudaStream_t)stream;
    range_f64_kernel<<<(unsigned)(A * C_used), 256, sm_s, st>>>((const double2*)cube128, (const double2*)ref128, window,
                                                               (const double2*)twiddle_s128, (double2*)rds128, ps, C_total,
                                                               chirp0, C_used, S, dc_removal);
    RS_CHECK_LAUNCH("rs_range_doppler_f64(range)");
    doppler_f64_kernel<<<(unsigned)(A * S), 128, sm_c, st>>>((double2*)rds128, (const double2*)twiddle_c128, pc, C_used);
    RS_CHECK_LAUNCH("rs_range_doppler_f64(doppler)");
    return RS_OK;
}

extern "C" int rs_detect_f64(const void* rds128, const uint8_t* gate, double threshold_db, double* power_db, int* row_count,
                             long long* row_offset, uint32_t* keys, long long cap, long long* total, int A, int R, int D,
                             void* stream) {
    RS_CHECK_ARG(rds128 && gate && power_db && row_count && row_offset && keys && total, "rs_detect_f64: null pointer");
    RS_CHECK_ARG(A > 0 && A <= RS_MAX_ANTENNAS && R > 0 && R <= RS_MAX_RANGE_BINS && D > 0 && D <= RS_MAX_DOPPLER_BINS && cap >= 0,
                 "rs_detect_f64: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)A * R * D, rows = (long long)A * R;
    power_db_c128_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 1 << 16), 256, 0, st>>>((const double2*)rds128,
                                                                                                  power_db, n);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    detect_rows_f64_kernel<false><<<grid, 256, 0, st>>>(power_db, gate, threshold_db, row_count, nullptr, nullptr, 0, A, R, D);
    scan_rows_kernel<<<1, 1024, 0, st>>>(row_count, row_offset, rows, total);
    detect_rows_f64_kernel<true><<<grid, 256, 0, st>>>(power_db, gate, threshold_db, row_count, row_offset, keys, cap, A, R, D);
    RS_CHECK_LAUNCH("rs_detect_f64");
    return RS_OK;
}

extern "C" int rs_signatures_c128(const void* rds128, const int* range_bin, const int* doppler_bin, int n, void* out128,
                                  int A, int R, int D, void* stream) {
    RS_CHECK_ARG(rds128 && out128 && n >= 0 && A > 0 && R > 0 && D > 0, "rs_signatures_c128: bad args");
    if (n == 0) return RS_OK;
    RS_CHECK_ARG(range_bin && doppler_bin, "rs_signatures_c128: null index list");
    signatures_c128_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const double2*)rds128, range_bin, doppler_bin,
                                                                             n, (double2*)out128, A, R, D);
    RS_CHECK_LAUNCH("rs_signatures_c128");
    return RS_OK;
}
