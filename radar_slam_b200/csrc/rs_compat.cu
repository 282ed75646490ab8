// fp64 helper kernels behind the legacy (numpy in / list-of-dict out) class API.  These keep the
// reference's fp64 output precision where the legacy return values are fp64 arrays; none of them is
// on the batched throughput path.
#include "rs_common.cuh"

namespace {

// 10 log10(|X|^2 + 1e-12) of the RDS rds[F][R][A][D], written in the reference layout [F][A][R][D]
// (dechirp.py:235-238, 277).
__global__ void power_db_kernel(const float2* __restrict__ rds, double* __restrict__ out, int R, int A, int D,
                                long long rows_total) {
    // one CTA per Doppler row (f, r, a) -> out row (f, a, r)
    for (long long row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const long long f = row / ((long long)R * A);
        const long long rem = row - f * R * A;
        const int r = (int)(rem / A), a = (int)(rem - (long long)r * A);
        const float2* src = rds + row * D;
        double* dst = out + ((f * A + a) * R + r) * D;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float2 x = src[d];
            const double p = (double)x.x * x.x + (double)x.y * x.y;
            dst[d] = 10.0 * log10(p + 1e-12);
        }
    }
}

// (x * conj(ref)) * w - mean per row, fp64 (dechirp.py:139, 108, 120 in that order); one CTA per row
__global__ void process_chirps_kernel(const double2* __restrict__ in, const double2* __restrict__ ref,
                                      const double* __restrict__ win, int S, int dc, double2* __restrict__ out) {
    __shared__ double sr[256], si[256];
    const double2* x = in + (size_t)blockIdx.x * S;
    double2* o = out + (size_t)blockIdx.x * S;
    double ar = 0, ai = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const double2 v = x[s], r = ref[s];
        double2 b = make_double2(v.x * r.x + v.y * r.y, v.y * r.x - v.x * r.y);   // v * conj(r)
        b.x *= win[s];
        b.y *= win[s];
        o[s] = b;
        ar += b.x;
        ai += b.y;
    }
    if (!dc) return;
    sr[threadIdx.x] = ar;
    si[threadIdx.x] = ai;
    __syncthreads();
    for (int h = blockDim.x >> 1; h; h >>= 1) {
        if (threadIdx.x < h) { sr[threadIdx.x] += sr[threadIdx.x + h]; si[threadIdx.x] += si[threadIdx.x + h]; }
        __syncthreads();
    }
    const double mr = sr[0] / S, mi = si[0] / S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double2 b = o[s];
        o[s] = make_double2(b.x - mr, b.y - mi);
    }
}

// ESPRIT closed form (angle_estimation.py:195-221) in fp64; one thread per snapshot
__global__ void esprit_f64_kernel(const double2* __restrict__ sig, int n, int M, double scale, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* s = sig + (size_t)i * M;
    double alpha = 0, gamma = 0, br = 0, bi = 0;
    for (int m = 0; m < M - 1; ++m) {
        const double2 x = s[m], y = s[m + 1];
        alpha += x.x * x.x + x.y * x.y;
        gamma += y.x * y.x + y.y * y.y;
        br += x.x * y.x + x.y * y.y;
        bi += x.x * y.y - x.y * y.x;
    }
    const double half = 0.5 * (alpha - gamma);
    const double lam = 0.5 * (alpha + gamma) + sqrt(half * half + br * br + bi * bi);
    double v0r, v0i, v1r, v1i;
    const double na = br * br + bi * bi + (lam - alpha) * (lam - alpha);
    const double nb = (lam - gamma) * (lam - gamma) + br * br + bi * bi;
    if (na >= nb) { v0r = br; v0i = bi; v1r = lam - alpha; v1i = 0; }
    else          { v0r = lam - gamma; v0i = 0; v1r = br; v1i = -bi; }
    double nr = 0, ni = 0, pr = 0, pi = 0;
    for (int m = 0; m < M - 1; ++m) {
        const double2 x = s[m], y = s[m + 1];
        const double ur = v0r * x.x - v0i * x.y + v1r * y.x - v1i * y.y;
        const double ui = v0r * x.y + v0i * x.x + v1r * y.y + v1i * y.x;
        if (m > 0) { nr += pr * ur + pi * ui; ni += pr * ui - pi * ur; }
        pr = ur; pi = ui;
    }
    out[i] = asin(atan2(ni, nr) * scale) * (180.0 / 3.14159265358979323846);
}

// ---- general 6-parameter bounded least squares for VelocitySolver.two_step_optimization ---------
// rows: k * [d_i, r_i x d_i] . [v, w] = y_i     (velocity_solver.py:84-113)
__device__ void design_row(const double* pos, const double* ang, double k, double* row) {
    const double az = ang[0], el = ang[1];
    const double d0 = cos(el) * cos(az), d1 = cos(el) * sin(az), d2 = sin(el);
    row[0] = k * d0; row[1] = k * d1; row[2] = k * d2;
    // (w x r).d = w.(r x d)
    row[3] = k * (pos[1] * d2 - pos[2] * d1);
    row[4] = k * (pos[2] * d0 - pos[0] * d2);
    row[5] = k * (pos[0] * d1 - pos[1] * d0);
}

// active-set bounded-variable least squares on the 6x6 normal equations (single thread)
__device__ void bvls6(const double (*G)[6], const double* b, const double* lo, const double* hi, double* x) {
    int st[6];   // 0 free, -1 at lo, +1 at hi, 2 unobservable, 3 pinned (lo == hi)
    double dmax = 0;
    for (int i = 0; i < 6; ++i) dmax = fmax(dmax, G[i][i]);
    for (int i = 0; i < 6; ++i) {
        x[i] = fmin(hi[i], fmax(lo[i], 0.0));
        st[i] = (G[i][i] <= 1e-13 * dmax || dmax == 0) ? 2 : 0;
        if (hi[i] <= lo[i]) st[i] = 3;      // pinned (step 1 pins omega at 0)
    }
    for (int iter = 0; iter < 64; ++iter) {
        int idx[6], nf = 0;
        for (int i = 0; i < 6; ++i) if (st[i] == 0) idx[nf++] = i;
        double M[6][7];
        for (int p = 0; p < nf; ++p) {
            double rhs = b[idx[p]];
            for (int j = 0; j < 6; ++j) if (st[j] != 0) rhs -= G[idx[p]][j] * x[j];
            for (int q = 0; q < nf; ++q) M[p][q] = G[idx[p]][idx[q]];
            M[p][nf] = rhs;
        }
        // Gauss-Jordan with partial pivoting; a vanishing pivot keeps that variable where it is
        double sol[6];
        bool sing[6] = {false, false, false, false, false, false};
        for (int c = 0; c < nf; ++c) {
            int piv = c;
            for (int r = c + 1; r < nf; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
            if (fabs(M[piv][c]) <= 1e-13 * dmax) { sing[c] = true; continue; }
            if (piv != c) for (int q = 0; q <= nf; ++q) { double t = M[c][q]; M[c][q] = M[piv][q]; M[piv][q] = t; }
            const double inv = 1.0 / M[c][c];
            for (int r = 0; r < nf; ++r) {
                if (r == c) continue;
                const double fct = M[r][c] * inv;
                if (fct != 0.0) for (int q = c; q <= nf; ++q) M[r][q] -= fct * M[c][q];
            }
        }
        for (int c = 0; c < nf; ++c) sol[c] = sing[c] ? x[idx[c]] : M[c][nf] / M[c][c];
        // step towards the free-set solution until the first bound is hit
        double alpha = 1.0;
        int hit = -1, side = 0;
        for (int p = 0; p < nf; ++p) {
            const int i = idx[p];
            const double dlt = sol[p] - x[i];
            if (sol[p] > hi[i] && dlt > 0) { const double a = (hi[i] - x[i]) / dlt; if (a < alpha) { alpha = a; hit = i; side = 1; } }
            if (sol[p] < lo[i] && dlt < 0) { const double a = (lo[i] - x[i]) / dlt; if (a < alpha) { alpha = a; hit = i; side = -1; } }
        }
        for (int p = 0; p < nf; ++p) x[idx[p]] += alpha * (sol[p] - x[idx[p]]);
        if (hit >= 0) { x[hit] = side > 0 ? hi[hit] : lo[hit]; st[hit] = side; continue; }
        // KKT check on the bound variables
        int rel = -1;
        double worst = 0;
        for (int i = 0; i < 6; ++i) {
            if (st[i] != 1 && st[i] != -1) continue;
            double g = -b[i];
            for (int j = 0; j < 6; ++j) g += G[i][j] * x[j];
            const double viol = st[i] == -1 ? -g : g;      // at lo need g >= 0, at hi need g <= 0
            if (viol > 1e-12 * (fabs(b[i]) + dmax) && viol > worst) { worst = viol; rel = i; }
        }
        if (rel < 0) break;
        st[rel] = 0;
    }
}

__global__ void velocity_ls6_kernel(const double* __restrict__ pos, const double* __restrict__ ang,
                                    const double* __restrict__ y, int n, double k, const double* __restrict__ lo,
                                    const double* __restrict__ hi, int nvar, double* __restrict__ out,
                                    double* __restrict__ pred) {
    // out: x[6], cost;   nvar = 3 (omega fixed at 0, step 1) or 6 (step 2)
    __shared__ double acc[28 * 8];
    __shared__ double xs[6];
    double loc[28];
    for (int q = 0; q < 28; ++q) loc[q] = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double row[6];
        design_row(pos + 3 * (size_t)i, ang + 2 * (size_t)i, k, row);
        int q = 0;
        for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) loc[q++] += row[a] * row[c];
        for (int a = 0; a < 6; ++a) loc[21 + a] += row[a] * y[i];
        loc[27] += y[i] * y[i];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = 0; q < 28; ++q) {
        double v = loc[q];
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) acc[q * 8 + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double G[6][6], b[6], l6[6], h6[6];
        int q = 0;
        for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) {
                double t = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += acc[q * 8 + w];
                G[a][c] = G[c][a] = t;
                ++q;
            }
        for (int a = 0; a < 6; ++a) {
            double t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += acc[(21 + a) * 8 + w];
            b[a] = t;
            l6[a] = a < nvar ? lo[a] : 0.0;
            h6[a] = a < nvar ? hi[a] : 0.0;
        }
        double x[6];
        bvls6(G, b, l6, h6, x);
        for (int a = 0; a < 6; ++a) { xs[a] = x[a]; out[a] = x[a]; }
    }
    __syncthreads();
    // predicted phases and the cost sum (y - X)^T (y - X)   (velocity_solver.py:171-174)
    double c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double row[6];
        design_row(pos + 3 * (size_t)i, ang + 2 * (size_t)i, k, row);
        double p = 0;
        for (int a = 0; a < 6; ++a) p += row[a] * xs[a];
        pred[i] = p;
        c += (y[i] - p) * (y[i] - p);
    }
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    __syncthreads();
    if (lane == 0) acc[wid] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += acc[w];
        out[6] = t;
    }
}


// RobustAngleEstimator.compute_angle_confidence (robust_angle_estimation.py:88-138) in fp64; one thread per target
__global__ void robust_confidence_kernel(const double2* __restrict__ sig, const double* __restrict__ angle_deg,
                                         const double* __restrict__ positions, double lambda_c, int n, int M,
                                         double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* s = sig + (size_t)i * M;
    const double st = sin(angle_deg[i] * (3.14159265358979323846 / 180.0));
    double cr = 0, ci = 0, sp = 0, perr = 0;
    double pw[RS_MAX_ANTENNAS];
    for (int m = 0; m < M; ++m) {
        const double ph = 2.0 * 3.14159265358979323846 * positions[m] * st / lambda_c;
        double sn, cs;
        sincos(ph, &sn, &cs);
        const double2 x = s[m];
        const double zr = cs * x.x + sn * x.y;        // conj(a) * s
        const double zi = cs * x.y - sn * x.x;
        cr += zr; ci += zi;
        const double p = x.x * x.x + x.y * x.y;
        sp += p;
        perr += fabs(atan2(zi, zr));                  // |wrap(arg s - arg a)|
        // insertion sort for the 20th percentile
        int j = m;
        while (j > 0 && pw[j - 1] > p) { pw[j] = pw[j - 1]; --j; }
        pw[j] = p;
    }
    const double corr = sqrt(cr * cr + ci * ci);
    const double ncorr = sp > 0 ? corr / sqrt(sp) : 0.0;
    const double pcons = exp(-perr / M);
    const double h = 0.2 * (M - 1);
    const int lo = (int)floor(h);
    const double g = h - lo;
    const double a = pw[lo], b = pw[lo + 1 < M ? lo + 1 : lo];
    double nf = a + (b - a) * g;
    if (g >= 0.5) nf = b - (b - a) * (1.0 - g);
    double snr_c = 0.0;
    if (nf > 0) snr_c = fmin(1.0, log10((sp / M) / nf) / 3.0);
    const double c = ncorr * 0.4 + pcons * 0.3 + snr_c * 0.3;
    out[i] = fmin(1.0, fmax(0.0, c));
}

}  // namespace

extern "C" int rs_power_db_f64(const void* rds, double* out, int F, int A, int C, int S, void* stream) {
    RS_CHECK_ARG(rds && out && F > 0 && F <= 65535 && A > 0 && C > 0 && S > 0, "rs_power_db_f64: bad args");
    const long long rows = (long long)F * S * A;
    const unsigned grid = (unsigned)(rows < (1ll << 20) ? rows : (1ll << 20));
    power_db_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float2*)rds, out, S, A, C, rows);
    RS_CHECK_LAUNCH("rs_power_db_f64");
    return RS_OK;
}

extern "C" int rs_process_chirps_f64(const void* in128, const void* ref128, const double* window, int rows, int S,
                                     int dc_removal, void* out128, void* stream) {
    RS_CHECK_ARG(in128 && ref128 && window && out128 && rows > 0 && S > 0, "rs_process_chirps_f64: bad args");
    process_chirps_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const double2*)in128, (const double2*)ref128, window, S,
                                                                 dc_removal, (double2*)out128);
    RS_CHECK_LAUNCH("rs_process_chirps_f64");
    return RS_OK;
}

extern "C" int rs_esprit_f64(const void* sig128, int n, int A, double esprit_scale, double* out, void* stream) {
    RS_CHECK_ARG(sig128 && out && n >= 0 && A >= 2, "rs_esprit_f64: bad args");
    if (n == 0) return RS_OK;
    esprit_f64_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const double2*)sig128, n, A, esprit_scale, out);
    RS_CHECK_LAUNCH("rs_esprit_f64");
    return RS_OK;
}

extern "C" int rs_velocity_ls6(const double* pos, const double* ang, const double* y, int n, double k_phase,
                               const double* lo, const double* hi, int nvar, double* out7, double* pred, void* stream) {
    RS_CHECK_ARG(pos && ang && y && lo && hi && out7 && pred && n > 0, "rs_velocity_ls6: bad args");
    RS_CHECK_ARG(nvar == 3 || nvar == 6, "rs_velocity_ls6: nvar must be 3 or 6");
    velocity_ls6_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pos, ang, y, n, k_phase, lo, hi, nvar, out7, pred);
    RS_CHECK_LAUNCH("rs_velocity_ls6");
    return RS_OK;
}

extern "C" int rs_robust_confidence_f64(const void* sig128, const double* angle_deg, const double* positions,
                                        double lambda_c, int n, int A, double* out, void* stream) {
    RS_CHECK_ARG(sig128 && angle_deg && positions && out && n >= 0 && A >= 1 && A <= RS_MAX_ANTENNAS,
                 "rs_robust_confidence_f64: bad args");
    if (n == 0) return RS_OK;
    robust_confidence_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const double2*)sig128, angle_deg, positions,
                                                                           lambda_c, n, A, out);
    RS_CHECK_LAUNCH("rs_robust_confidence_f64");
    return RS_OK;
}
