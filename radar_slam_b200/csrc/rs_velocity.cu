// Ego-velocity least squares.
//
// Replaces VelocitySolver.solve_velocity / two_step_optimization (velocity_solver.py:178-355).
// The reference minimises sum (y - 4 pi dt/lambda * d.(v + w x r))^2 with differential evolution.
// Elevation is hard-wired to 0 and the target position is range * direction (velocity_solver.py:
// 334-342), so (w x r).d == 0 and the v_z column is zero: the cost is a convex quadratic in
// (v_x, v_y) only and DE converges onto its minimiser inside the box |v| <= 50 (:216, :250).
// One CTA per frame accumulates the six normal-equation sums in fp64 in a fixed order
// (deterministic), then one thread solves the 2x2 system with the box active-set enumeration.
#include "rs_common.cuh"

namespace {

constexpr int VEL_THREADS = 256;

struct Sums {
    double cc, ss, cs, yc, ys, yy, n;
};

__device__ __forceinline__ void block_reduce(Sums& s, double* sh) {
    double v[7] = {s.cc, s.ss, s.cs, s.yc, s.ys, s.yy, s.n};
#pragma unroll
    for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int off = 16; off; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0)
        for (int q = 0; q < 7; ++q) sh[wid * 7 + q] = v[q];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < 7; ++q) {
            double t = 0;
            for (int w = 0; w < VEL_THREADS / 32; ++w) t += sh[w * 7 + q];
            sh[q + 64] = t;
        }
    }
    __syncthreads();
    s.cc = sh[64]; s.ss = sh[65]; s.cs = sh[66]; s.yc = sh[67]; s.ys = sh[68]; s.yy = sh[69]; s.n = sh[70];
}

__device__ __forceinline__ double clampd(double x, double b) { return fmin(b, fmax(-b, x)); }

// exact minimiser of  g11 vx^2 + 2 g12 vx vy + g22 vy^2 - 2 (b1 vx + b2 vy)  over |vx|,|vy| <= bound
__device__ void box_ls(double g11, double g22, double g12, double b1, double b2, double bound, double* vx_out,
                       double* vy_out) {
    const double det = g11 * g22 - g12 * g12;
    const bool ok = det > 1e-12 * fmax(g11 * g22, 1e-300);
    if (ok) {
        const double vx = (g22 * b1 - g12 * b2) / det, vy = (g11 * b2 - g12 * b1) / det;
        if (fabs(vx) <= bound && fabs(vy) <= bound) { *vx_out = vx; *vy_out = vy; return; }
    }
    double bestf = 1.0e300, bx = 0, by = 0;
    auto consider = [&](double vx, double vy) {
        const double fv = g11 * vx * vx + 2 * g12 * vx * vy + g22 * vy * vy - 2 * (b1 * vx + b2 * vy);
        if (fv < bestf) { bestf = fv; bx = vx; by = vy; }
    };
    for (int sgn = -1; sgn <= 1; sgn += 2) {
        const double vx = sgn * bound;
        consider(vx, clampd(g22 > 0 ? (b2 - g12 * vx) / g22 : 0.0, bound));
    }
    for (int sgn = -1; sgn <= 1; sgn += 2) {
        const double vy = sgn * bound;
        consider(clampd(g11 > 0 ? (b1 - g12 * vy) / g11 : 0.0, bound), vy);
    }
    if (!ok) {
        const double tr = g11 + g22;
        if (tr > 0) consider(clampd(b1 / tr, bound), clampd(b2 / tr, bound));
        else consider(0.0, 0.0);
    }
    *vx_out = bx; *vy_out = by;
}

__global__ void __launch_bounds__(VEL_THREADS)
velocity_kernel(const int32_t* __restrict__ det_aidx, const float* __restrict__ det_adeg,
                const float* __restrict__ det_phase, const uint8_t* __restrict__ det_flags,
                const int32_t* __restrict__ det_count, const double* __restrict__ grid_cs, double kph, double bound,
                int irls_iters, double huber, double* __restrict__ vel, int seg_cap, int nseg) {
    __shared__ double sh[80];
    __shared__ double v_sh[2];
    const int f = blockIdx.x;
    double vx = 0, vy = 0;
    Sums tot{};
    for (int it = 0; it <= irls_iters; ++it) {
        Sums s{};
        for (int sg = 0; sg < nseg; ++sg) {
            const size_t seg = (size_t)f * nseg + sg;
            const int n = det_count[seg];
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const size_t o = seg * seg_cap + i;
                if (det_flags[o] & RS_FLAG_DROPPED) continue;
                double c, sn;
                const int ai = det_aidx[o];
                if (grid_cs != nullptr && ai >= 0) { c = grid_cs[2 * ai]; sn = grid_cs[2 * ai + 1]; }
                else { sincos((double)det_adeg[o] * (3.14159265358979323846 / 180.0), &sn, &c); }
                const double y = (double)det_phase[o];
                double w = 1.0;
                if (it > 0) {
                    const double res = fabs(y - kph * (vx * c + vy * sn));
                    w = res > huber ? huber / res : 1.0;
                }
                s.cc += w * c * c; s.ss += w * sn * sn; s.cs += w * c * sn;
                s.yc += w * y * c; s.ys += w * y * sn; s.yy += w * y * y; s.n += 1.0;
            }
        }
        block_reduce(s, sh);
        if (threadIdx.x == 0) {
            double ox = 0, oy = 0;
            if (s.n >= 3.0)
                box_ls(kph * kph * s.cc, kph * kph * s.ss, kph * kph * s.cs, kph * s.yc, kph * s.ys, bound, &ox, &oy);
            v_sh[0] = ox; v_sh[1] = oy;
        }
        __syncthreads();
        vx = v_sh[0]; vy = v_sh[1];
        tot = s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double* o = vel + (size_t)f * 8;
        const bool ok = tot.n >= 3.0;       // velocity_solver.py:202-204
        o[0] = ok ? vx : 0.0; o[1] = ok ? vy : 0.0; o[2] = 0.0;
        o[3] = 0.0; o[4] = 0.0; o[5] = 0.0;
        o[6] = ok ? 1.0 : 0.0; o[7] = tot.n;
    }
}

// The normal-equation sums of ONE segment from its detection list (what rs_angles fuses into its scan for the grid
// methods at A <= 16): one CTA per segment, fixed reduction order -> deterministic.  ESPRIT and A > 16 go through here,
// so a frame's millions of detections are summed by all SMs instead of by the one CTA of velocity_kernel.
__global__ void __launch_bounds__(VEL_THREADS)
velocity_partials_kernel(const int32_t* __restrict__ det_aidx, const float* __restrict__ det_adeg,
                         const float* __restrict__ det_phase, const uint8_t* __restrict__ det_flags,
                         const int32_t* __restrict__ det_count, const double* __restrict__ grid_cs,
                         double* __restrict__ partials, int seg_cap) {
    __shared__ double sh[80];
    const size_t seg = blockIdx.x;
    const int n = det_count[seg];
    Sums s{};
    // Thread t takes the contiguous entries [t per, t per + per): the entries of a cell are contiguous and share their
    // angle, so off the grid (ESPRIT: an fp64 sincos per angle) the previous entry's (cos, sin) is reused while the angle
    // repeats -- one sincos per cell and thread instead of one per detection (17 detections per cell at 192 channels:
    // 0.24 -> 0.1 ms per 8 frames of 512 x 256 x 192).
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int i0 = (int)threadIdx.x * per, i1 = min(n, i0 + per);
    float last = 0.f;
    bool have = false;
    double c = 1.0, sn = 0.0;
    for (int i = i0; i < i1; ++i) {
        const size_t o = seg * seg_cap + i;
        if (det_flags[o] & RS_FLAG_DROPPED) continue;
        const int ai = det_aidx[o];
        if (grid_cs != nullptr && ai >= 0) {
            c = grid_cs[2 * ai]; sn = grid_cs[2 * ai + 1];
            have = false;
        } else {
            const float a = det_adeg[o];
            if (!have || a != last) {
                sincos((double)a * (3.14159265358979323846 / 180.0), &sn, &c);
                last = a;
                have = true;
            }
        }
        const double y = (double)det_phase[o];
        s.cc += c * c; s.ss += sn * sn; s.cs += c * sn;
        s.yc += y * c; s.ys += y * sn; s.yy += y * y; s.n += 1.0;
    }
    block_reduce(s, sh);
    if (threadIdx.x == 0) {
        double* q = partials + seg * 8;
        q[0] = s.cc; q[1] = s.ss; q[2] = s.cs; q[3] = s.yc; q[4] = s.ys; q[5] = s.yy; q[6] = s.n; q[7] = 0.0;
    }
}

// One warp per frame: add the per-segment partial sums written by rs_angles in segment order
// (lane-strided, then a fixed shuffle tree: deterministic) and solve.
__global__ void __launch_bounds__(128)
velocity_from_partials_kernel(const double* __restrict__ partials, int nseg, int F, double kph, double bound,
                              const int32_t* __restrict__ overflow, double* __restrict__ vel) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= F) return;
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int sg = lane; sg < nseg; sg += 32) {
        const double* q = partials + ((size_t)f * nseg + sg) * 8;
#pragma unroll
        for (int j = 0; j < 7; ++j) s[j] += q[j];
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int off = 16; off; off >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
    }
    if (lane == 0) {
        double vx = 0, vy = 0;
        const bool ok = s[6] >= 3.0;
        if (ok) box_ls(kph * kph * s[0], kph * kph * s[1], kph * kph * s[2], kph * s[3], kph * s[4], bound, &vx, &vy);
        double* o = vel + (size_t)f * 8;
        o[0] = vx; o[1] = vy; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 0.0;
        // a frame whose detection segments overflowed (rs_detect: det_overflow) was solved from a truncated list: the row
        // carries the best-effort solution but reports failure, so no caller can mistake it for the full answer
        const bool complete = overflow == nullptr || overflow[f] == 0;
        o[6] = (ok && complete) ? 1.0 : 0.0; o[7] = s[6];
    }
}

}  // namespace

extern "C" int rs_velocity_from_partials(const double* ls_partials, int nseg_per_frame, int F, double k_phase,
                                         double bound, const int32_t* det_overflow, double* vel, void* stream) {
    RS_CHECK_ARG(ls_partials && vel && nseg_per_frame > 0 && F > 0 && bound > 0, "rs_velocity_from_partials: bad args");
    velocity_from_partials_kernel<<<(F + 3) / 4, 128, 0, (cudaStream_t)stream>>>(ls_partials, nseg_per_frame, F, k_phase,
                                                                              bound, det_overflow, vel);
    RS_CHECK_LAUNCH("rs_velocity_from_partials");
    return RS_OK;
}

extern "C" int rs_velocity_partials(const int32_t* det_aidx, const float* det_adeg, const float* det_phase,
                                    const uint8_t* det_flags, const int32_t* det_count, const double* grid_cs,
                                    double* ls_partials, int seg_cap, int nseg_per_frame, int F, void* stream) {
    RS_CHECK_ARG(det_aidx && det_adeg && det_phase && det_flags && det_count && ls_partials, "rs_velocity_partials: null pointer");
    RS_CHECK_ARG(F > 0 && seg_cap > 0 && nseg_per_frame > 0 && (long long)F * nseg_per_frame < (1ll << 31),
                 "rs_velocity_partials: bad args");
    velocity_partials_kernel<<<(unsigned)((long long)F * nseg_per_frame), VEL_THREADS, 0, (cudaStream_t)stream>>>(
        det_aidx, det_adeg, det_phase, det_flags, det_count, grid_cs, ls_partials, seg_cap);
    RS_CHECK_LAUNCH("rs_velocity_partials");
    return RS_OK;
}

extern "C" int rs_velocity_ls(const int32_t* det_aidx, const float* det_adeg, const float* det_phase,
                              const uint8_t* det_flags, const int32_t* det_count, const double* grid_cs, double k_phase,
                              double bound, int irls_iters, double huber_delta, double* vel, int seg_cap,
                              int nseg_per_frame, int F, void* stream) {
    RS_CHECK_ARG(det_aidx && det_adeg && det_phase && det_flags && det_count && vel, "rs_velocity_ls: null pointer");
    RS_CHECK_ARG(F > 0 && seg_cap > 0 && nseg_per_frame > 0 && irls_iters >= 0 && bound > 0, "rs_velocity_ls: bad args");
    RS_CHECK_ARG(irls_iters == 0 || huber_delta > 0, "rs_velocity_ls: huber_delta must be > 0");
    velocity_kernel<<<F, VEL_THREADS, 0, (cudaStream_t)stream>>>(det_aidx, det_adeg, det_phase, det_flags, det_count,
                                                                 grid_cs, k_phase, bound, irls_iters, huber_delta, vel,
                                                                 seg_cap, nseg_per_frame);
    RS_CHECK_LAUNCH("rs_velocity_ls");
    return RS_OK;
}
