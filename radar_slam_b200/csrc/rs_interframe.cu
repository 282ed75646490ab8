// Inter-frame ego-velocity path: ImprovedVelocitySolver
// (src/algorithms/velocity_solver_improved.py).
//
//   rs_associate_targets      greedy nearest-neighbour association of the targets of consecutive frames
//                             (velocity_solver_improved.py:74-129), batched over frame pairs.
//   rs_wrapped_cost           the solver's cost  sum_i wrap(y_i - x_i(v, w))^2 + 0.01 |v|^2 + 0.01 |w|^2
//                             (:223-266) in fp64 for a batch of candidate motions.
//   rs_wrapped_lattice_search global search of that cost over the (v_x, v_y) box.  With elevation 0 and positions
//                             range * direction the model is  x_i = k (v_x cos az_i + v_y sin az_i)  -- v_z and w do
//                             not enter and the regulariser pins them to 0 -- and the residual is wrapped to (-pi, pi],
//                             so the cost has a local minimum every 2 pi / k (2 cm/s at dt = 0.1 s, 77 GHz) in every
//                             target direction: differential_evolution (:389-420) lands in one of ~1e7 of them.  The
//                             kernel evaluates the cost on a lattice finer than the basin (step 2 pi / (6 k)) over the
//                             whole box -- ~1e9 points x N targets, fp32 with an fp64 re-base of the phase every 32
//                             points -- and returns the best point of every tile; the caller polishes the best tiles
//                             in fp64 (Gauss-Newton inside the basin) and keeps the smallest cost.
#include "rs_common.cuh"

namespace {

constexpr int IF_THREADS = 256;

// ---- association ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IF_THREADS)
associate_kernel(const double2* __restrict__ cur_xy, const int32_t* __restrict__ n_cur, const double2* __restrict__ prev_xy,
                 const int32_t* __restrict__ n_prev, double threshold, int32_t* __restrict__ match,
                 double* __restrict__ dist, int nc_max, int np_max) {
    extern __shared__ unsigned char used[];         // [np_max]
    __shared__ double red_d[IF_THREADS / 32];
    __shared__ int red_j[IF_THREADS / 32];
    __shared__ int win_j;
    const int pair = blockIdx.x;
    const int nc = min(n_cur[pair], nc_max), np = min(n_prev[pair], np_max);
    const double2* c = cur_xy + (size_t)pair * nc_max;
    const double2* p = prev_xy + (size_t)pair * np_max;
    for (int j = threadIdx.x; j < np; j += blockDim.x) used[j] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = 0; i < nc; ++i) {
        const double2 ci = c[i];
        double best = INFINITY;
        int bj = 0x7fffffff;
        for (int j = threadIdx.x; j < np; j += blockDim.x) {
            if (used[j]) continue;
            const double dx = ci.x - p[j].x, dy = ci.y - p[j].y;
            const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // scipy cdist (euclidean), no FMA contraction
            if (d < threshold && d < best) { best = d; bj = j; }           // :111-113, j ascending: first minimum wins
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
        }
        if (lane == 0) { red_d[wid] = best; red_j[wid] = bj; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = red_d[0];
            int j = red_j[0];
            for (int w = 1; w < IF_THREADS / 32; ++w)
                if (red_d[w] < b || (red_d[w] == b && red_j[w] < j)) { b = red_d[w]; j = red_j[w]; }
            const bool ok = j != 0x7fffffff;
            match[(size_t)pair * nc_max + i] = ok ? j : -1;
            dist[(size_t)pair * nc_max + i] = ok ? b : INFINITY;
            win_j = ok ? j : -1;
        }
        __syncthreads();
        if (threadIdx.x == 0 && win_j >= 0) used[win_j] = 1;
        __syncthreads();
    }
    for (int i = nc + threadIdx.x; i < nc_max; i += blockDim.x) {
        match[(size_t)pair * nc_max + i] = -1;
        dist[(size_t)pair * nc_max + i] = INFINITY;
    }
}

// ---- fp64 cost of a batch of candidate motions: one warp per candidate --------------------------------------------
__global__ void __launch_bounds__(IF_THREADS)
wrapped_cost_kernel(const double* __restrict__ params, const double* __restrict__ pos, const double* __restrict__ ang,
                    const double* __restrict__ y, double k_phase, double reg_v, double reg_w, double* __restrict__ cost,
                    int n, int nq) {
    const int q = blockIdx.x * (IF_THREADS / 32) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    const double* m = params + (size_t)q * 6;
    const double vx = m[0], vy = m[1], vz = m[2], wx = m[3], wy = m[4], wz = m[5];
    double acc = 0;
    for (int i = lane; i < n; i += 32) {
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        const double az = ang[2 * i], el = ang[2 * i + 1];
        const double dx = cos(el) * cos(az), dy = cos(el) * sin(az), dz = sin(el);       // :195-199
        const double rx = vx + (wy * pz - wz * py), ry = vy + (wz * px - wx * pz), rz = vz + (wx * py - wy * px);   // v + w x r
        const double pred = k_phase * (rx * dx + ry * dy + rz * dz);                      // :211
        const double r = y[i] - pred;
        const double wr = atan2(sin(r), cos(r));                                          // :250
        acc += wr * wr;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0)
        cost[q] = acc + reg_v * (vx * vx + vy * vy + vz * vz) + reg_w * (wx * wx + wy * wy + wz * wz);   // :255-260
}

// ---- AdvancedVelocityOptimizer.compute_regularized_cost_function (advanced_velocity_optimization.py:153-223) ----------
// wrapped residual sum + five regularisers, fp64, one warp per candidate motion
__global__ void __launch_bounds__(IF_THREADS)
regularized_cost_kernel(const double* __restrict__ params, const double* __restrict__ pos, const double* __restrict__ ang,
                        const double* __restrict__ y, double k_phase, double max_v, double max_w, double weight,
                        const double* __restrict__ previous, double* __restrict__ cost, int n, int nq) {
    const int q = blockIdx.x * (IF_THREADS / 32) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    const double* m = params + (size_t)q * 6;
    const double vx = m[0], vy = m[1], vz = m[2], wx = m[3], wy = m[4], wz = m[5];
    double acc = 0;
    for (int i = lane; i < n; i += 32) {
        const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        const double az = ang[2 * i], el = ang[2 * i + 1];
        const double dx = cos(el) * cos(az), dy = cos(el) * sin(az), dz = sin(el);       // :239-243
        const double rx = vx + (wy * pz - wz * py), ry = vy + (wz * px - wx * pz), rz = vz + (wx * py - wy * px);   // :246-247
        const double pred = k_phase * (rx * dx + ry * dy + rz * dz);                      // :253
        const double r = y[i] - pred;
        const double wr = atan2(sin(r), cos(r));                                          // :185
        acc += wr * wr;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        double reg = 0;
        const double vm = sqrt(vx * vx + vy * vy + vz * vz), wm = sqrt(wx * wx + wy * wy + wz * wz);
        if (vm > max_v * 0.8) reg += weight * (vm - max_v * 0.8) * (vm - max_v * 0.8);     // 1. speed           :194-196
        if (wm > max_w * 0.8) reg += weight * (wm - max_w * 0.8) * (wm - max_w * 0.8);     // 2. rotation rate   :199-201
        if (previous != nullptr) {                                                        // 3. temporal        :204-207
            double t = 0;
            for (int j = 0; j < 6; ++j) t += (m[j] - previous[j]) * (m[j] - previous[j]);
            reg += weight * 0.1 * t;
        }
        if (vm > 20 && wm > 5) reg += weight * 0.01 * (vm - 20) * (wm - 5);               // 4. combination     :211-213
        reg += weight * 10.0 * vz * vz;                                                   // 5. vertical motion :216-217
        cost[q] = acc + reg;
    }
}

// ---- Gauss-Newton polish inside a basin of the wrapped cost -------------------------------------------------------------
// minimise  sum_i wrap(y_i - k (v_x c_i + v_y s_i))^2 + reg |v - centre|^2  from a batch of starting points: inside a
// basin the wrapped residual is linear in v, so the normal matrix H = G^T G + reg I is constant.  One warp per start.
__global__ void __launch_bounds__(IF_THREADS)
gn_polish_kernel(const double* __restrict__ c, const double* __restrict__ s, const double* __restrict__ y, int n, double k,
                 double reg, double cx, double cy, double lo_x, double hi_x, double lo_y, double hi_y,
                 double* __restrict__ v, double* __restrict__ cost, int nq, int iters) {
    const int q = blockIdx.x * (IF_THREADS / 32) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    double hxx = 0, hxy = 0, hyy = 0;
    for (int i = lane; i < n; i += 32) {
        const double gx = k * c[i], gy = k * s[i];
        hxx += gx * gx; hxy += gx * gy; hyy += gy * gy;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        hxx += __shfl_xor_sync(0xffffffffu, hxx, off);
        hxy += __shfl_xor_sync(0xffffffffu, hxy, off);
        hyy += __shfl_xor_sync(0xffffffffu, hyy, off);
    }
    hxx += reg; hyy += reg;
    double det = hxx * hyy - hxy * hxy;
    if (!(fabs(det) > 1e-300)) det = 1e-300;
    double vx = v[2 * q], vy = v[2 * q + 1];
    double f = 0;
    for (int it = 0; it <= iters; ++it) {
        double gx = 0, gy = 0;
        f = 0;
        for (int i = lane; i < n; i += 32) {
            const double r0 = y[i] - k * (vx * c[i] + vy * s[i]);
            const double r = atan2(sin(r0), cos(r0));
            gx += k * c[i] * r; gy += k * s[i] * r;
            f += r * r;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, off);
            gy += __shfl_xor_sync(0xffffffffu, gy, off);
            f += __shfl_xor_sync(0xffffffffu, f, off);
        }
        f += reg * ((vx - cx) * (vx - cx) + (vy - cy) * (vy - cy));
        if (it == iters) break;
        gx -= reg * (vx - cx);
        gy -= reg * (vy - cy);
        const double sx = (hyy * gx - hxy * gy) / det, sy = (hxx * gy - hxy * gx) / det;
        vx = fmin(hi_x, fmax(lo_x, vx + sx));
        vy = fmin(hi_y, fmax(lo_y, vy + sy));
        if (fmax(fabs(sx), fabs(sy)) < 1e-13) iters = it + 1;       // converged: one more pass evaluates the cost
    }
    if (lane == 0) { v[2 * q] = vx; v[2 * q + 1] = vy; cost[q] = f; }
}

// ---- lattice search ----------------------------------------------------------------------------------------------
constexpr int LS_L = 32;          // consecutive v_x points per thread
constexpr int LS_ROWS = 8;        // v_y rows per CTA (one warp each)
constexpr int LS_MAX_N = 2048;

__global__ void __launch_bounds__(IF_THREADS)
lattice_search_kernel(const double* __restrict__ ax, const double* __restrict__ by, const double* __restrict__ yc, int n,
                      double vx_lo, double vy_lo, double h, long long nx, long long ny, double reg,
                      float* __restrict__ tile_cost, int32_t* __restrict__ tile_ix, int32_t* __restrict__ tile_iy) {
    // phases in CYCLES: yc_i = y_i / 2 pi, ax_i = k cos(az_i) / 2 pi, by_i = k sin(az_i) / 2 pi
    extern __shared__ double tgt[];                  // [3][n]
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x)
        tgt[i] = i < n ? ax[i] : (i < 2 * n ? by[i - n] : yc[i - 2 * n]);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long iy = (long long)blockIdx.y * LS_ROWS + wid;
    const long long ix0 = ((long long)blockIdx.x * 32 + lane) * LS_L;
    float best = 3.0e38f;
    int bix = 0;
    if (iy < ny && ix0 < nx) {
        const double vy = vy_lo + (double)iy * h, vx0 = vx_lo + (double)ix0 * h;
        float acc[LS_L];
#pragma unroll
        for (int l = 0; l < LS_L; ++l) acc[l] = 0.f;
        const float MAGIC = 12582912.f;              // 1.5 * 2^23: (x + M) - M = rint(x) for |x| < 2^22
        for (int i = 0; i < n; ++i) {
            // fp64 re-base of the phase at the first point of the segment, reduced to [-0.5, 0.5] cycles
            double b = tgt[2 * n + i] - tgt[i] * vx0 - tgt[n + i] * vy;
            b -= rint(b);
            const float base = (float)b, step = (float)(-tgt[i] * h);
#pragma unroll
            for (int l = 0; l < LS_L; ++l) {
                const float ph = fmaf((float)l, step, base);
                const float r = ph - ((ph + MAGIC) - MAGIC);
                acc[l] = fmaf(r, r, acc[l]);
            }
        }
        const float four_pi2 = 39.47841760435743f;
#pragma unroll
        for (int l = 0; l < LS_L; ++l) {
            if (ix0 + l < nx) {
                const double vx = vx0 + (double)l * h;
                const float c = fmaf(four_pi2, acc[l], (float)(reg * (vx * vx + vy * vy)));
                if (c < best) { best = c; bix = l; }
            }
        }
    }
    // tile = the whole CTA (8 rows x 1024 columns): keep its best point
    long long gix = ix0 + bix, giy = iy;
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const long long ox = __shfl_xor_sync(0xffffffffu, gix, off);
        if (ob < best || (ob == best && ox < gix)) { best = ob; gix = ox; }
    }
    __shared__ float s_c[LS_ROWS];
    __shared__ long long s_x[LS_ROWS], s_y[LS_ROWS];
    if (lane == 0) { s_c[wid] = best; s_x[wid] = gix; s_y[wid] = giy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int w0 = 0;
        for (int w = 1; w < LS_ROWS; ++w)
            if (s_c[w] < s_c[w0]) w0 = w;
        const size_t t = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        tile_cost[t] = s_c[w0];
        tile_ix[t] = (int32_t)s_x[w0];
        tile_iy[t] = (int32_t)s_y[w0];
    }
}

}  // namespace

extern "C" int rs_associate_targets(const double* cur_xy, const int32_t* n_cur, const double* prev_xy,
                                    const int32_t* n_prev, double threshold, int32_t* match_idx, double* match_dist,
                                    int pairs, int nc_max, int np_max, void* stream) {
    RS_CHECK_ARG(cur_xy && n_cur && prev_xy && n_prev && match_idx && match_dist, "rs_associate_targets: null pointer");
    RS_CHECK_ARG(pairs > 0 && nc_max > 0 && np_max > 0 && np_max <= 128 * 1024, "rs_associate_targets: bad dims");
    cudaFuncSetAttribute(associate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, np_max);
    associate_kernel<<<pairs, IF_THREADS, (size_t)np_max, (cudaStream_t)stream>>>(
        (const double2*)cur_xy, n_cur, (const double2*)prev_xy, n_prev, threshold, match_idx, match_dist, nc_max, np_max);
    RS_CHECK_LAUNCH("rs_associate_targets");
    return RS_OK;
}

extern "C" int rs_wrapped_cost(const double* params, const double* pos, const double* ang, const double* y, int n, int nq,
                               double k_phase, double reg_v, double reg_w, double* cost, void* stream) {
    RS_CHECK_ARG(params && pos && ang && y && cost && n >= 0 && nq > 0, "rs_wrapped_cost: bad args");
    const int wpb = IF_THREADS / 32;
    wrapped_cost_kernel<<<(nq + wpb - 1) / wpb, IF_THREADS, 0, (cudaStream_t)stream>>>(params, pos, ang, y, k_phase, reg_v,
                                                                                      reg_w, cost, n, nq);
    RS_CHECK_LAUNCH("rs_wrapped_cost");
    return RS_OK;
}

extern "C" int rs_regularized_cost(const double* params, const double* pos, const double* ang, const double* y, int n, int nq,
                                   double k_phase, double max_velocity, double max_angular_velocity, double weight,
                                   const double* previous_motion, double* cost, void* stream) {
    RS_CHECK_ARG(params && pos && ang && y && cost && n >= 0 && nq > 0, "rs_regularized_cost: bad args");
    const int wpb = IF_THREADS / 32;
    regularized_cost_kernel<<<(nq + wpb - 1) / wpb, IF_THREADS, 0, (cudaStream_t)stream>>>(
        params, pos, ang, y, k_phase, max_velocity, max_angular_velocity, weight, previous_motion, cost, n, nq);
    RS_CHECK_LAUNCH("rs_regularized_cost");
    return RS_OK;
}

extern "C" int rs_wrapped_gn_polish(const double* cos_az, const double* sin_az, const double* y, int n, double k_phase,
                                    double reg, double centre_x, double centre_y, double lo_x, double hi_x, double lo_y,
                                    double hi_y, double* v_xy, double* cost, int nq, int iters, void* stream) {
    RS_CHECK_ARG(cos_az && sin_az && y && v_xy && cost && n > 0 && nq > 0 && iters >= 0 && reg >= 0, "rs_wrapped_gn_polish: bad args");
    const int wpb = IF_THREADS / 32;
    gn_polish_kernel<<<(nq + wpb - 1) / wpb, IF_THREADS, 0, (cudaStream_t)stream>>>(
        cos_az, sin_az, y, n, k_phase, reg, centre_x, centre_y, lo_x, hi_x, lo_y, hi_y, v_xy, cost, nq, iters);
    RS_CHECK_LAUNCH("rs_wrapped_gn_polish");
    return RS_OK;
}

extern "C" int rs_wrapped_lattice_tiles(long long nx, long long ny, int* tiles_x, int* tiles_y) {
    RS_CHECK_ARG(nx > 0 && ny > 0 && tiles_x && tiles_y, "rs_wrapped_lattice_tiles: bad args");
    *tiles_x = (int)((nx + 32 * LS_L - 1) / (32 * LS_L));
    *tiles_y = (int)((ny + LS_ROWS - 1) / LS_ROWS);
    return RS_OK;
}

extern "C" int rs_wrapped_lattice_search(const double* ax_cycles, const double* by_cycles, const double* y_cycles, int n,
                                         double vx_lo, double vy_lo, double h, long long nx, long long ny, double reg,
                                         float* tile_cost, int32_t* tile_ix, int32_t* tile_iy, void* stream) {
    RS_CHECK_ARG(ax_cycles && by_cycles && y_cycles && tile_cost && tile_ix && tile_iy, "rs_wrapped_lattice_search: null pointer");
    RS_CHECK_ARG(n > 0 && n <= LS_MAX_N && h > 0 && nx > 0 && ny > 0 && nx < (1ll << 31) && ny < (1ll << 31),
                 "rs_wrapped_lattice_search: bad dims (1 <= n <= %d)", LS_MAX_N);
    int tx = 0, ty = 0;
    rs_wrapped_lattice_tiles(nx, ny, &tx, &ty);
    RS_CHECK_ARG(ty <= 65535, "rs_wrapped_lattice_search: too many lattice rows (%lld)", ny);
    const size_t smem = (size_t)3 * n * sizeof(double);
    cudaFuncSetAttribute(lattice_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lattice_search_kernel<<<dim3((unsigned)tx, (unsigned)ty), IF_THREADS, smem, (cudaStream_t)stream>>>(
        ax_cycles, by_cycles, y_cycles, n, vx_lo, vy_lo, h, nx, ny, reg, tile_cost, tile_ix, tile_iy);
    RS_CHECK_LAUNCH("rs_wrapped_lattice_search");
    return RS_OK;
}
