// fp64 re-evaluation of the decisions the fp32 path flags as undecidable.
//
// The reference works in fp64 end to end.  From complex64 inputs the fp32 spectra carry ~1e-7 relative
// error, so two kinds of decisions cannot be settled in fp32 and are flagged by the fast kernels:
//   RS_FLAG_NEARMAX  a cell whose power is within det_eps of its best neighbour or of the threshold
//                    (detections AND near-misses: the latter are emitted as RS_FLAG_DROPPED candidates)
//   RS_FLAG_TIE / RS_FLAG_GUARD  a grid argmax whose runner-up is within tie_eps, or a MUSIC denominator
//                    near the reference's 1e-12 guard (angle_estimation.py:149)
// These kernels settle exactly those few decisions with the reference's rule (dechirp.py:250-254,
// angle_estimation.py:143-152, 173):
//   * detections: fp64 direct DFT of the 3x3 neighbourhood from the raw cube (fp64 dechirp*window table,
//     fp64 twiddles), exact comparison.
//   * angles, stage A: fp64 grid scan of the fp32 snapshot.  The snapshot carries the fp32 FFT's rounding
//     error dS with |dS_m| <= sigma = fft_eps * rms(|X|) of the frame, which moves P_g - P_h by at most
//     2 |conj(b_g) a_g - conj(b_h) a_h| |dS|  (b_g = a_g^H s).  If the winner beats EVERY other grid point by
//     more than that bound the decision is final -- adjacent grid points, whose steering vectors are
//     nearly equal, have a tiny bound, so almost all fp32 ties are settled here.
//   * angles, stage B (the rest, and anything near the MUSIC guard): the snapshot itself is recomputed in
//     fp64 from the raw cube and scanned exactly.
// One CTA per detection segment walks its flagged items in list order, so results (and the corrections
// to the velocity sums) are deterministic.  Segments without flagged items exit on a per-segment counter.
#include "rs_common.cuh"

namespace {

constexpr int RC_THREADS = 256;
constexpr int RC_MAX_ITEMS = 512;

struct CubeView {
    const float2* cube;    // [F][A][C_total][S]
    const double2* tab;    // [S] conj(ref) * window, fp64
    int A, C_total, chirp0, C, S, dc;
};

__device__ __forceinline__ double2 dmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ void build_twiddles(double2* ws, int S, double2* wc, int C) {
    for (int m = threadIdx.x; m < S; m += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)m / (double)S, &sn, &cs);
        ws[m] = make_double2(cs, sn);
    }
    for (int m = threadIdx.x; m < C; m += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)m / (double)C, &sn, &cs);
        wc[m] = make_double2(cs, sn);
    }
    __syncthreads();
}

// X[a][rp_i][dp_j] (shifted bin positions, as stored in the RDS) for i < nr, j < nd, in fp64 from the raw cube.
// Whole CTA participates; out[i * 3 + j] in shared memory.  T: scratch double2 [3][C].
__device__ void dft_cells(const CubeView& v, int f, int a, const int* rp, int nr, const int* dp, int nd,
                          const double2* ws, const double2* wc, double2* T, double2* out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int kr[3], kd[3];
    for (int i = 0; i < 3; ++i) {
        kr[i] = i < nr ? (rp[i] - v.S / 2 + v.S) % v.S : 0;          // undo np.fft.fftshift
        kd[i] = i < nd ? (dp[i] - v.C / 2 + v.C) % v.C : 0;
    }
    for (int c = wid; c < v.C; c += nw) {
        const float2* x = v.cube + (((size_t)f * v.A + a) * v.C_total + v.chirp0 + c) * v.S;
        double2 acc[3] = {{0, 0}, {0, 0}, {0, 0}};
        int idx[3], step[3];                       // (s * kr_i) mod S, advanced incrementally
        for (int i = 0; i < 3; ++i) {
            idx[i] = (int)(((long long)lane * kr[i]) % v.S);
            step[i] = (int)((32ll * kr[i]) % v.S);
        }
        for (int s0 = lane; s0 < v.S; s0 += 256) {
            // eight loads in flight per lane before any of them is consumed (the item is latency bound)
            float2 xs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = s0 + 32 * j;
                xs[j] = s < v.S ? __ldg(x + s) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = s0 + 32 * j;
                if (s < v.S) {
                    const double2 y = dmul(make_double2((double)xs[j].x, (double)xs[j].y), v.tab[s]);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < nr) {
                            const double2 t = dmul(y, ws[idx[i]]);
                            acc[i].x += t.x;
                            acc[i].y += t.y;
                            idx[i] += step[i];
                            if (idx[i] >= v.S) idx[i] -= v.S;
                        }
                    }
                }
            }
        }
        for (int i = 0; i < nr; ++i) {
            const double re = warp_sum(acc[i].x), im = warp_sum(acc[i].y);
            if (lane == 0) T[i * v.C + c] = make_double2(re, im);
        }
    }
    __syncthreads();
    for (int o = wid; o < nr * nd; o += nw) {
        const int i = o / nd, j = o - i * nd;
        double ar = 0, ai = 0;
        for (int c = lane; c < v.C; c += 32) {
            const double2 t = dmul(T[i * v.C + c], wc[(int)(((long long)c * kd[j]) % v.C)]);
            ar += t.x;
            ai += t.y;
        }
        ar = warp_sum(ar);
        ai = warp_sum(ai);
        // the reference subtracts the per-chirp mean (dechirp.py:120): range bin 0 is ~1e-15, taken as 0
        if (lane == 0) out[i * 3 + j] = (v.dc && kr[i] == 0) ? make_double2(0, 0) : make_double2(ar, ai);
    }
    __syncthreads();
}

// The A-channel snapshot of ONE cell in fp64: z[s] = table[s] w_S^{s kr} and u[c] = w_C^{c kd} are formed once,
// then every antenna costs one complex multiply-add per raw sample.
__device__ void snapshot_f64(const CubeView& v, int f, int rpos, int dpos, const double2* ws, const double2* wc,
                             double2* z, double2* u, double2* T, double2* snap) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int kr = (rpos - v.S / 2 + v.S) % v.S, kd = (dpos - v.C / 2 + v.C) % v.C;
    for (int s = threadIdx.x; s < v.S; s += blockDim.x) z[s] = dmul(v.tab[s], ws[(int)(((long long)s * kr) % v.S)]);
    for (int c = threadIdx.x; c < v.C; c += blockDim.x) u[c] = wc[(int)(((long long)c * kd) % v.C)];
    __syncthreads();
    for (int a = 0; a < v.A; ++a) {
        for (int c = wid; c < v.C; c += nw) {
            const float2* x = v.cube + (((size_t)f * v.A + a) * v.C_total + v.chirp0 + c) * v.S;
            // four independent accumulator pairs: the fp64 pipe is latency bound on a single chain
            double ar0 = 0, ai0 = 0, ar1 = 0, ai1 = 0, ar2 = 0, ai2 = 0, ar3 = 0, ai3 = 0;
            int s = lane;
            for (; s + 96 < v.S; s += 128) {
                const float2 x0 = x[s], x1 = x[s + 32], x2 = x[s + 64], x3 = x[s + 96];
                const double2 w0 = z[s], w1 = z[s + 32], w2 = z[s + 64], w3 = z[s + 96];
                ar0 += (double)x0.x * w0.x - (double)x0.y * w0.y; ai0 += (double)x0.x * w0.y + (double)x0.y * w0.x;
                ar1 += (double)x1.x * w1.x - (double)x1.y * w1.y; ai1 += (double)x1.x * w1.y + (double)x1.y * w1.x;
                ar2 += (double)x2.x * w2.x - (double)x2.y * w2.y; ai2 += (double)x2.x * w2.y + (double)x2.y * w2.x;
                ar3 += (double)x3.x * w3.x - (double)x3.y * w3.y; ai3 += (double)x3.x * w3.y + (double)x3.y * w3.x;
            }
            for (; s < v.S; s += 32) {
                const float2 xv = x[s];
                const double2 w = z[s];
                ar0 += (double)xv.x * w.x - (double)xv.y * w.y;
                ai0 += (double)xv.x * w.y + (double)xv.y * w.x;
            }
            double ar = warp_sum((ar0 + ar1) + (ar2 + ar3));
            double ai = warp_sum((ai0 + ai1) + (ai2 + ai3));
            if (lane == 0) T[c] = dmul(make_double2(ar, ai), u[c]);
        }
        __syncthreads();
        if (wid == 0) {
            double ar = 0, ai = 0;
            for (int c = lane; c < v.C; c += 32) { ar += T[c].x; ai += T[c].y; }
            ar = warp_sum(ar);
            ai = warp_sum(ai);
            if (lane == 0) snap[a] = (v.dc && kr == 0) ? make_double2(0, 0) : make_double2(ar, ai);
        }
        __syncthreads();
    }
}

// ordered list of the flagged items of a segment: item i qualifies when flags[pos(i)] & mask and not & done
template <typename PosFn>
__device__ int collect_items(int n, PosFn pos_of, const uint8_t* flags, uint8_t mask, uint8_t done, int* items,
                             int* scratch) {
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(n, lo + per);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) {
        const uint8_t fl = flags[pos_of(i)];
        cnt += ((fl & mask) && !(fl & done)) ? 1 : 0;
    }
    scratch[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const int c = scratch[t];
            scratch[t] = run;
            run += c;
        }
        scratch[blockDim.x] = run;
    }
    __syncthreads();
    int at = scratch[threadIdx.x];
    const int total = scratch[blockDim.x];
    for (int i = lo; i < hi; ++i) {
        const uint8_t fl = flags[pos_of(i)];
        if ((fl & mask) && !(fl & done)) {
            if (at < RC_MAX_ITEMS) items[at] = i;
            ++at;
        }
    }
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------
// detections: exact local-maximum / threshold decision for NEARMAX entries (hits and candidates)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RC_THREADS)
recheck_detect_kernel(CubeView v, double thr64, const uint32_t* __restrict__ det_key, uint8_t* __restrict__ det_flags,
                      const int32_t* __restrict__ det_count, const int32_t* __restrict__ det_nnear, int seg_cap,
                      int nseg, int R, int D, const int32_t* __restrict__ det_aidx, const float* __restrict__ det_phase,
                      const double* __restrict__ grid_cs, double* __restrict__ ls_partials,
                      int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];
    double2* ws = smd;
    double2* wc = ws + v.S;
    double2* T = wc + v.C;
    __shared__ double2 out[9];
    __shared__ int items[RC_MAX_ITEMS];
    __shared__ int scratch[RC_THREADS + 1];
    __shared__ int rp[3], dp[3];
    const int seg = blockIdx.x;
    if (det_nnear != nullptr && det_nnear[seg] == 0) return;
    const int n = det_count[seg];
    if (n == 0) return;
    const size_t base = (size_t)seg * seg_cap;
    const int total = collect_items(n, [&](int i) { return base + i; }, det_flags, RS_FLAG_NEARMAX, RS_FLAG_DETFIXED, items,
                                    scratch);
    if (total == 0) return;
    build_twiddles(ws, v.S, wc, v.C);
    const int f = seg / nseg;
    const int todo = min(total, RC_MAX_ITEMS);
    for (int it = 0; it < todo; ++it) {
        const size_t o = base + items[it];
        int a, r, d;
        rs_split_key(det_key[o], a, r, d);
        const int r_lo = max(r - 1, 0), r_hi = min(r + 1, R - 1), d_lo = max(d - 1, 0), d_hi = min(d + 1, D - 1);
        const int nr = r_hi - r_lo + 1, nd = d_hi - d_lo + 1;
        if (threadIdx.x < 3) {
            rp[threadIdx.x] = r_lo + threadIdx.x;
            dp[threadIdx.x] = d_lo + threadIdx.x;
        }
        __syncthreads();
        dft_cells(v, f, a, rp, nr, dp, nd, ws, wc, T, out);
        if (threadIdx.x == 0) {
            double c = 0, m = -1;
            for (int i = 0; i < nr; ++i)
                for (int j = 0; j < nd; ++j) {
                    const double2 x = out[i * 3 + j];
                    const double p = x.x * x.x + x.y * x.y;
                    if (r_lo + i == r && d_lo + j == d) c = p;
                    else m = fmax(m, p);
                }
            // dechirp.py:250-254 on exact powers: equality with the 3x3 maximum and dB strictly above threshold
            const bool is_det = (c >= m) && (c + 1e-12 > thr64);
            uint8_t fl = det_flags[o];
            const bool was_det = !(fl & RS_FLAG_DROPPED);
            if (is_det != was_det) {
                atomicAdd(stats + (is_det ? 2 : 1), 1);
                // run after rs_angles: move this detection's row into / out of the velocity sums
                if (ls_partials != nullptr && det_aidx[o] >= 0) {
                    const double sg = is_det ? 1.0 : -1.0, y = (double)det_phase[o];
                    const double cc = grid_cs[2 * det_aidx[o]], ss = grid_cs[2 * det_aidx[o] + 1];
                    double* ps = ls_partials + (size_t)seg * 8;
                    ps[0] += sg * cc * cc; ps[1] += sg * ss * ss; ps[2] += sg * cc * ss;
                    ps[3] += sg * y * cc;  ps[4] += sg * y * ss;  ps[5] += sg * y * y;  ps[6] += sg;
                }
            }
            fl = is_det ? (fl & ~RS_FLAG_DROPPED) : (fl | RS_FLAG_DROPPED);
            det_flags[o] = fl | RS_FLAG_DETFIXED;
            atomicAdd(stats + 0, 1);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && total > RC_MAX_ITEMS) atomicAdd(stats + 3, total - RC_MAX_ITEMS);
}

// ---------------------------------------------------------------------------------------------
// angles: exact grid argmax for TIE / GUARD cells
// ---------------------------------------------------------------------------------------------
struct AngleFix {
    const float2* rds;
    const double2* steer;      // [A][G] exp(+i m phi_g), fp64
    const float* grid_deg;
    const double* grid_cs;
    int G, method;
    double fft_eps;            // bound on the fp32 FFT's absolute error per element, in units of rms(|X|)
    const float* det_psum;     // per segment sum of |X|^2 (rs_detect)
    const int32_t* det_ntie;
    const uint32_t* det_key;
    const uint32_t* det_lead;
    const int32_t* det_nlead;
    uint8_t* det_flags;
    int32_t* det_aidx;
    float* det_adeg;
    const float* det_phase;
    double* ls_partials;
    int seg_cap, nseg, R, D;
};

struct ScanRes {
    int idx;            // first-index argmax of the method's pseudo-spectrum
    bool undecided;     // some other grid point is within the error bound of the winner, or guard zone
};

// |a_g^H s| etc. for one grid point
__device__ __forceinline__ double2 beam(const AngleFix& q, const double2* snap, int A, int g) {
    double ar = 0, ai = 0;
    for (int m = 0; m < A; ++m) {
        const double2 w = q.steer[(size_t)m * q.G + g];
        const double2 x = snap[m];
        ar += w.x * x.x + w.y * x.y;      // conj(w) * x
        ai += w.x * x.y - w.y * x.x;
    }
    return make_double2(ar, ai);
}

// ds_norm: bound on |dS|_2 of the snapshot (0 when the snapshot is exact)
__device__ ScanRes scan_f64(const AngleFix& q, const double2* snap, int A, double energy, double ds_norm, double* red_v,
                            int* red_i) {
    double best = -1.0, bestp = -1.0;
    int bi = 0x7fffffff;
    for (int g = threadIdx.x; g < q.G; g += blockDim.x) {
        const double2 b = beam(q, snap, A, g);
        const double pwr = b.x * b.x + b.y * b.y;
        double val = pwr;
        if (q.method == RS_METHOD_MUSIC) {
            const double den = fabs((double)A - (energy > 0 ? pwr / energy : 1.0));
            val = den > 1e-12 ? 1.0 / den : 0.0;                   // angle_estimation.py:149-152
        }
        if (val > best) { best = val; bi = g; }
        bestp = fmax(bestp, pwr);
    }
    __shared__ double r_p[RC_THREADS];
    red_v[threadIdx.x] = best; red_i[threadIdx.x] = bi; r_p[threadIdx.x] = bestp;
    __syncthreads();
    for (int s = blockDim.x >> 1; s; s >>= 1) {
        if (threadIdx.x < s) {
            const double ov = red_v[threadIdx.x + s];
            const int oi = red_i[threadIdx.x + s];
            if (ov > red_v[threadIdx.x] || (ov == red_v[threadIdx.x] && oi < red_i[threadIdx.x])) {
                red_v[threadIdx.x] = ov;
                red_i[threadIdx.x] = oi;
            }
            r_p[threadIdx.x] = fmax(r_p[threadIdx.x], r_p[threadIdx.x + s]);
        }
        __syncthreads();
    }
    ScanRes res;
    res.idx = red_i[0];
    const double pmax = r_p[0];
    __syncthreads();
    bool und = false;
    if (ds_norm > 0) {
        // near the guard the pseudo-spectrum is not monotone in the beam power: needs the exact snapshot
        if (q.method == RS_METHOD_MUSIC && ((double)A * energy - pmax <= 1e-6 * (double)A * energy)) und = true;
        const int gs = res.idx;
        const double2 bg = beam(q, snap, A, gs);
        const double pg = bg.x * bg.x + bg.y * bg.y;
        for (int h = threadIdx.x; h < q.G && !und; h += blockDim.x) {
            if (h == gs) continue;
            const double2 bh = beam(q, snap, A, h);
            const double ph = bh.x * bh.x + bh.y * bh.y;
            // |conj(b_g) a_g - conj(b_h) a_h|^2 = A (|b_g|^2 + |b_h|^2) - 2 Re(b_g conj(b_h) (a_g^H a_h))
            double cr = 0, ci = 0;
            for (int m = 0; m < A; ++m) {
                const double2 wg = q.steer[(size_t)m * q.G + gs], wh = q.steer[(size_t)m * q.G + h];
                cr += wg.x * wh.x + wg.y * wh.y;                   // conj(a_g) a_h
                ci += wg.x * wh.y - wg.y * wh.x;
            }
            const double tr = bg.x * bh.x + bg.y * bh.y, ti = bg.y * bh.x - bg.x * bh.y;      // b_g conj(b_h)
            const double n2 = fmax(0.0, (double)A * (pg + ph) - 2.0 * (tr * cr - ti * ci));
            const double bound = 2.0 * sqrt(n2) * ds_norm;
            if (pg - ph <= bound) und = true;
        }
    }
    res.undecided = __syncthreads_or(und ? 1 : 0) != 0;
    return res;
}

// writes one settled cell to all of its detections; dl = correction of the segment's velocity sums
__device__ void apply_angle(const AngleFix& q, size_t o, int k, int idx, double* dl, int32_t* stats, int stage) {
    const int old = q.det_aidx[o];
    int live = 0;
    for (int e = 0; e < k; ++e) {
        const uint8_t fl = q.det_flags[o + e];
        live += (fl & RS_FLAG_DROPPED) ? 0 : 1;
        q.det_flags[o + e] = fl | RS_FLAG_FIXED;
        q.det_aidx[o + e] = idx;
        q.det_adeg[o + e] = q.grid_deg[idx];
    }
    if (old != idx && live > 0) {
        const double w = (double)live, y = (double)q.det_phase[o];
        const double c0 = q.grid_cs[2 * old], s0 = q.grid_cs[2 * old + 1];
        const double c1 = q.grid_cs[2 * idx], s1 = q.grid_cs[2 * idx + 1];
        dl[0] = w * (c1 * c1 - c0 * c0); dl[1] = w * (s1 * s1 - s0 * s0); dl[2] = w * (c1 * s1 - c0 * s0);
        dl[3] = w * y * (c1 - c0);       dl[4] = w * y * (s1 - s0);
    }
    atomicAdd(stats + 0, 1);
    if (old != idx) atomicAdd(stats + 1, 1);
    if (stage) atomicAdd(stats + 2, 1);
}

// warp-level variant of scan_f64 for stage A: lanes over grid points, snapshot in shared memory.
// Outside the MUSIC guard zone 1/(M - P/E) is monotone in the beam power P, so the argmax is taken on P; the
// guard zone and every pair that the error bound cannot separate go to stage B.
constexpr int SW_MAX_PER_LANE = 12;            // grids up to 384 points keep (b_h, P_h) in registers
__device__ ScanRes scan_warp(const AngleFix& q, const double2* snap, int A, double energy, double ds_norm) {
    const int lane = threadIdx.x & 31;
    double2 bl[SW_MAX_PER_LANE];
    double bestp = -1.0;
    int bi = 0x7fffffff;
    const bool cached = q.G <= 32 * SW_MAX_PER_LANE;
#pragma unroll
    for (int j = 0; j < SW_MAX_PER_LANE; ++j) {
        const int g = lane + 32 * j;
        if (g < q.G) {
            const double2 b = beam(q, snap, A, g);
            bl[j] = b;
            const double pwr = b.x * b.x + b.y * b.y;
            if (pwr > bestp) { bestp = pwr; bi = g; }
        }
    }
    for (int g = lane + 32 * SW_MAX_PER_LANE; g < q.G; g += 32) {
        const double2 b = beam(q, snap, A, g);
        const double pwr = b.x * b.x + b.y * b.y;
        if (pwr > bestp) { bestp = pwr; bi = g; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const double op = __shfl_xor_sync(0xffffffffu, bestp, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (op > bestp || (op == bestp && oi < bi)) { bestp = op; bi = oi; }
    }
    ScanRes res;
    res.idx = bi;
    bool und = (q.method == RS_METHOD_MUSIC) && ((double)A * energy - bestp <= 1e-6 * (double)A * energy);
    const double2 bg = beam(q, snap, A, bi);
    const double pg = bestp;
    const double wg = sqrt((double)A * pg);
    auto test = [&](int h, double2 bh) {
        if (h == bi || und) return;
        const double ph = bh.x * bh.x + bh.y * bh.y;
        const double gap = pg - ph;
        // cheap sufficient test: |conj(b_g) a_g - conj(b_h) a_h| <= sqrt(A) (|b_g| + |b_h|)
        if (gap > 2.0 * (wg + sqrt((double)A * ph)) * ds_norm) return;
        double cr = 0, ci = 0;
        for (int m = 0; m < A; ++m) {
            const double2 wgm = q.steer[(size_t)m * q.G + bi], wh = q.steer[(size_t)m * q.G + h];
            cr += wgm.x * wh.x + wgm.y * wh.y;
            ci += wgm.x * wh.y - wgm.y * wh.x;
        }
        const double tr = bg.x * bh.x + bg.y * bh.y, ti = bg.y * bh.x - bg.x * bh.y;
        const double n2 = fmax(0.0, (double)A * (pg + ph) - 2.0 * (tr * cr - ti * ci));
        if (gap <= 2.0 * sqrt(n2) * ds_norm) und = true;
    };
    if (cached) {
#pragma unroll
        for (int j = 0; j < SW_MAX_PER_LANE; ++j) {
            const int h = lane + 32 * j;
            if (h < q.G) test(h, bl[j]);
        }
    } else {
        for (int h = lane; h < q.G; h += 32) test(h, beam(q, snap, A, h));
    }
    res.undecided = __any_sync(0xffffffffu, und);
    return res;
}

constexpr int RA_MAX_ITEMS = 256;
constexpr int RB_MAX = 16;          // undecided cells per segment handed to stage B
constexpr int RB_BATCH = 4;         // cells whose range twiddles are staged together in stage B

// -------- stage A: one CTA per segment, one warp per flagged cell; undecided cells go to the stage-B list
__global__ void __launch_bounds__(RC_THREADS)
recheck_angles_kernel(CubeView v, AngleFix q, int32_t* __restrict__ work_idx, int32_t* __restrict__ work_cnt,
                      int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];
    double2* snap = smd;                 // [warps][A]
    __shared__ int items[RC_MAX_ITEMS];
    __shared__ int scratch[RC_THREADS + 1];
    __shared__ double delta[RA_MAX_ITEMS][5];
    __shared__ unsigned char need_b[RA_MAX_ITEMS];
    const int seg = blockIdx.x;
    if (threadIdx.x == 0) work_cnt[seg] = 0;
    if (q.det_ntie != nullptr && q.det_ntie[seg] == 0) return;
    const int n = q.det_nlead[seg];
    if (n == 0) return;
    const size_t base = (size_t)seg * q.seg_cap;
    const int total = collect_items(n, [&](int i) { return base + (q.det_lead[base + i] & 0xFFFFu); }, q.det_flags,
                                    RS_FLAG_TIE | RS_FLAG_GUARD, RS_FLAG_FIXED, items, scratch);
    if (total == 0) return;
    const int f = seg / q.nseg;
    const int A = v.A;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // rms(|X|) of the frame from the per-tile power sums; |dS|_2 <= fft_eps * rms * sqrt(A)
    double psum = 0;
    for (int sg = 0; sg < q.nseg; ++sg) psum += (double)q.det_psum[(size_t)f * q.nseg + sg];
    const double ds_norm = fmax(q.fft_eps * sqrt(psum / ((double)q.R * q.D * A)) * sqrt((double)A), 1e-300);
    const int todo = min(total, RA_MAX_ITEMS);
    for (int it = wid; it < todo; it += nw) {
        const uint32_t ld = q.det_lead[base + items[it]];
        const size_t o = base + (ld & 0xFFFFu);
        const int k = (int)(ld >> 16);
        int a, r, d;
        rs_split_key(q.det_key[o], a, r, d);
        const float2* cell = q.rds + (((size_t)f * q.R + r) * q.D + d) * A;
        double2* sw = snap + wid * A;
        __syncwarp();
        for (int m = lane; m < A; m += 32) sw[m] = make_double2((double)cell[m].x, (double)cell[m].y);
        __syncwarp();
        double energy = 0;
        for (int m = 0; m < A; ++m) energy += sw[m].x * sw[m].x + sw[m].y * sw[m].y;
        // |dS|_2 <= fft_eps * rms * sqrt(A)  (noise-like part)  +  1.5e-7 |s|_2  (part that scales with the cell itself)
        const ScanRes sr = scan_warp(q, sw, A, energy, ds_norm + 1.5e-7 * sqrt(energy));
        if (lane == 0) {
            need_b[it] = sr.undecided ? 1 : 0;
            double dl[5] = {0, 0, 0, 0, 0};
            if (!sr.undecided) apply_angle(q, o, k, sr.idx, dl, stats, 0);
            for (int j = 0; j < 5; ++j) delta[it][j] = dl[j];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nb = 0, lost = total > RA_MAX_ITEMS ? total - RA_MAX_ITEMS : 0;
        double* ps = q.ls_partials ? q.ls_partials + (size_t)seg * 8 : nullptr;
        for (int it = 0; it < todo; ++it) {
            if (ps) for (int j = 0; j < 5; ++j) ps[j] += delta[it][j];       // item order: deterministic
            if (need_b[it]) {
                if (nb < RB_MAX) work_idx[(size_t)seg * RB_MAX + nb++] = items[it];
                else ++lost;
            }
        }
        work_cnt[seg] = nb;
        if (lost) atomicAdd(stats + 3, lost);
    }
}

// inner loop of stage B1 for NB cells at once: every raw sample is converted to fp64 once and multiplied into
// NB accumulators; the Doppler twiddle index advances incrementally (no integer division in the loop)
template <int NB>
__device__ __forceinline__ void b1_accumulate(const float2* __restrict__ plane, int S, int C, const double2* z,
                                              const double2* wc, const int* kd, double2* acc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int ci[NB], cstep[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        ci[i] = (int)(((long long)wid * kd[i]) % C);
        cstep[i] = (int)(((long long)nw * kd[i]) % C);
    }
    for (int c = wid; c < C; c += nw) {
        const float2* x = plane + (size_t)c * S;
        double tr[NB], ti[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) { tr[i] = 0; ti[i] = 0; }
#pragma unroll 4
        for (int s = lane; s < S; s += 32) {
            const float2 xv = __ldg(x + s);
            const double xr = (double)xv.x, xi = (double)xv.y;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const double2 w = z[i * S + s];
                tr[i] = fma(xr, w.x, fma(-xi, w.y, tr[i]));
                ti[i] = fma(xr, w.y, fma(xi, w.x, ti[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double2 u = wc[ci[i]];
            acc[i].x += tr[i] * u.x - ti[i] * u.y;       // per-lane partial times the Doppler twiddle
            acc[i].y += tr[i] * u.y + ti[i] * u.x;
            ci[i] += cstep[i];
            if (ci[i] >= C) ci[i] -= C;
        }
    }
}

// -------- stage B1: one CTA per (frame, antenna) reads that antenna's raw rows ONCE and forms its element of the
// fp64 snapshot of every undecided cell of the frame
__global__ void __launch_bounds__(RC_THREADS)
recheck_snapshots_kernel(CubeView v, AngleFix q, const int32_t* __restrict__ work_idx,
                         const int32_t* __restrict__ work_cnt, double2* __restrict__ work_snap) {
    extern __shared__ double2 smd[];
    double2* ws = smd;                          // [S]
    double2* wc = ws + v.S;                     // [C]
    double2* z = wc + v.C;                      // [RB_BATCH][S]
    double2* red = z + RB_BATCH * v.S;          // [warps][RB_BATCH]
    __shared__ int cell_r[RB_MAX * 64], cell_d[RB_MAX * 64], cell_slot[RB_MAX * 64];
    __shared__ int n_cells;
    const int a = blockIdx.x, f = blockIdx.y;
    if (threadIdx.x == 0) {
        int nc = 0;
        for (int sg = 0; sg < q.nseg; ++sg) {
            const int seg = f * q.nseg + sg;
            const int cnt = work_cnt[seg];
            for (int k = 0; k < cnt && nc < RB_MAX * 64; ++k) {       // same cap in recheck_finish_kernel
                const size_t base = (size_t)seg * q.seg_cap;
                const uint32_t ld = q.det_lead[base + work_idx[(size_t)seg * RB_MAX + k]];
                int aa, r, d;
                rs_split_key(q.det_key[base + (ld & 0xFFFFu)], aa, r, d);
                cell_r[nc] = r; cell_d[nc] = d; cell_slot[nc] = seg * RB_MAX + k;
                ++nc;
            }
        }
        n_cells = nc;
    }
    __syncthreads();
    const int nc = n_cells;
    if (nc == 0) return;
    build_twiddles(ws, v.S, wc, v.C);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b0 = 0; b0 < nc; b0 += RB_BATCH) {
        const int nb = min(RB_BATCH, nc - b0);
        int kd[RB_BATCH];
        for (int i = 0; i < RB_BATCH; ++i) kd[i] = i < nb ? (cell_d[b0 + i] - v.C / 2 + v.C) % v.C : 0;
        for (int i = 0; i < nb; ++i) {
            const int kr = (cell_r[b0 + i] - v.S / 2 + v.S) % v.S;
            for (int s = threadIdx.x; s < v.S; s += blockDim.x)
                z[i * v.S + s] = dmul(v.tab[s], ws[(int)(((long long)s * kr) % v.S)]);
        }
        __syncthreads();
        double2 acc[RB_BATCH];
        for (int i = 0; i < RB_BATCH; ++i) acc[i] = make_double2(0, 0);
        const float2* plane = v.cube + (((size_t)f * v.A + a) * v.C_total + v.chirp0) * v.S;
        switch (nb) {
            case 1: b1_accumulate<1>(plane, v.S, v.C, z, wc, kd, acc); break;
            case 2: b1_accumulate<2>(plane, v.S, v.C, z, wc, kd, acc); break;
            case 3: b1_accumulate<3>(plane, v.S, v.C, z, wc, kd, acc); break;
            default: b1_accumulate<4>(plane, v.S, v.C, z, wc, kd, acc); break;
        }
        for (int i = 0; i < nb; ++i) {
            const double re = warp_sum(acc[i].x), im = warp_sum(acc[i].y);
            if (lane == 0) red[wid * RB_BATCH + i] = make_double2(re, im);
        }
        __syncthreads();
        if (threadIdx.x < nb) {
            const int i = threadIdx.x;
            double re = 0, im = 0;
            for (int w = 0; w < nw; ++w) { re += red[w * RB_BATCH + i].x; im += red[w * RB_BATCH + i].y; }
            const int kr = (cell_r[b0 + i] - v.S / 2 + v.S) % v.S;
            work_snap[(size_t)cell_slot[b0 + i] * v.A + a] = (v.dc && kr == 0) ? make_double2(0, 0) : make_double2(re, im);
        }
        __syncthreads();
    }
}

// -------- stage B2: one CTA per segment scans the exact snapshots and settles the remaining cells
__global__ void __launch_bounds__(RC_THREADS)
recheck_finish_kernel(CubeView v, AngleFix q, const int32_t* __restrict__ work_idx, const int32_t* __restrict__ work_cnt,
                      const double2* __restrict__ work_snap, int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];
    double2* snap = smd;                 // [A]
    __shared__ double red_v[RC_THREADS];
    __shared__ int red_i[RC_THREADS];
    const int seg = blockIdx.x;
    const int cnt = work_cnt[seg];
    if (cnt == 0) return;
    const int A = v.A;
    const size_t base = (size_t)seg * q.seg_cap;
    // cells of earlier segments of this frame (stage B1 handles at most RB_MAX * 64 cells per frame)
    int before = 0;
    for (int sg = (seg / q.nseg) * q.nseg; sg < seg; ++sg) before += work_cnt[sg];
    for (int k = 0; k < cnt; ++k) {
        if (before + k >= RB_MAX * 64) {
            if (threadIdx.x == 0) atomicAdd(stats + 3, 1);
            continue;
        }
        const uint32_t ld = q.det_lead[base + work_idx[(size_t)seg * RB_MAX + k]];
        const size_t o = base + (ld & 0xFFFFu);
        const int mult = (int)(ld >> 16);
        for (int m = threadIdx.x; m < A; m += blockDim.x) snap[m] = work_snap[((size_t)seg * RB_MAX + k) * A + m];
        __syncthreads();
        double energy = 0;
        for (int m = 0; m < A; ++m) energy += snap[m].x * snap[m].x + snap[m].y * snap[m].y;
        const ScanRes sr = scan_f64(q, snap, A, energy, 0.0, red_v, red_i);
        if (threadIdx.x == 0) {
            double dl[5] = {0, 0, 0, 0, 0};
            apply_angle(q, o, mult, sr.idx, dl, stats, 1);
            if (q.ls_partials) {
                double* ps = q.ls_partials + (size_t)seg * 8;
                for (int j = 0; j < 5; ++j) ps[j] += dl[j];
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int rs_recheck_detections_f64(const void* cube, const void* table128, int C_total, int chirp0, int dc_removal,
                                         double thr_power64, const uint32_t* det_key, uint8_t* det_flags,
                                         const int32_t* det_count, const int32_t* det_nnear, int seg_cap,
                                         int nseg_per_frame, int F, int A, int C, int S, const int32_t* det_aidx,
                                         const float* det_phase, const double* grid_cs, double* ls_partials,
                                         int32_t* stats, void* stream) {
    RS_CHECK_ARG(cube && table128 && det_key && det_flags && det_count && stats, "rs_recheck_detections_f64: null pointer");
    RS_CHECK_ARG(ls_partials == nullptr || (det_aidx && det_phase && grid_cs),
                 "rs_recheck_detections_f64: ls_partials needs det_aidx, det_phase and grid_cs");
    RS_CHECK_ARG(F > 0 && A > 0 && C > 0 && S > 0 && chirp0 >= 0 && chirp0 + C <= C_total && seg_cap > 0 &&
                     nseg_per_frame > 0,
                 "rs_recheck_detections_f64: bad dims");
    CubeView v{(const float2*)cube, (const double2*)table128, A, C_total, chirp0, C, S, dc_removal};
    const size_t smem = (size_t)(S + 4 * C) * sizeof(double2);
    if (smem > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_recheck_detections_f64: needs %zu B of shared memory", smem);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(recheck_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), (cudaStream_t)stream);
    const long long blocks = (long long)F * nseg_per_frame;
    recheck_detect_kernel<<<(unsigned)blocks, RC_THREADS, smem, (cudaStream_t)stream>>>(
        v, thr_power64, det_key, det_flags, det_count, det_nnear, seg_cap, nseg_per_frame, S, C, det_aidx, det_phase,
        grid_cs, ls_partials, stats);
    RS_CHECK_LAUNCH("rs_recheck_detections_f64");
    return RS_OK;
}

extern "C" int rs_recheck_angles_f64(const void* cube, const void* table128, int C_total, int chirp0, int dc_removal,
                                     const void* rds, const void* steer128, const float* grid_deg, const double* grid_cs,
                                     int G, int method, double fft_eps, const float* det_psum, const int32_t* det_ntie,
                                     const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead,
                                     uint8_t* det_flags, int32_t* det_aidx, float* det_adeg, const float* det_phase,
                                     double* ls_partials, int seg_cap, int nseg_per_frame, int F, int A, int C, int S,
                                     int32_t* work_idx, int32_t* work_cnt, void* work_snap, int32_t* stats,
                                     void* stream) {
    RS_CHECK_ARG(cube && table128 && rds && steer128 && grid_deg && grid_cs && det_psum && det_key && det_lead &&
                     det_nlead && det_flags && det_aidx && det_adeg && det_phase && work_idx && work_cnt && work_snap &&
                     stats,
                 "rs_recheck_angles_f64: null pointer");
    RS_CHECK_ARG(method == RS_METHOD_MUSIC || method == RS_METHOD_BEAMFORMING, "rs_recheck_angles_f64: grid methods only");
    RS_CHECK_ARG(F > 0 && F <= 65535 && A >= 2 && C > 0 && S > 0 && G > 0 && chirp0 >= 0 && chirp0 + C <= C_total &&
                     fft_eps >= 0 && nseg_per_frame > 0,
                 "rs_recheck_angles_f64: bad dims");
    CubeView v{(const float2*)cube, (const double2*)table128, A, C_total, chirp0, C, S, dc_removal};
    AngleFix q{(const float2*)rds, (const double2*)steer128, grid_deg, grid_cs, G, method, fft_eps, det_psum, det_ntie,
               det_key, det_lead, det_nlead, det_flags, det_aidx, det_adeg, det_phase, ls_partials, seg_cap,
               nseg_per_frame, S, C};
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_a = (size_t)(RC_THREADS / 32) * A * sizeof(double2);
    const size_t smem_b1 = (size_t)(S + C + RB_BATCH * S + (RC_THREADS / 32) * RB_BATCH) * sizeof(double2);
    const size_t smem_b2 = (size_t)A * sizeof(double2);
    if (smem_b1 > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_recheck_angles_f64: needs %zu B of shared memory", smem_b1);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(recheck_angles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    cudaFuncSetAttribute(recheck_snapshots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b1);
    cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), st);
    const long long blocks = (long long)F * nseg_per_frame;
    recheck_angles_kernel<<<(unsigned)blocks, RC_THREADS, smem_a, st>>>(v, q, work_idx, work_cnt, stats);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage A)");
    recheck_snapshots_kernel<<<dim3((unsigned)A, (unsigned)F), RC_THREADS, smem_b1, st>>>(v, q, work_idx, work_cnt,
                                                                                        (double2*)work_snap);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage B1)");
    recheck_finish_kernel<<<(unsigned)blocks, RC_THREADS, smem_b2, st>>>(v, q, work_idx, work_cnt,
                                                                        (const double2*)work_snap, stats);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage B2)");
    return RS_OK;
}
