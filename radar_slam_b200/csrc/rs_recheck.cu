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
// Work decomposition.  A segment's flagged items are always settled in list order by ONE owner, so the results and
// the corrections to the velocity sums are deterministic:
//   * detections: one CTA per 8 segments; segments without NEARMAX entries are skipped on a per-segment counter.
//   * angles, stage A and the final exact scan: one WARP per segment (lanes over grid points).
//   * angles, stage B snapshots: stage A appends the undecided cells of a frame to one per-frame list; one CTA per
//     (frame, antenna) then reads that antenna's raw plane ONCE for up to 8 cells at a time.
// The direct DFT is organised by fast-time column: a thread owns sample index s, walks the chirps with eight
// coalesced loads in flight, applies the Doppler twiddle of every requested cell (broadcast from shared memory)
// and only at the end the range factor table[s] w_S^{s k_r}; a block reduction over s finishes the cell.
#include "rs_common.cuh"

namespace {

constexpr int RC_THREADS = 256;
constexpr int RC_WARPS = RC_THREADS / 32;
constexpr int RC_MAX_ITEMS = 512;
constexpr int RB_MAX = 16;                       // undecided cells per segment handed to stage B
constexpr int FB_CAP = RS_RECHECK_FRAME_CAP;     // undecided cells per frame handed to stage B
constexpr int NB_MAX = 8;                        // cells accumulated per pass over a raw plane

struct CubeView {
    const float2* cube;    // [F][A][C_total][S]
    const double2* tab;    // [S] conj(ref) * window, fp64
    const double2* tw_s;   // [S] exp(-2 pi i k / S), fp64
    const double2* tw_c;   // [C] exp(-2 pi i k / C), fp64
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
__device__ __forceinline__ int unshift(int pos, int n) { return (pos - n / 2 + n) % n; }      // undo np.fft.fftshift

// Y[i] = sum_c x[c][s] u[i][c]  for one fast-time column (col = plane + s); LD loads in flight
template <int NB>
__device__ __forceinline__ void column_dft(const float2* __restrict__ col, int S, int C, const double2* __restrict__ u,
                                           double2 (&Y)[NB]) {
    constexpr int LD = NB <= 2 ? 16 : 8;          // few cells: the column walk is latency bound, keep more loads in flight
#pragma unroll
    for (int i = 0; i < NB; ++i) Y[i] = make_double2(0, 0);
    int c = 0;
    for (; c + LD <= C; c += LD) {
        float2 x[LD];
#pragma unroll
        for (int j = 0; j < LD; ++j) x[j] = __ldg(col + (size_t)(c + j) * S);
#pragma unroll
        for (int j = 0; j < LD; ++j) {
            const double xr = (double)x[j].x, xi = (double)x[j].y;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const double2 w = u[i * C + c + j];
                Y[i].x = fma(xr, w.x, fma(-xi, w.y, Y[i].x));
                Y[i].y = fma(xr, w.y, fma(xi, w.x, Y[i].y));
            }
        }
    }
    for (; c < C; ++c) {
        const float2 xv = __ldg(col + (size_t)c * S);
        const double xr = (double)xv.x, xi = (double)xv.y;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double2 w = u[i * C + c];
            Y[i].x = fma(xr, w.x, fma(-xi, w.y, Y[i].x));
            Y[i].y = fma(xr, w.y, fma(xi, w.x, Y[i].y));
        }
    }
}

// block sums of n complex values held per thread; result of value i lands in out[i] (thread 0 .. n-1 write)
template <int N>
__device__ __forceinline__ void block_reduce(const double2 (&acc)[N], int n, double2* red /* [RC_WARPS][N] */,
                                             double2* out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < n) {
            const double re = warp_sum(acc[i].x), im = warp_sum(acc[i].y);
            if (lane == 0) red[wid * N + i] = make_double2(re, im);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
        double re = 0, im = 0;
        for (int w = 0; w < RC_WARPS; ++w) { re += red[w * N + threadIdx.x].x; im += red[w * N + threadIdx.x].y; }
        out[threadIdx.x] = make_double2(re, im);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// detections: exact local-maximum / threshold decision for NEARMAX entries (hits and candidates)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RC_THREADS)
recheck_detect_kernel(CubeView v, double thr64, const uint32_t* __restrict__ det_key, uint8_t* __restrict__ det_flags,
                      const int32_t* __restrict__ det_count, const int32_t* __restrict__ det_nnear, int seg_cap,
                      int nseg, int nseg_total, int R, int D, const int32_t* __restrict__ det_aidx,
                      const float* __restrict__ det_phase, const double* __restrict__ grid_cs,
                      double* __restrict__ ls_partials, int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];
    double2* u = smd;                    // [3][C] Doppler twiddles of the three columns
    double2* red = u + 3 * v.C;          // [RC_WARPS][9]
    __shared__ double2 out[9];
    __shared__ int items[RC_MAX_ITEMS];
    __shared__ int wcnt[RC_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // which of the CTA's segments have anything to settle: all counters are fetched at once (one memory latency instead of
    // one per segment -- almost every segment is clean, and a clean CTA used to walk eight dependent loads)
    __shared__ unsigned todo_s;
    if (threadIdx.x < 32) {
        const int seg = blockIdx.x * RC_WARPS + lane;
        bool work = false;
        if (lane < RC_WARPS && seg < nseg_total) work = (det_nnear == nullptr || det_nnear[seg] != 0) && det_count[seg] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, work);
        if (lane == 0) todo_s = m;
    }
    __syncthreads();
    const unsigned todo_mask = todo_s;
    for (int sg = 0; sg < RC_WARPS; ++sg) {
        if (!((todo_mask >> sg) & 1u)) continue;
        const int seg = blockIdx.x * RC_WARPS + sg;
        const int n = det_count[seg];
        const size_t base = (size_t)seg * seg_cap;
        // ordered list of the segment's undecided entries
        int total = 0;
        for (int c0 = 0; c0 < n; c0 += RC_THREADS) {
            const int i = c0 + threadIdx.x;
            bool fl = false;
            if (i < n) {
                const uint8_t f8 = det_flags[base + i];
                fl = (f8 & RS_FLAG_NEARMAX) && !(f8 & RS_FLAG_DETFIXED);
            }
            const unsigned m = __ballot_sync(0xffffffffu, fl);
            if (lane == 0) wcnt[wid] = __popc(m);
            __syncthreads();
            int off = total, all = 0;
            for (int w = 0; w < RC_WARPS; ++w) {
                if (w < wid) off += wcnt[w];
                all += wcnt[w];
            }
            if (fl) {
                const int at = off + __popc(m & ((1u << lane) - 1u));
                if (at < RC_MAX_ITEMS) items[at] = i;
            }
            total += all;
            __syncthreads();
        }
        if (total == 0) continue;
        const int f = seg / nseg;
        const int todo = min(total, RC_MAX_ITEMS);
        for (int it = 0; it < todo; ++it) {
            const size_t o = base + items[it];
            int a, r, d;
            rs_split_key(det_key[o], a, r, d);
            const int r_lo = max(r - 1, 0), r_hi = min(r + 1, R - 1), d_lo = max(d - 1, 0), d_hi = min(d + 1, D - 1);
            const int nr = r_hi - r_lo + 1, nd = d_hi - d_lo + 1;
            for (int idx = threadIdx.x; idx < 3 * v.C; idx += RC_THREADS) {
                const int j = idx / v.C, c = idx - j * v.C;
                const int kd = unshift(min(d_lo + j, D - 1), v.C);
                u[idx] = j < nd ? v.tw_c[(int)(((long long)c * kd) % v.C)] : make_double2(0, 0);
            }
            __syncthreads();
            int kr[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) kr[i] = unshift(min(r_lo + i, R - 1), v.S);
            const float2* plane = v.cube + (((size_t)f * v.A + a) * v.C_total + v.chirp0) * v.S;
            double2 acc[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = make_double2(0, 0);
            for (int s = threadIdx.x; s < v.S; s += RC_THREADS) {
                double2 Y[3];
                column_dft<3>(plane + s, v.S, v.C, u, Y);
                const double2 t = v.tab[s];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double2 z = dmul(t, v.tw_s[(int)(((long long)s * kr[i]) % v.S)]);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double2 p = dmul(Y[j], z);
                        acc[i * 3 + j].x += p.x;
                        acc[i * 3 + j].y += p.y;
                    }
                }
            }
            block_reduce<9>(acc, 9, red, out);
            if (threadIdx.x == 0) {
                double c = 0, m = -1;
                for (int i = 0; i < nr; ++i)
                    for (int j = 0; j < nd; ++j) {
                        // the reference subtracts the per-chirp mean (dechirp.py:120): range bin 0 is ~1e-15, taken as 0
                        const double2 x = (v.dc && kr[i] == 0) ? make_double2(0, 0) : out[i * 3 + j];
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
                        const double sg2 = is_det ? 1.0 : -1.0, y = (double)det_phase[o];
                        const double cc = grid_cs[2 * det_aidx[o]], ss = grid_cs[2 * det_aidx[o] + 1];
                        double* ps = ls_partials + (size_t)seg * 8;
                        ps[0] += sg2 * cc * cc; ps[1] += sg2 * ss * ss; ps[2] += sg2 * cc * ss;
                        ps[3] += sg2 * y * cc;  ps[4] += sg2 * y * ss;  ps[5] += sg2 * y * y;  ps[6] += sg2;
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
    const int32_t* det_tielist;   // [nseg_total][RS_TIE_LIST_CAP] leader indices flagged by rs_angles (unordered) or null
    const uint32_t* det_key;
    const uint32_t* det_lead;
    const int32_t* det_nlead;
    uint8_t* det_flags;
    int32_t* det_aidx;
    float* det_adeg;
    const float* det_phase;
    double* ls_partials;
    int seg_cap, nseg, nseg_total, R, D;
};

struct ScanRes {
    int idx;            // first-index argmax of the method's pseudo-spectrum
    bool undecided;     // some other grid point is within the error bound of the winner, or guard zone
};

// a_g^H s for one grid point
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

// writes one settled cell to all of its detections; dl += correction of the segment's velocity sums
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
        dl[0] += w * (c1 * c1 - c0 * c0); dl[1] += w * (s1 * s1 - s0 * s0); dl[2] += w * (c1 * s1 - c0 * s0);
        dl[3] += w * y * (c1 - c0);       dl[4] += w * y * (s1 - s0);
    }
    atomicAdd(stats + 0, 1);
    if (old != idx) atomicAdd(stats + 1, 1);
    if (stage) atomicAdd(stats + 2, 1);
}

// Stage A scan by one warp: lanes over grid points, snapshot in shared memory.  Outside the MUSIC guard zone
// 1/(M - P/E) is monotone in the beam power P, so the argmax is taken on P; the guard zone and every pair that the
// error bound cannot separate go to stage B.  ds_norm bounds |dS|_2 of the snapshot.
constexpr int SW_MAX_PER_LANE = 12;            // grids up to 384 points keep P_h in registers
__device__ ScanRes scan_warp(const AngleFix& q, const double2* snap, int A, double energy, double ds_norm) {
    const int lane = threadIdx.x & 31;
    double pl[SW_MAX_PER_LANE];
    double bestp = -1.0;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < SW_MAX_PER_LANE; ++j) {
        const int g = lane + 32 * j;
        pl[j] = -1.0;
        if (g < q.G) {
            const double2 b = beam(q, snap, A, g);
            const double pwr = b.x * b.x + b.y * b.y;
            pl[j] = pwr;
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
    auto test = [&](int h, double ph) {
        if (h == bi || und) return;
        const double gap = pg - ph;
        // cheap sufficient tests: |conj(b_g) a_g - conj(b_h) a_h| <= sqrt(A) (|b_g| + |b_h|) <= 2 sqrt(A) |b_g|
        if (gap > 4.0 * wg * ds_norm) return;
        if (gap > 2.0 * (wg + sqrt((double)A * ph)) * ds_norm) return;
        const double2 bh = beam(q, snap, A, h);
        double cr = 0, ci = 0;
        for (int m = 0; m < A; ++m) {
            const double2 wgm = q.steer[(size_t)m * q.G + bi], wh = q.steer[(size_t)m * q.G + h];
            cr += wgm.x * wh.x + wgm.y * wh.y;      // conj(a_g) a_h
            ci += wgm.x * wh.y - wgm.y * wh.x;
        }
        // |conj(b_g) a_g - conj(b_h) a_h|^2 = A (|b_g|^2 + |b_h|^2) - 2 Re(b_g conj(b_h) (a_g^H a_h))
        const double tr = bg.x * bh.x + bg.y * bh.y, ti = bg.y * bh.x - bg.x * bh.y;
        const double n2 = fmax(0.0, (double)A * (pg + ph) - 2.0 * (tr * cr - ti * ci));
        if (gap <= 2.0 * sqrt(n2) * ds_norm) und = true;
    };
#pragma unroll
    for (int j = 0; j < SW_MAX_PER_LANE; ++j) {
        const int h = lane + 32 * j;
        if (h < q.G) test(h, pl[j]);
    }
    for (int h = lane + 32 * SW_MAX_PER_LANE; h < q.G; h += 32) {
        const double2 bh = beam(q, snap, A, h);
        test(h, bh.x * bh.x + bh.y * bh.y);
    }
    res.undecided = __any_sync(0xffffffffu, und);
    return res;
}

// exact scan of an exact snapshot by one warp: the method's pseudo-spectrum with the reference's guard
// (angle_estimation.py:143-152), first-index argmax (:173)
__device__ int scan_exact_warp(const AngleFix& q, const double2* snap, int A, double energy) {
    const int lane = threadIdx.x & 31;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int g = lane; g < q.G; g += 32) {
        const double2 b = beam(q, snap, A, g);
        const double pwr = b.x * b.x + b.y * b.y;
        double val = pwr;
        if (q.method == RS_METHOD_MUSIC) {
            const double den = fabs((double)A - (energy > 0 ? pwr / energy : 1.0));
            val = den > 1e-12 ? 1.0 / den : 0.0;
        }
        if (val > best) { best = val; bi = g; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    return bi;
}

// -------- stage A: one CTA per 8 segments.  Every warp lists the flagged cells of its own segment (in list order),
// the CTA's warps then EVALUATE all listed cells round-robin -- the fp64 scans are the expensive part and segments hold
// anything from zero to a dozen cells -- and finally every warp COMMITS its own segment's results in list order
// (velocity-sum corrections, stage-B slots), which keeps the outcome independent of who evaluated what.
struct StageAResult {
    double dl[5];      // correction of the segment's velocity sums (zero when undecided or unchanged)
    int undecided, r, d, li;
};

// Everything the evaluation and the commit of one listed cell read from global memory, fetched by ONE thread per cell for
// all cells of the CTA at once (A <= SA_AMAX): the chain leader -> key / old index / phase / flags -> snapshot is three
// memory latencies for the whole CTA instead of eight per cell, which is what the kernel's time was (14 % issue, 16 %
// occupancy: 0.22 ms per 1000 frames for 0.5 GFLOP).
constexpr int SA_STAGE = 64;          // cells staged per round (a CTA of the benchmark scene lists ~14)
constexpr int SA_AMAX = 16;
struct StagedCell {
    uint32_t pos;      // position of the cell's first entry in its segment
    int k, r, d, old, w, li;
    float y;
    uint8_t fl[8];     // flags of its first 8 entries
};

__global__ void __launch_bounds__(RC_THREADS, 4)
recheck_angles_kernel(CubeView v, AngleFix q, int32_t* __restrict__ work_idx, int32_t* __restrict__ work_cnt,
                      int32_t* __restrict__ frame_cnt, int4* __restrict__ frame_list, int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];                        // [RC_WARPS][A] snapshots
    __shared__ int items[RC_WARPS][RS_TIE_LIST_CAP];        // flagged leader indices per segment, ascending
    __shared__ int cnt[RC_WARPS], off[RC_WARPS + 1];
    __shared__ double ds_seg[RC_WARPS];
    __shared__ StageAResult res[RC_WARPS * RS_TIE_LIST_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = blockIdx.x * RC_WARPS + wid;
    const int A = v.A;
    double2* sw = smd + wid * A;
    const bool seg_ok = seg < q.nseg_total;
    const size_t base = (size_t)seg * q.seg_cap;
    const int f = seg_ok ? seg / q.nseg : 0;
    const int n = seg_ok ? q.det_nlead[seg] : 0;
    const int ntie = !seg_ok ? 0 : (q.det_ntie != nullptr ? q.det_ntie[seg] : 0x7fffffff);
    const bool has_work = seg_ok && n > 0 && ntie != 0;
    const bool listed = has_work && q.det_tielist != nullptr && ntie <= RS_TIE_LIST_CAP;

    auto undecided = [&](int li) {
        const uint8_t f8 = q.det_flags[base + (q.det_lead[base + li] & 0xFFFFu)];
        return (f8 & (RS_FLAG_TIE | RS_FLAG_GUARD)) && !(f8 & RS_FLAG_FIXED);
    };
    // fp64 scan of leader li of segment sg (whole warp); result valid on lane 0
    auto evaluate = [&](int sg, int li, double ds_norm, StageAResult& out) {
        const size_t sbase = (size_t)sg * q.seg_cap;
        const int sf = sg / q.nseg;
        const uint32_t ldj = q.det_lead[sbase + li];
        const size_t o = sbase + (ldj & 0xFFFFu);
        const int k = (int)(ldj >> 16);
        int a, r, d;
        rs_split_key(q.det_key[o], a, r, d);
        const float2* cell = q.rds + ((size_t)sf * q.R + r) * A * q.D + d;
        __syncwarp();
        for (int m = lane; m < A; m += 32) {
            const float2 x = cell[(size_t)m * q.D];
            sw[m] = make_double2((double)x.x, (double)x.y);
        }
        __syncwarp();
        double energy = 0;
        for (int m = 0; m < A; ++m) energy += sw[m].x * sw[m].x + sw[m].y * sw[m].y;
        // |dS|_2 <= fft_eps * rms * sqrt(A)  (noise-like part)  +  1.5e-7 |s|_2  (part that scales with the cell)
        const ScanRes sr = scan_warp(q, sw, A, energy, ds_norm + 1.5e-7 * sqrt(energy));
        if (lane == 0) {
            for (int j = 0; j < 5; ++j) out.dl[j] = 0;
            out.undecided = sr.undecided ? 1 : 0;
            out.r = r; out.d = d; out.li = li;
            if (!sr.undecided) apply_angle(q, o, k, sr.idx, out.dl, stats, 0);      // this cell's entries only
        }
    };

    // ---- phase 1: this warp's segment: noise level of its frame, ordered list of its flagged cells
    if (lane == 0) cnt[wid] = 0;
    if (has_work) {
        // rms(|X|) of the frame from the per-tile power sums; |dS|_2 <= fft_eps * rms * sqrt(A)
        double psum = 0;
        for (int sg = lane; sg < q.nseg; sg += 32) psum += (double)q.det_psum[(size_t)f * q.nseg + sg];
        psum = warp_sum(psum);
        if (lane == 0) ds_seg[wid] = fmax(q.fft_eps * sqrt(psum / ((double)q.R * q.D * A)) * sqrt((double)A), 1e-300);
    }
    if (listed) {
        // the list rs_angles wrote (in atomic order): rank-sort it
        int li = lane < ntie ? q.det_tielist[(size_t)seg * RS_TIE_LIST_CAP + lane] : 0x7fffffff;
        const bool live = lane < ntie && li >= 0 && li < n && undecided(li);
        if (!live) li = 0x7fffffff;
        int rank = 0;
        for (int t = 0; t < ntie; ++t) rank += __shfl_sync(0xffffffffu, li, t) < li ? 1 : 0;
        if (live) items[wid][rank] = li;
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        if (lane == 0) cnt[wid] = __popc(lm);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < RC_WARPS; ++w) { off[w] = run; run += cnt[w]; }
        off[RC_WARPS] = run;
    }
    __syncthreads();

    // ---- phase 2: all listed cells of the CTA, round-robin over its warps
    const int total = off[RC_WARPS];
    if (A <= SA_AMAX) {
        __shared__ StagedCell st[SA_STAGE];
        for (int t0 = 0; t0 < total; t0 += SA_STAGE) {
            const int nt = min(SA_STAGE, total - t0);
            if ((int)threadIdx.x < nt) {                    // 2a: one thread per cell gathers what the cell needs
                const int t = t0 + threadIdx.x;
                int w = 0;
                while (t >= off[w + 1]) ++w;
                StagedCell& c = st[threadIdx.x];
                c.w = w;
                c.li = items[w][t - off[w]];
                const int sg = blockIdx.x * RC_WARPS + w;
                const size_t sbase = (size_t)sg * q.seg_cap;
                const uint32_t ldj = q.det_lead[sbase + c.li];
                c.pos = ldj & 0xFFFFu;
                c.k = (int)(ldj >> 16);
                const size_t o = sbase + c.pos;
                int a;
                rs_split_key(q.det_key[o], a, c.r, c.d);
                c.old = q.det_aidx[o];
                c.y = q.det_phase[o];
#pragma unroll
                for (int e = 0; e < 8; ++e) c.fl[e] = e < c.k ? q.det_flags[o + e] : (uint8_t)0;
                const float2* cell = q.rds + ((size_t)(sg / q.nseg) * q.R + c.r) * A * q.D + c.d;
                double2* sn = smd + (size_t)threadIdx.x * A;
                for (int m = 0; m < A; ++m) {
                    const float2 x = cell[(size_t)m * q.D];
                    sn[m] = make_double2((double)x.x, (double)x.y);
                }
            }
            __syncthreads();
            for (int j = wid; j < nt; j += RC_WARPS) {      // 2b: fp64 scans out of shared memory
                const StagedCell& c = st[j];
                const double2* sn = smd + (size_t)j * A;
                double energy = 0;
                for (int m = 0; m < A; ++m) energy += sn[m].x * sn[m].x + sn[m].y * sn[m].y;
                // |dS|_2 <= fft_eps * rms * sqrt(A)  (noise-like part)  +  1.5e-7 |s|_2  (part that scales with the cell)
                const ScanRes sr = scan_warp(q, sn, A, energy, ds_seg[c.w] + 1.5e-7 * sqrt(energy));
                if (lane == 0) {
                    StageAResult& out = res[t0 + j];
                    for (int i = 0; i < 5; ++i) out.dl[i] = 0;
                    out.undecided = sr.undecided ? 1 : 0;
                    out.r = c.r; out.d = c.d; out.li = c.li;
                    if (!sr.undecided) {                    // apply_angle() from the staged values
                        const size_t o = (size_t)(blockIdx.x * RC_WARPS + c.w) * q.seg_cap + c.pos;
                        const int idx = sr.idx;
                        const float deg = q.grid_deg[idx];
                        int live = 0;
                        for (int e = 0; e < c.k; ++e) {
                            const uint8_t fl = e < 8 ? c.fl[e] : q.det_flags[o + e];
                            live += (fl & RS_FLAG_DROPPED) ? 0 : 1;
                            q.det_flags[o + e] = fl | RS_FLAG_FIXED;
                            q.det_aidx[o + e] = idx;
                            q.det_adeg[o + e] = deg;
                        }
                        if (c.old != idx && live > 0) {
                            const double wgt = (double)live, y = (double)c.y;
                            const double c0 = q.grid_cs[2 * c.old], s0 = q.grid_cs[2 * c.old + 1];
                            const double c1 = q.grid_cs[2 * idx], s1 = q.grid_cs[2 * idx + 1];
                            out.dl[0] = wgt * (c1 * c1 - c0 * c0); out.dl[1] = wgt * (s1 * s1 - s0 * s0);
                            out.dl[2] = wgt * (c1 * s1 - c0 * s0);
                            out.dl[3] = wgt * y * (c1 - c0);       out.dl[4] = wgt * y * (s1 - s0);
                        }
                        atomicAdd(stats + 0, 1);
                        if (c.old != idx) atomicAdd(stats + 1, 1);
                    }
                }
            }
            __syncthreads();
        }
    } else {
        for (int t = wid; t < total; t += RC_WARPS) {
            int w = 0;
            while (t >= off[w + 1]) ++w;
            evaluate(blockIdx.x * RC_WARPS + w, items[w][t - off[w]], ds_seg[w], res[t]);
        }
        __syncthreads();
    }

    // ---- phase 3: commit this warp's segment in list order
    if (!seg_ok) return;
    double dl[5] = {0, 0, 0, 0, 0};
    int nb = 0, lost = 0;
    auto commit = [&](const StageAResult& rr) {             // lane 0
        if (!rr.undecided) {
            for (int j = 0; j < 5; ++j) dl[j] += rr.dl[j];
        } else if (nb < RB_MAX) {
            const int slot = atomicAdd(frame_cnt + f, 1);
            if (slot < FB_CAP) {
                work_idx[(size_t)seg * RB_MAX + nb] = rr.li;
                frame_list[(size_t)f * FB_CAP + slot] = make_int4(seg * RB_MAX + nb, rr.r, rr.d, 0);
                ++nb;
            } else {
                ++lost;
            }
        } else {
            ++lost;
        }
    };
    if (listed) {
        if (lane == 0)
            for (int k = 0; k < cnt[wid]; ++k) commit(res[off[wid] + k]);
    } else if (has_work) {
        // more flagged cells than the list holds (or no list): this warp scans its leaders and settles them itself
        StageAResult* mine = &res[wid * RS_TIE_LIST_CAP];    // after the barrier above nobody else uses the array
        const double ds_norm = ds_seg[wid];
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            unsigned mask = __ballot_sync(0xffffffffu, i < n && undecided(i));
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                evaluate(seg, c0 + j, ds_norm, *mine);
                if (lane == 0) commit(*mine);
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        if (q.ls_partials && has_work) {
            double* ps = q.ls_partials + (size_t)seg * 8;
            for (int j = 0; j < 5; ++j) ps[j] += dl[j];
        }
        work_cnt[seg] = nb;
        if (lost) atomicAdd(stats + 3, lost);
    }
}

// -------- stage B1: one CTA per (frame, antenna) reads that antenna's raw plane once per batch of NB_MAX undecided
// cells of the frame and forms its element of their fp64 snapshots
template <int NB>
__device__ __forceinline__ void b1_batch(const CubeView& v, const float2* plane, const double2* u, const int* kr,
                                         double2 (&acc)[NB_MAX]) {
    for (int s = threadIdx.x; s < v.S; s += RC_THREADS) {
        double2 Y[NB];
        column_dft<NB>(plane + s, v.S, v.C, u, Y);
        const double2 t = v.tab[s];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double2 p = dmul(Y[i], dmul(t, v.tw_s[(int)(((long long)s * kr[i]) % v.S)]));
            acc[i].x += p.x;
            acc[i].y += p.y;
        }
    }
}

__global__ void __launch_bounds__(RC_THREADS)
recheck_snapshots_kernel(CubeView v, const int32_t* __restrict__ frame_cnt, const int4* __restrict__ frame_list,
                         double2* __restrict__ work_snap) {
    extern __shared__ double2 smd[];
    double2* u = smd;                           // [NB_MAX][C]
    double2* red = u + NB_MAX * v.C;            // [RC_WARPS][NB_MAX]
    __shared__ double2 out[NB_MAX];
    __shared__ int kr_s[NB_MAX], kd_s[NB_MAX], slot_s[NB_MAX];
    const int a = blockIdx.x, f = blockIdx.y;
    const int nc = min(frame_cnt[f], FB_CAP);
    if (nc == 0) return;
    const float2* plane = v.cube + (((size_t)f * v.A + a) * v.C_total + v.chirp0) * v.S;
    for (int b0 = 0; b0 < nc; b0 += NB_MAX) {
        const int nb = min(NB_MAX, nc - b0);
        if ((int)threadIdx.x < NB_MAX) {
            const int i = threadIdx.x;
            if (i < nb) {
                const int4 e = frame_list[(size_t)f * FB_CAP + b0 + i];
                slot_s[i] = e.x;
                kr_s[i] = unshift(e.y, v.S);
                kd_s[i] = unshift(e.z, v.C);
            } else {
                slot_s[i] = -1; kr_s[i] = 0; kd_s[i] = 0;
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < NB_MAX * v.C; idx += RC_THREADS) {
            const int i = idx / v.C, c = idx - i * v.C;
            u[idx] = i < nb ? v.tw_c[(int)(((long long)c * kd_s[i]) % v.C)] : make_double2(0, 0);
        }
        __syncthreads();
        int kr[NB_MAX];
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) kr[i] = kr_s[i];
        double2 acc[NB_MAX];
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) acc[i] = make_double2(0, 0);
        if (nb <= 1) b1_batch<1>(v, plane, u, kr, acc);
        else if (nb <= 2) b1_batch<2>(v, plane, u, kr, acc);
        else if (nb <= 3) b1_batch<3>(v, plane, u, kr, acc);
        else if (nb <= 4) b1_batch<4>(v, plane, u, kr, acc);
        else if (nb <= 6) b1_batch<6>(v, plane, u, kr, acc);
        else b1_batch<8>(v, plane, u, kr, acc);
        block_reduce<NB_MAX>(acc, nb, red, out);
        if ((int)threadIdx.x < nb) {
            const int i = threadIdx.x;
            work_snap[(size_t)slot_s[i] * v.A + a] = (v.dc && kr_s[i] == 0) ? make_double2(0, 0) : out[i];
        }
        __syncthreads();
    }
}

// -------- stage B2: one warp per segment scans the exact snapshots and settles the remaining cells in list order
__global__ void __launch_bounds__(RC_THREADS)
recheck_finish_kernel(CubeView v, AngleFix q, const int32_t* __restrict__ work_idx, const int32_t* __restrict__ work_cnt,
                      const double2* __restrict__ work_snap, int32_t* __restrict__ stats) {
    extern __shared__ double2 smd[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = blockIdx.x * RC_WARPS + wid;
    if (seg >= q.nseg_total) return;
    const int cnt = work_cnt[seg];
    if (cnt == 0) return;
    const int A = v.A;
    double2* snap = smd + wid * A;
    const size_t base = (size_t)seg * q.seg_cap;
    double dl[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < cnt; ++k) {
        const uint32_t ld = q.det_lead[base + work_idx[(size_t)seg * RB_MAX + k]];
        const size_t o = base + (ld & 0xFFFFu);
        const int mult = (int)(ld >> 16);
        __syncwarp();
        for (int m = lane; m < A; m += 32) snap[m] = work_snap[((size_t)seg * RB_MAX + k) * A + m];
        __syncwarp();
        double energy = 0;
        for (int m = 0; m < A; ++m) energy += snap[m].x * snap[m].x + snap[m].y * snap[m].y;
        const int idx = scan_exact_warp(q, snap, A, energy);
        if (lane == 0) apply_angle(q, o, mult, idx, dl, stats, 1);
    }
    if (lane == 0 && q.ls_partials) {
        double* ps = q.ls_partials + (size_t)seg * 8;
        for (int j = 0; j < 5; ++j) ps[j] += dl[j];
    }
}

}  // namespace

extern "C" int rs_recheck_detections_f64(const void* cube, const void* table128, const void* tw_s128,
                                         const void* tw_c128, int C_total, int chirp0, int dc_removal,
                                         double thr_power64, const uint32_t* det_key, uint8_t* det_flags,
                                         const int32_t* det_count, const int32_t* det_nnear, int seg_cap,
                                         int nseg_per_frame, int F, int A, int C, int S, const int32_t* det_aidx,
                                         const float* det_phase, const double* grid_cs, double* ls_partials,
                                         int32_t* stats, void* stream) {
    RS_CHECK_ARG(cube && table128 && tw_s128 && tw_c128 && det_key && det_flags && det_count && stats,
                 "rs_recheck_detections_f64: null pointer");
    RS_CHECK_ARG(ls_partials == nullptr || (det_aidx && det_phase && grid_cs),
                 "rs_recheck_detections_f64: ls_partials needs det_aidx, det_phase and grid_cs");
    RS_CHECK_ARG(F > 0 && A > 0 && C > 0 && S > 0 && chirp0 >= 0 && chirp0 + C <= C_total && seg_cap > 0 &&
                     nseg_per_frame > 0,
                 "rs_recheck_detections_f64: bad dims");
    CubeView v{(const float2*)cube, (const double2*)table128, (const double2*)tw_s128, (const double2*)tw_c128,
               A, C_total, chirp0, C, S, dc_removal};
    const size_t smem = (size_t)(3 * C + RC_WARPS * 9) * sizeof(double2);
    if (smem > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_recheck_detections_f64: needs %zu B of shared memory", smem);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(recheck_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), (cudaStream_t)stream);
    const long long nseg_total = (long long)F * nseg_per_frame;
    RS_CHECK_ARG(nseg_total < (1ll << 31), "rs_recheck_detections_f64: too many segments");
    const unsigned blocks = (unsigned)((nseg_total + RC_WARPS - 1) / RC_WARPS);
    recheck_detect_kernel<<<blocks, RC_THREADS, smem, (cudaStream_t)stream>>>(
        v, thr_power64, det_key, det_flags, det_count, det_nnear, seg_cap, nseg_per_frame, (int)nseg_total, S, C,
        det_aidx, det_phase, grid_cs, ls_partials, stats);
    RS_CHECK_LAUNCH("rs_recheck_detections_f64");
    return RS_OK;
}

extern "C" int rs_recheck_angles_f64(const void* cube, const void* table128, const void* tw_s128, const void* tw_c128,
                                     int C_total, int chirp0, int dc_removal, const void* rds, const void* steer128,
                                     const float* grid_deg, const double* grid_cs, int G, int method, double fft_eps,
                                     const float* det_psum, const int32_t* det_ntie, const int32_t* det_tielist,
                                     const uint32_t* det_key,
                                     const uint32_t* det_lead, const int32_t* det_nlead, uint8_t* det_flags,
                                     int32_t* det_aidx, float* det_adeg, const float* det_phase, double* ls_partials,
                                     int seg_cap, int nseg_per_frame, int F, int A, int C, int S, int32_t* work_idx,
                                     int32_t* work_cnt, void* work_snap, int32_t* frame_cnt, int32_t* frame_list,
                                     int32_t* stats, void* stream) {
    RS_CHECK_ARG(cube && table128 && tw_s128 && tw_c128 && rds && steer128 && grid_deg && grid_cs && det_psum &&
                     det_key && det_lead && det_nlead && det_flags && det_aidx && det_adeg && det_phase && work_idx &&
                     work_cnt && work_snap && frame_cnt && frame_list && stats,
                 "rs_recheck_angles_f64: null pointer");
    RS_CHECK_ARG(method == RS_METHOD_MUSIC || method == RS_METHOD_BEAMFORMING, "rs_recheck_angles_f64: grid methods only");
    RS_CHECK_ARG(F > 0 && F <= 65535 && A >= 2 && C > 0 && S > 0 && G > 0 && chirp0 >= 0 && chirp0 + C <= C_total &&
                     fft_eps >= 0 && nseg_per_frame > 0,
                 "rs_recheck_angles_f64: bad dims");
    RS_CHECK_ARG(((uintptr_t)frame_list & 15u) == 0, "rs_recheck_angles_f64: frame_list must be 16-byte aligned");
    const long long nseg_total = (long long)F * nseg_per_frame;
    RS_CHECK_ARG(nseg_total < (1ll << 27), "rs_recheck_angles_f64: too many segments");
    CubeView v{(const float2*)cube, (const double2*)table128, (const double2*)tw_s128, (const double2*)tw_c128,
               A, C_total, chirp0, C, S, dc_removal};
    AngleFix q{(const float2*)rds, (const double2*)steer128, grid_deg, grid_cs, G, method, fft_eps, det_psum, det_ntie,
               det_tielist, det_key, det_lead, det_nlead, det_flags, det_aidx, det_adeg, det_phase, ls_partials, seg_cap,
               nseg_per_frame, (int)nseg_total, S, C};
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_a = (size_t)(A <= SA_AMAX ? SA_STAGE : RC_WARPS) * A * sizeof(double2);
    const size_t smem_b1 = (size_t)(NB_MAX * C + RC_WARPS * NB_MAX) * sizeof(double2);
    if (smem_b1 > (size_t)rs_smem_optin_limit() || smem_a > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_recheck_angles_f64: needs %zu B of shared memory", smem_b1 > smem_a ? smem_b1 : smem_a);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(recheck_angles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    cudaFuncSetAttribute(recheck_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    cudaFuncSetAttribute(recheck_snapshots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b1);
    cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), st);
    cudaMemsetAsync(frame_cnt, 0, (size_t)F * sizeof(int32_t), st);
    const unsigned blocks = (unsigned)((nseg_total + RC_WARPS - 1) / RC_WARPS);
    recheck_angles_kernel<<<blocks, RC_THREADS, smem_a, st>>>(v, q, work_idx, work_cnt, frame_cnt, (int4*)frame_list,
                                                              stats);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage A)");
    recheck_snapshots_kernel<<<dim3((unsigned)A, (unsigned)F), RC_THREADS, smem_b1, st>>>(
        v, frame_cnt, (const int4*)frame_list, (double2*)work_snap);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage B1)");
    recheck_finish_kernel<<<blocks, RC_THREADS, smem_a, st>>>(v, q, work_idx, work_cnt, (const double2*)work_snap,
                                                              stats);
    RS_CHECK_LAUNCH("rs_recheck_angles_f64(stage B2)");
    return RS_OK;
}
