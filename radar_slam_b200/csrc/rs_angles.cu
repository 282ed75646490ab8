// Per-detection angle estimation on the RDS rds[F][R][A][D] (rs_common.cuh):
// the A-channel snapshot of cell (r, d) is A elements at stride D.
//
// Replaces AngleEstimator.extract_spatial_signature / music_spectrum / estimate_angle_* /
// process_targets (angle_estimation.py:67-309).
//
// MUSIC in the reference is single-snapshot: R = s s^H (angle_estimation.py:127), eigh, noise
// subspace = all but the top eigenvector, so a^H E_n E_n^H a = M - |a^H s|^2 / |s|^2 exactly and
// argmax(1/den) == argmax |a^H s|^2 except inside the den <= 1e-12 guard (:149), which fp32
// cannot resolve: those detections get RS_FLAG_GUARD and are re-evaluated in fp64 by the caller.
// The scan uses the lag form  |a^H s|^2 = R_0 + 2 sum_k Re(R_k e^{-ik phi}),  R_k = sum_n s[n+k] conj(s[n]),
// which halves the multiply-adds per grid point; cos/sin(k phi_g) come from an fp64-built table.
//
// ESPRIT (angle_estimation.py:195-221) reduces, for one snapshot, to the principal eigenvector of a
// 2x2 Hermitian matrix; it is evaluated in fp64 from the fp32 snapshot.
#include <type_traits>
#include <cstdlib>
#include <cuda_fp16.h>
#include "rs_common.cuh"

namespace {

constexpr int ANG_THREADS = 128;

struct AngleArgs {
    const float2* rds;
    const float* scan_table;
    int scan_stride;
    const float2* steer;
    const float* grid_deg;
    int G;
    int method;
    float tie_eps;
    double esprit_scale;
    const uint32_t* det_key;
    const uint32_t* det_lead;     // per segment: position | multiplicity << 16 of each distinct cell
    const int32_t* det_nlead;
    int32_t* det_ntie;            // per segment: cells flagged TIE / GUARD (zeroed by the entry point; may be null)
    int32_t* det_tielist;         // per segment: leader indices of the first RS_TIE_LIST_CAP flagged cells (unordered; may be null)
    uint8_t* det_flags;
    int32_t* det_aidx;
    float* det_adeg;
    float* det_phase;
    int seg_cap, nseg_per_frame, R, D, A;
    float* det_power;             // optional: |X|^2 of every entry (lists from rs_range_doppler_detect without det_power)
    const int32_t* det_nnear;     // optional: RS_FLAG_NEARMAX entries per segment (rs_detect): 0 = no entry can be DROPPED
};

// |X|^2 of the k entries of a cell from its snapshot (the entries are the antennas that flagged the cell; their keys sit
// in the sector the leader's key came from).  Same expression as the detection kernels: the values are bit-identical.
template <int AP>
__device__ __forceinline__ void emit_power(const AngleArgs& p, size_t o, int k, const float2 (&s)[AP]) {
    if (p.det_power == nullptr) return;
    for (int e = 0; e < k; ++e) {
        const int a = (int)(p.det_key[o + e] >> 24);
        float pw = 0.f;
#pragma unroll
        for (int m = 0; m < AP; ++m)
            if (m == a) pw = fmaf(s[m].x, s[m].x, s[m].y * s[m].y);
        p.det_power[o + e] = pw;
    }
}

// write one cell's result to all of its detections (same snapshot on every antenna that flagged the cell)
// returns how many of them are live (not RS_FLAG_DROPPED): the weight of the cell in the velocity sums
// clean: the segment holds no RS_FLAG_NEARMAX entry (det_nnear == 0), hence no RS_FLAG_DROPPED one (only near-miss candidates
// and the detection recheck, which touches NEARMAX entries only, carry / set that flag): every entry is live and the flags
// need not be read -- one dependent global load per entry less in 99.98 % of the segments.
template <bool BATCH = false>
__device__ __forceinline__ int emit(const AngleArgs& p, int seg, int lead_i, size_t o, int k, int aidx, float adeg,
                                    float phase, uint8_t extra_flags, bool clean = false) {
    int live = 0;
    if ((extra_flags & (RS_FLAG_TIE | RS_FLAG_GUARD)) && p.det_ntie) {
        const int slot = atomicAdd(p.det_ntie + seg, 1);
        if (p.det_tielist && slot < RS_TIE_LIST_CAP) p.det_tielist[(size_t)seg * RS_TIE_LIST_CAP + slot] = lead_i;
    }
    if (clean) {
        for (int e = 0; e < k; ++e) {
            p.det_aidx[o + e] = aidx;
            p.det_adeg[o + e] = adeg;
            p.det_phase[o + e] = phase;
            if (extra_flags) p.det_flags[o + e] = extra_flags;      // the entry's flags were 0
        }
        return k;
    }
    // BATCH (the tcgen05 kernel: few resident warps, registers to spare): the flags of up to eight entries are fetched
    // together -- one latency instead of k; longer groups finish in the loop below
    int e0 = 0;
    if (BATCH) {
        uint8_t fl8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) fl8[e] = e < k ? p.det_flags[o + e] : (uint8_t)RS_FLAG_DROPPED;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e < k) {
                p.det_aidx[o + e] = aidx;
                p.det_adeg[o + e] = adeg;
                p.det_phase[o + e] = phase;
                live += (fl8[e] & RS_FLAG_DROPPED) ? 0 : 1;
                if (extra_flags) p.det_flags[o + e] = fl8[e] | extra_flags;
            }
        }
        e0 = 8;
    }
    for (int e = e0; e < k; ++e) {
        p.det_aidx[o + e] = aidx;
        p.det_adeg[o + e] = adeg;
        p.det_phase[o + e] = phase;
        const uint8_t fl = p.det_flags[o + e];
        live += (fl & RS_FLAG_DROPPED) ? 0 : 1;
        if (extra_flags) p.det_flags[o + e] = fl | extra_flags;
    }
    return live;
}

// ESPRIT closed form from a snapshot held in registers (s[0..M-1]); returns degrees.
template <int AP>
__device__ __forceinline__ float esprit_deg(const float2 (&s)[AP], int M, double scale) {
    double alpha = 0, gamma = 0, br = 0, bi = 0;
#pragma unroll
    for (int i = 0; i < AP - 1; ++i) {
        if (i < M - 1) {
            const double xr = s[i].x, xi = s[i].y, yr = s[i + 1].x, yi = s[i + 1].y;
            alpha += xr * xr + xi * xi;
            gamma += yr * yr + yi * yi;
            br += xr * yr + xi * yi;     // conj(x) * y
            bi += xr * yi - xi * yr;
        }
    }
    const double half = 0.5 * (alpha - gamma);
    const double lam = 0.5 * (alpha + gamma) + sqrt(half * half + br * br + bi * bi);
    // eigenvector of [[alpha, beta],[conj(beta), gamma]] for lam: take the better conditioned row
    double v0r, v0i, v1r, v1i;
    const double na = br * br + bi * bi + (lam - alpha) * (lam - alpha);
    const double nb = (lam - gamma) * (lam - gamma) + br * br + bi * bi;
    if (na >= nb) { v0r = br; v0i = bi; v1r = lam - alpha; v1i = 0; }
    else          { v0r = lam - gamma; v0i = 0; v1r = br; v1i = -bi; }
    // u_i = v0 s_i + v1 s_{i+1};  num = sum_{i<M-2} conj(u_i) u_{i+1}
    double nr = 0, ni = 0, pr = 0, pi = 0;
#pragma unroll
    for (int i = 0; i < AP - 1; ++i) {
        if (i < M - 1) {
            const double xr = s[i].x, xi = s[i].y, yr = s[i + 1].x, yi = s[i + 1].y;
            const double ur = v0r * xr - v0i * xi + v1r * yr - v1i * yi;
            const double ui = v0r * xi + v0i * xr + v1r * yi + v1i * yr;
            if (i > 0) {
                nr += pr * ur + pi * ui;
                ni += pr * ui - pi * ur;
            }
            pr = ur; pi = ui;
        }
    }
    const double phase = atan2(ni, nr);
    return (float)(asin(phase * scale) * (180.0 / 3.14159265358979323846));
}

// thread-per-detection kernel for A <= AP (AP in {2,4,8,16})
template <int AP>
__global__ void __launch_bounds__(ANG_THREADS) angles_small_kernel(AngleArgs p) {
    extern __shared__ float tab[];   // [G][scan_stride]
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    if (n == 0) return;
    const bool scan = p.method != RS_METHOD_ESPRIT;
    if (scan) {
        const int nt = p.G * p.scan_stride;
        for (int i = threadIdx.x; i < nt; i += blockDim.x) tab[i] = p.scan_table[i];
        __syncthreads();
    }
    const int f = seg / p.nseg_per_frame;
    const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
    const int M = p.A;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t ld = p.det_lead[(size_t)seg * p.seg_cap + i];
        const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
        const int mult = (int)(ld >> 16);
        int a, r, d;
        rs_split_key(p.det_key[o], a, r, d);
        const float2* cell = frame + (size_t)r * M * p.D + d;
        float2 s[AP];
#pragma unroll
        for (int m = 0; m < AP; ++m) s[m] = (m < M) ? __ldg(cell + (size_t)m * p.D) : make_float2(0.f, 0.f);

        // inter-antenna phase angle(s[1] conj(s[0]))  (velocity_solver.py:136)
        const float pr = s[1].x * s[0].x + s[1].y * s[0].y;
        const float pi = s[1].y * s[0].x - s[1].x * s[0].y;
        const float phase = atan2f(pi, pr);

        uint8_t flags = 0;
        if (!scan) {
            emit(p, seg, i, o, mult, -1, esprit_deg<AP>(s, M, p.esprit_scale), phase, 0);
            continue;
        }
        // lags R_k, k = 0..AP-1
        float rr[AP], ri[AP];
#pragma unroll
        for (int k = 0; k < AP; ++k) {
            float xr = 0.f, xi = 0.f;
#pragma unroll
            for (int m = 0; m + k < AP; ++m) {
                xr = fmaf(s[m + k].x, s[m].x, xr);
                xr = fmaf(s[m + k].y, s[m].y, xr);
                xi = fmaf(s[m + k].y, s[m].x, xi);
                xi = fmaf(-s[m + k].x, s[m].y, xi);
            }
            rr[k] = xr;
            ri[k] = xi;
        }
        float best = -3.0e38f, second = -3.0e38f;
        int bi = 0;
        const float* t = tab;
        for (int g = 0; g < p.G; ++g, t += p.scan_stride) {
            float acc = 0.f;
#pragma unroll
            for (int k = 1; k < AP; ++k) {
                acc = fmaf(rr[k], t[2 * (k - 1)], acc);
                acc = fmaf(ri[k], t[2 * (k - 1) + 1], acc);
            }
            if (acc > best) { second = best; best = acc; bi = g; }
            else if (acc > second) second = acc;
        }
        const float pbest = rr[0] + 2.f * best;
        if (2.f * (best - second) <= p.tie_eps * fabsf(pbest)) flags |= RS_FLAG_TIE;
        if (p.method == RS_METHOD_MUSIC) {
            const float full = (float)M * rr[0];
            if (full - pbest <= 1e-4f * full) flags |= RS_FLAG_GUARD;
        }
        emit(p, seg, i, o, mult, bi, p.grid_deg[bi], phase, flags);
    }
}

// ---------------------------------------------------------------------------------------------
// Scan kernel, A <= AP: each thread carries ND detections through the grid at once so one
// broadcast shared-memory read of the (cos k phi, sin k phi) row feeds ND * 2(AP-1) FMAs.
// SYM: the grid is symmetric about 0 (grid[G-1-g] == -grid[g]), so cos k phi is shared by the
// pair and sin k phi flips sign:  P(+-theta) = E +- O  with  E = sum Re R_k cos, O = sum Im R_k sin,
// i.e. (AP-1) FMAs per grid point instead of 2(AP-1).  First-index argmax (np.argmax) is kept by
// tracking the ascending left half with '>' and the descending right half with '>='.
// The fp64 normal-equation sums of the velocity solve are accumulated here per segment
// (deterministic block reduction), so the solve does not re-read the detection lists.
// ---------------------------------------------------------------------------------------------
struct Track {
    float best, second;
    int idx;
};
__device__ __forceinline__ void track_first(Track& t, float v, int g) {     // keeps the earliest index on ties
    const bool up = v > t.best;
    t.second = fmaxf(t.second, fminf(t.best, v));
    t.best = fmaxf(t.best, v);
    t.idx = up ? g : t.idx;
}
__device__ __forceinline__ void track_last(Track& t, float v, int g) {      // keeps the latest processed on ties
    const bool up = v >= t.best;
    t.second = fmaxf(t.second, fminf(t.best, v));
    t.best = fmaxf(t.best, v);
    t.idx = up ? g : t.idx;
}

template <int AP, int ND, bool SYM>
__global__ void __launch_bounds__(ANG_THREADS) angles_scan_kernel(AngleArgs p, const double* __restrict__ grid_cs,
                                                                   double* __restrict__ ls_partials) {
    extern __shared__ float tab[];   // [rows][scan_stride]
    __shared__ double red[ANG_THREADS / 32][8];
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    double acc_ls[7] = {0, 0, 0, 0, 0, 0, 0};
    if (n > 0) {
        const int rows = SYM ? (p.G + 1) / 2 : p.G;
        for (int i = threadIdx.x; i < rows * p.scan_stride; i += blockDim.x) tab[i] = p.scan_table[i];
        __syncthreads();
        const int f = seg / p.nseg_per_frame;
        const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
        const int M = p.A;
        for (int base = 0; base < n; base += ND * ANG_THREADS) {
            float rr[ND][AP], ri[ND][AP], yv[ND];
            size_t o[ND];
            int mult[ND];
            bool valid[ND];
#pragma unroll
            for (int q = 0; q < ND; ++q) {
                const int i = base + q * ANG_THREADS + threadIdx.x;
                valid[q] = i < n;
                const uint32_t ld = valid[q] ? p.det_lead[(size_t)seg * p.seg_cap + i] : (1u << 16);
                o[q] = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
                mult[q] = (int)(ld >> 16);
                float2 s[AP];
                if (valid[q]) {
                    int a, r, d;
                    rs_split_key(p.det_key[o[q]], a, r, d);
                    const float2* cell = frame + (size_t)r * M * p.D + d;
#pragma unroll
                    for (int m = 0; m < AP; ++m) s[m] = (m < M) ? __ldg(cell + (size_t)m * p.D) : make_float2(0.f, 0.f);
                    const float pr = s[1].x * s[0].x + s[1].y * s[0].y;
                    const float pi = s[1].y * s[0].x - s[1].x * s[0].y;
                    yv[q] = atan2f(pi, pr);
                } else {
                    yv[q] = 0.f;
#pragma unroll
                    for (int m = 0; m < AP; ++m) s[m] = make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < AP; ++k) {
                    float xr = 0.f, xi = 0.f;
#pragma unroll
                    for (int m = 0; m + k < AP; ++m) {
                        xr = fmaf(s[m + k].x, s[m].x, xr);
                        xr = fmaf(s[m + k].y, s[m].y, xr);
                        xi = fmaf(s[m + k].y, s[m].x, xi);
                        xi = fmaf(-s[m + k].x, s[m].y, xi);
                    }
                    rr[q][k] = xr;
                    ri[q][k] = xi;
                }
            }

            Track L[ND], Rt[ND];
#pragma unroll
            for (int q = 0; q < ND; ++q) {
                L[q] = Track{-3.0e38f, -3.0e38f, 0};
                Rt[q] = Track{-3.0e38f, -3.0e38f, 0};
            }
            if (SYM) {
                const int half = p.G / 2;
                const float* t = tab;
                for (int g = 0; g < half; ++g, t += p.scan_stride) {
                    float tv[2 * (AP - 1)];
#pragma unroll
                    for (int k4 = 0; k4 < (2 * (AP - 1) + 3) / 4; ++k4) {
                        const float4 v = *reinterpret_cast<const float4*>(t + 4 * k4);
                        if (4 * k4 + 0 < 2 * (AP - 1)) tv[4 * k4 + 0] = v.x;
                        if (4 * k4 + 1 < 2 * (AP - 1)) tv[4 * k4 + 1] = v.y;
                        if (4 * k4 + 2 < 2 * (AP - 1)) tv[4 * k4 + 2] = v.z;
                        if (4 * k4 + 3 < 2 * (AP - 1)) tv[4 * k4 + 3] = v.w;
                    }
#pragma unroll
                    for (int q = 0; q < ND; ++q) {
                        float e = 0.f, od = 0.f;
#pragma unroll
                        for (int k = 1; k < AP; ++k) {
                            e = fmaf(rr[q][k], tv[2 * (k - 1)], e);
                            od = fmaf(ri[q][k], tv[2 * (k - 1) + 1], od);
                        }
                        track_first(L[q], e + od, g);
                        track_last(Rt[q], e - od, p.G - 1 - g);
                    }
                }
                if (p.G & 1) {      // the middle angle (0 deg for a symmetric grid): its own row, no partner
#pragma unroll
                    for (int q = 0; q < ND; ++q) {
                        float e = 0.f;
#pragma unroll
                        for (int k = 1; k < AP; ++k) {
                            e = fmaf(rr[q][k], t[2 * (k - 1)], e);
                            e = fmaf(ri[q][k], t[2 * (k - 1) + 1], e);
                        }
                        track_first(L[q], e, half);     // larger than every left index, '>' keeps earlier ties
                    }
                }
            } else {
                const float* t = tab;
                for (int g = 0; g < p.G; ++g, t += p.scan_stride) {
#pragma unroll
                    for (int q = 0; q < ND; ++q) {
                        float e = 0.f;
#pragma unroll
                        for (int k = 1; k < AP; ++k) {
                            e = fmaf(rr[q][k], t[2 * (k - 1)], e);
                            e = fmaf(ri[q][k], t[2 * (k - 1) + 1], e);
                        }
                        track_first(L[q], e, g);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < ND; ++q) {
                if (!valid[q]) continue;
                // merge: left indices are all smaller than right indices, so a tie goes to the left
                const bool left = L[q].best >= Rt[q].best;
                const float best = left ? L[q].best : Rt[q].best;
                const float second = left ? fmaxf(Rt[q].best, L[q].second) : fmaxf(L[q].best, Rt[q].second);
                const int bi = left ? L[q].idx : Rt[q].idx;
                const float pbest = rr[q][0] + 2.f * best;
                uint8_t flags = 0;
                if (2.f * (best - second) <= p.tie_eps * fabsf(pbest)) flags |= RS_FLAG_TIE;
                if (p.method == RS_METHOD_MUSIC) {
                    const float full = (float)M * rr[q][0];
                    if (full - pbest <= 1e-4f * full) flags |= RS_FLAG_GUARD;
                }
                const int live = emit(p, seg, base + q * ANG_THREADS + threadIdx.x, o[q], mult[q], bi, p.grid_deg[bi], yv[q], flags);
                if (ls_partials != nullptr) {      // every antenna's detection of the cell adds the same row
                    const double c = grid_cs[2 * bi], sn = grid_cs[2 * bi + 1], y = (double)yv[q], w = (double)live;
                    acc_ls[0] += w * c * c; acc_ls[1] += w * sn * sn; acc_ls[2] += w * c * sn;
                    acc_ls[3] += w * y * c; acc_ls[4] += w * y * sn; acc_ls[5] += w * y * y; acc_ls[6] += w;
                }
            }
        }
    }
    if (ls_partials == nullptr) return;
    // deterministic block reduction of the seven sums -> ls_partials[seg][0..6]
#pragma unroll
    for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int off = 16; off; off >>= 1) acc_ls[q] += __shfl_xor_sync(0xffffffffu, acc_ls[q], off);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 7; ++q) red[wid][q] = acc_ls[q];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0;
        for (int w = 0; w < ANG_THREADS / 32; ++w) t += red[w][threadIdx.x];
        ls_partials[(size_t)seg * 8 + threadIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core scan for 5 <= A <= 16 on a grid that is symmetric about 0.
//
// The lag-form scan is a dense contraction: for the grid pair (g, G-1-g)
//     E[cell][g] = sum_k Re R_k cos(k phi_g),   O[cell][g] = sum_k Im R_k sin(k phi_g),   P(+-theta) = R_0 + 2 (E +- O),
// i.e. [cells x lags] . [lags x pairs] with AP - 1 lags, zero padded to AP.  It runs on the tensor cores through
// mma.sync.m16n8k16 with fp16 operands, fp32 accumulation and a two-way split of both operands (x = hi + lo, both fp16;
// the lags are divided by R_0 first, so everything lives in [-1, 1]): the products hi*hi + lo*hi + hi*lo keep ~2^-21
// relative accuracy per product, the same order as an fp32 FMA chain.  The split is packed along K: one k-step holds
// eight lags as [hi(8) | lo(8)], so  A = [a_hi | a_lo], B = [b_hi ; b_hi]  gives hi*hi + lo*hi in ONE instruction and
// A = [a_hi | 0], B = [b_lo ; 0]  adds hi*lo -- 2 instructions per matrix, k-step and 8 pairs x 16 cells (the issue
// limit of mma.sync, 0.46 /clk/SM, is what the tensor part costs).  The B fragments (cos / sin tables) are split and
// packed on the host.  A warp carries 32 cells (two 16-row tiles); lane L computes the lags of cell L, publishes them
// through shared memory, every lane builds the A fragments of its rows, and then tracks (best, runner-up, pair index)
// for its accumulator rows.  The pair maximum is E + |O| and the pair minimum E - |O|; the scan tracks the maxima only
// (6 ALU operations per pair and cell); which side won is read off the sign of O once, after the scan, where the
// losing side also enters the runner-up (|O| ~ 0 means the two sides tie, and a tie is flagged RS_FLAG_TIE anyway).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// the k = 8 form: its A fragment is the first half (a0, a1) of the k = 16 fragment, its B fragment one word
__device__ __forceinline__ void mma_f16_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
// (x, y) -> packed fp16 pair of the leading parts (x in the low half) and of the remainders
__device__ __forceinline__ void split_f16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Only the pair maxima are tracked in the scan: the minimum E - |O| of a pair can be the runner-up only for the winning
// pair itself (elsewhere its own maximum is a better candidate), and that one is added after the scan from the odd part
// of the winner.
__device__ __forceinline__ void track_max(Track& t, float hi, int pair) {
    const float m = fminf(t.best, hi);
    const bool up = hi > t.best;
    t.best = fmaxf(t.best, hi);
    t.second = fmaxf(t.second, m);
    t.idx = up ? pair : t.idx;
}

template <int AP, int MINB, bool POWER = false>
__global__ void __launch_bounds__(ANG_THREADS, MINB) angles_mma_kernel(AngleArgs p, const uint32_t* __restrict__ mma_table,
                                                                  int ntiles, const double* __restrict__ grid_cs,
                                                                  double* __restrict__ ls_partials) {
    constexpr int K = AP, KS = AP / 8, LSTRIDE = 2 * K + 8, TILE_WORDS = KS * 192;
    extern __shared__ uint32_t smw[];
    // per tile: [KS][32 lanes] (cos_hi, cos_hi, sin_hi, sin_hi), then [KS][32 lanes] (cos_lo, sin_lo) -- every B operand of
    // the loop below is an aligned register pair / single register of one LDS.128 or LDS.64 result
    uint32_t* tabs = smw;
    float* Lx = reinterpret_cast<float*>(tabs + (size_t)ntiles * TILE_WORDS);   // [warps][32 cells][LSTRIDE]: E lags / R_0, then O lags / R_0
    __shared__ double red[ANG_THREADS / 32][8];
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    const bool seg_clean = p.det_nnear != nullptr && p.det_nnear[seg] == 0;
    double acc_ls[7] = {0, 0, 0, 0, 0, 0, 0};
    if (n > 0) {
        for (int i = threadIdx.x; i < ntiles * KS * 32; i += blockDim.x) {
            const int j = i / (KS * 32), sl = i - j * (KS * 32);          // mma_table: [ntiles][cos, sin][KS][32][hi, lo]
            const uint2 c = __ldg(reinterpret_cast<const uint2*>(mma_table) + (size_t)(2 * j) * KS * 32 + sl);
            const uint2 sn = __ldg(reinterpret_cast<const uint2*>(mma_table) + (size_t)(2 * j + 1) * KS * 32 + sl);
            uint32_t* tile = tabs + (size_t)j * TILE_WORDS;
            reinterpret_cast<uint4*>(tile)[sl] = make_uint4(c.x, c.x, sn.x, sn.x);
            reinterpret_cast<uint2*>(tile + KS * 128)[sl] = make_uint2(c.y, sn.y);
        }
        __syncthreads();
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int gq = lane >> 2, tq = lane & 3;
        float* Lw = Lx + (size_t)wid * 32 * LSTRIDE;
        const int f = seg / p.nseg_per_frame;
        const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
        const int M = p.A, G = p.G;
        const int half = G / 2, odd = G & 1, last_pair = (G + 1) / 2 - 1;
        const float NEG = -3.0e38f;
        for (int base = 0; base < n; base += ANG_THREADS) {
            const int i = base + threadIdx.x;
            const bool valid = i < n;
            const uint32_t ld = valid ? p.det_lead[(size_t)seg * p.seg_cap + i] : (1u << 16);
            const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
            const int mult = (int)(ld >> 16);
            float rr0 = 0.f, yv = 0.f;
            {
                float2 s[AP];
                if (valid) {
                    int a, r, d;
                    rs_split_key(p.det_key[o], a, r, d);
                    const float2* cell = frame + (size_t)r * M * p.D + d;
#pragma unroll
                    for (int m = 0; m < AP; ++m) s[m] = (m < M) ? __ldg(cell + (size_t)m * p.D) : make_float2(0.f, 0.f);
                    yv = atan2f(s[1].y * s[0].x - s[1].x * s[0].y, s[1].x * s[0].x + s[1].y * s[0].y);
                    if (POWER) emit_power<AP>(p, o, mult, s);      // own instantiation: the extra registers cost the plain scan 2 %
                } else {
#pragma unroll
                    for (int m = 0; m < AP; ++m) s[m] = make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int m = 0; m < AP; ++m) rr0 = fmaf(s[m].x, s[m].x, fmaf(s[m].y, s[m].y, rr0));
                const float inv = rr0 > 0.f ? 1.f / rr0 : 0.f;       // |R_k| <= R_0: the normalised lags lie in [-1, 1]
                __syncwarp();                       // the previous pass no longer reads this warp's rows
                float* row = Lw + lane * LSTRIDE;
#pragma unroll
                for (int k = 1; k < AP; ++k) {
                    float xr = 0.f, xi = 0.f;
#pragma unroll
                    for (int m = 0; m + k < AP; ++m) {
                        xr = fmaf(s[m + k].x, s[m].x, xr);
                        xr = fmaf(s[m + k].y, s[m].y, xr);
                        xi = fmaf(s[m + k].y, s[m].x, xi);
                        xi = fmaf(-s[m + k].x, s[m].y, xi);
                    }
                    row[k - 1] = xr * inv;
                    row[K + k - 1] = xi * inv;
                }
                row[K - 1] = 0.f;
                row[2 * K - 1] = 0.f;
                __syncwarp();
            }
            // A fragments of both 16-cell tiles: rows gq, gq + 8 of tile t; columns 2 tq, 2 tq + 1 of k-step s
            uint32_t aE[2][KS][4], aO[2][KS][4];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float* r0 = Lw + (16 * t + gq) * LSTRIDE + 2 * tq;
                const float* r1 = r0 + 8 * LSTRIDE;
#pragma unroll
                for (int s_ = 0; s_ < KS; ++s_) {
                    const float2 e0 = *reinterpret_cast<const float2*>(r0 + 8 * s_), e1 = *reinterpret_cast<const float2*>(r1 + 8 * s_);
                    const float2 o0 = *reinterpret_cast<const float2*>(r0 + K + 8 * s_), o1 = *reinterpret_cast<const float2*>(r1 + K + 8 * s_);
                    split_f16x2(e0.x, e0.y, aE[t][s_][0], aE[t][s_][2]);
                    split_f16x2(e1.x, e1.y, aE[t][s_][1], aE[t][s_][3]);
                    split_f16x2(o0.x, o0.y, aO[t][s_][0], aO[t][s_][2]);
                    split_f16x2(o1.x, o1.y, aO[t][s_][1], aO[t][s_][3]);
                }
            }
            Track tr[2][2] = {{Track{NEG, NEG, 0}, Track{NEG, NEG, 0}}, {Track{NEG, NEG, 0}, Track{NEG, NEG, 0}}};
            const uint32_t* tb = tabs + 4 * lane;
            // one tile = 8 grid pairs x 32 cells; only the last tile can hold pairs beyond the grid, so it is peeled
            auto scan_tile = [&](const int j, auto masked_c) {
                constexpr bool MASKED = decltype(masked_c)::value;
                uint4 bh[KS];                                        // (cos_hi, cos_hi, sin_hi, sin_hi) of this lane's rows 2 tq, 2 tq + 1
                uint2 bl[KS];                                        // (cos_lo, sin_lo)
#pragma unroll
                for (int s_ = 0; s_ < KS; ++s_) {
                    bh[s_] = *reinterpret_cast<const uint4*>(tb + s_ * 128);
                    bl[s_] = *reinterpret_cast<const uint2*>(tb + KS * 128 - 2 * lane + s_ * 64);
                }
                const int pb = 8 * j + 2 * tq;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    float cE[4] = {0.f, 0.f, 0.f, 0.f}, cO[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int s_ = 0; s_ < KS; ++s_) {
                        mma_f16(cE, aE[t][s_][0], aE[t][s_][1], aE[t][s_][2], aE[t][s_][3], bh[s_].x, bh[s_].y);   // hi hi + lo hi
                        mma_f16(cO, aO[t][s_][0], aO[t][s_][1], aO[t][s_][2], aO[t][s_][3], bh[s_].z, bh[s_].w);
                        mma_f16_k8(cE, aE[t][s_][0], aE[t][s_][1], bl[s_].x);                                      // hi lo
                        mma_f16_k8(cO, aO[t][s_][0], aO[t][s_][1], bl[s_].y);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float hi = cE[q] + fabsf(cO[q]);
                        const int pair = pb + (q & 1);
                        if (MASKED && pair > last_pair) hi = NEG;
                        track_max(tr[t][q >> 1], hi, pair);
                    }
                }
            };
            for (int j = 0; j < ntiles - 1; ++j, tb += TILE_WORDS) scan_tile(j, std::false_type{});
            scan_tile(ntiles - 1, std::true_type{});
            // The four lanes of a quad hold different columns of the same four rows (t, h).  Transposing reduction: lanes
            // tq and tq ^ 1 split the tiles (the even lane keeps t = 0 and hands its t = 1 trackers over), then lanes tq
            // and tq ^ 2 split h: lane tq ends with the merged tracker of row (t, h) = (tq & 1, tq >> 1) -- 3 merges
            // and 9 shuffles per lane instead of 8 and 24.
            auto merge = [](Track& k, float ob, float os, int oi) {
                k.second = fmaxf(fmaxf(k.second, os), fminf(k.best, ob));
                if (ob > k.best || (ob == k.best && oi < k.idx)) { k.best = ob; k.idx = oi; }
            };
            Track kh[2];
            {
                const bool up1 = (tq & 1) != 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    Track keep = up1 ? tr[1][h] : tr[0][h];
                    const Track send = up1 ? tr[0][h] : tr[1][h];
                    const float ob = __shfl_xor_sync(0xffffffffu, send.best, 1);
                    const float os = __shfl_xor_sync(0xffffffffu, send.second, 1);
                    const int oi = __shfl_xor_sync(0xffffffffu, send.idx, 1);
                    merge(keep, ob, os, oi);
                    kh[h] = keep;
                }
            }
            const bool up2 = (tq & 2) != 0;
            Track mine = up2 ? kh[1] : kh[0];
            {
                const Track send = up2 ? kh[0] : kh[1];
                const float ob = __shfl_xor_sync(0xffffffffu, send.best, 2);
                const float os = __shfl_xor_sync(0xffffffffu, send.second, 2);
                const int oi = __shfl_xor_sync(0xffffffffu, send.idx, 2);
                merge(mine, ob, os, oi);
            }
            // row (t, h, gq) is cell 16 t + 8 h + gq of the warp: fetch it from lane 4 gq + t + 2 h
            const int src = 4 * (lane & 7) + (lane >> 4) + 2 * ((lane >> 3) & 1);
            float my_best = __shfl_sync(0xffffffffu, mine.best, src);
            float my_second = __shfl_sync(0xffffffffu, mine.second, src);
            const int my_pair = __shfl_sync(0xffffffffu, mine.idx, src);
            if (valid) {
                // which side of the pair won: sign of the odd part, re-evaluated with the fp32 table
                const float* row = Lw + lane * LSTRIDE + K;
                const float* trow = p.scan_table + (size_t)my_pair * p.scan_stride;
                float od = 0.f;
#pragma unroll
                for (int k = 1; k < AP; ++k) od = fmaf(row[k - 1], __ldg(trow + 2 * (k - 1) + 1), od);
                const bool middle = odd && my_pair == half;             // the middle angle has no partner
                const int bi = middle ? half : (od >= 0.f ? my_pair : G - 1 - my_pair);
                if (!middle) my_second = fmaxf(my_second, my_best - 2.f * fabsf(od));   // the losing side of the winning pair
                // my_best, my_second are (P - R_0) / (2 R_0) of the best and second-best grid point
                const float pnorm = 1.f + 2.f * my_best;
                uint8_t flags = 0;
                if (2.f * (my_best - my_second) <= p.tie_eps * fabsf(pnorm)) flags |= RS_FLAG_TIE;
                if (p.method == RS_METHOD_MUSIC) {
                    const float full = (float)M;
                    if (full - pnorm <= 1e-4f * full) flags |= RS_FLAG_GUARD;
                }
                const int live = emit(p, seg, i, o, mult, bi, p.grid_deg[bi], yv, flags, seg_clean);
                if (ls_partials != nullptr) {
                    const double c = grid_cs[2 * bi], sn = grid_cs[2 * bi + 1], y = (double)yv, w = (double)live;
                    acc_ls[0] += w * c * c; acc_ls[1] += w * sn * sn; acc_ls[2] += w * c * sn;
                    acc_ls[3] += w * y * c; acc_ls[4] += w * y * sn; acc_ls[5] += w * y * y; acc_ls[6] += w;
                }
            }
        }
    }
    if (ls_partials == nullptr) return;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int off = 16; off; off >>= 1) acc_ls[q] += __shfl_xor_sync(0xffffffffu, acc_ls[q], off);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 7; ++q) red[wid][q] = acc_ls[q];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0;
        for (int w = 0; w < ANG_THREADS / 32; ++w) t += red[w][threadIdx.x];
        ls_partials[(size_t)seg * 8 + threadIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 scan (the default for 5..16 antennas on a symmetric grid; RS_ANGLES_TC=0 selects the mma.sync kernel above): the
// same contraction on the 5th-generation tensor cores.
//
// One CTA = 128 threads = 128 cells = the 128 TMEM lanes of a UMMA tile.  Every thread gathers its cell's snapshot,
// forms the normalised lags and writes its ROW of the A operand ([Re R_k] for the even part, [Im R_k] for the odd
// part, fp16 hi / lo split packed along K exactly as in the mma.sync kernel) into shared memory in the canonical
// K-major no-swizzle UMMA layout (8 x 16-byte core matrices).  One elected thread issues
//     tcgen05.mma.cta_group::1.kind::f16   D[128 x 32] (+)= A[128 x 16] . B[16 x 32]
// per 16-wide K chunk for the cos table (E) and the sin table (O) of one "job" of 32 grid pairs into one of two 64-column TMEM
// buffers, commits to that buffer's mbarrier, and after the wait each thread reads ITS OWN cell's 32 + 32 accumulators with
// tcgen05.ld.32x32b -- a whole row per thread, so the pair maximum E + |O| is tracked without any cross-lane merge
// (the mma.sync fragment layout spreads a row over a quad: 3 merges and 12 shuffles per cell there).  Two jobs are in
// flight: the tensor core fills one buffer while the threads track the other, and the next tile's snapshot loads are
// issued before the tracking starts, so their latency is hidden too.  The B operand (cos / sin tables, split hi / lo,
// already in UMMA layout) is built on the host (tables.scan_tc_table).
// 128 TMEM columns per CTA: four persistent CTAs per SM keep 16 warps resident; they walk the segments (or, with 16
// antennas, the segment pairs of a tile: PAIR below).  The tracking is two-level (group maxima + stash, see the tile loop).
// ---------------------------------------------------------------------------------------------
namespace tc5 {

constexpr int NPH = 32;                    // grid pairs per job (UMMA N); two jobs in flight in 2 x 64 TMEM columns
constexpr int PAIR_CELLS = 2048;           // pair mode: cells of a detection tile (16 range x 128 Doppler bins)
constexpr int PAIR_WS_BYTES = PAIR_CELLS * (8 + 2);              // per CTA: result table + queue of cell indices
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, k) of a [rows x 16] fp16 chunk sits at (row / 8) * 256 + (k / 8) * 128 + (row % 8) * 16
// + (k % 8) * 2: leading (K) byte offset 128, stride (M / N) byte offset 256, descriptor version 1 (Blackwell)
__device__ __forceinline__ unsigned long long smem_desc(uint32_t addr) {
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | ((unsigned long long)(128 >> 4) << 16) |
           ((unsigned long long)(256 >> 4) << 32) | (1ull << 46);
}
// kind::f16: D fp32 (bit 4), A / B fp16 K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NPH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, unsigned long long da, unsigned long long db, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ unsigned long long pk2f(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

}  // namespace tc5

template <int AP, bool PAIR>
__global__ void __launch_bounds__(ANG_THREADS, 4) angles_tc5_kernel(AngleArgs p, const uint4* __restrict__ tc_table, int njobs,
                                                                 const double* __restrict__ grid_cs,
                                                                 double* __restrict__ ls_partials, int nsegs,
                                                                 unsigned char* __restrict__ ws, int pair_tr, int pair_td) {
    using namespace tc5;
    constexpr int K = AP;                              // lag slots per part (AP - 1 lags + one constant)
    constexpr int KC = AP == 8 ? 2 : 3;                // 16-wide K chunks per matrix: AP 8: [hi|lo][hi|0]; AP 16: [hi][lo][hi]
    constexpr int B_CHUNK = NPH * 32;                  // bytes of one [32 x 16] fp16 chunk
    constexpr int A_CHUNK = 128 * 32;                  // bytes of one [128 x 16] fp16 chunk
    constexpr int A_CHUNKS = 2;                        // stored A chunks per matrix (AP 16 re-uses [hi] for the third)
    constexpr int LSTRIDE = K + 4;                     // fp32 rows read as float4: stride 12 / 20 words is conflict free
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned char* Bt = smraw;                                                   // [njobs][E, O][KC][B_CHUNK]
    unsigned char* At = Bt + (size_t)njobs * 2 * KC * B_CHUNK;                   // [E, O][A_CHUNKS][A_CHUNK]
    float* Lx = reinterpret_cast<float*>(At + 2 * A_CHUNKS * A_CHUNK);           // [128][LSTRIDE] odd lags / R_0 (fp32)
    float* St = Lx + 128 * LSTRIDE;                                              // [pairs][LSTRIDE] sin(k phi_pair), fp32
    double2* Gc = reinterpret_cast<double2*>(St + (((p.G + 1) / 2) * LSTRIDE + 3) / 4 * 4);   // [G] cos / sin of the grid angle
    float* Gd = reinterpret_cast<float*>(Gc + p.G);                              // [G] grid angle in degrees
    __shared__ unsigned long long mbar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ double red[ANG_THREADS / 32][8];
    __shared__ uint32_t cellmap[PAIR_CELLS / 32];      // pair mode: cells of the tile flagged on either antenna octet
    __shared__ int wsum[ANG_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int M = p.A, G = p.G;
    const int half = G / 2, odd = G & 1, last_pair = (G + 1) / 2 - 1;
    // Persistent: a CTA allocates its TMEM columns, initialises its barriers and stages the tables ONCE and then walks the
    // segments blockIdx.x, blockIdx.x + gridDim.x, ... (a segment is ~9 tiles: the set-up was 13 % of the stall samples).
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const int n16 = njobs * 2 * KC * B_CHUNK / 16;
        uint4* dst = reinterpret_cast<uint4*>(Bt);
        for (int i = tid; i < n16; i += ANG_THREADS) dst[i] = __ldg(tc_table + i);
        // the fp32 sin rows of the scan table: the sign of the winner's odd part is re-evaluated from them after the scan
        for (int i = tid; i < (last_pair + 1) * LSTRIDE; i += ANG_THREADS) {
            const int pr = i / LSTRIDE, k = i - pr * LSTRIDE;
            St[i] = k < AP - 1 ? __ldg(p.scan_table + (size_t)pr * p.scan_stride + 2 * k + 1) : 0.f;
        }
#pragma unroll
        for (int k = AP - 1; k < LSTRIDE; ++k) Lx[tid * LSTRIDE + k] = 0.f;
        for (int i = tid; i < G; i += ANG_THREADS) {
            Gc[i] = grid_cs != nullptr ? make_double2(grid_cs[2 * i], grid_cs[2 * i + 1]) : make_double2(0.0, 0.0);
            Gd[i] = p.grid_deg[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t trow = tmem_base + ((uint32_t)(wid * 32) << 16);          // this warp's 32 TMEM lanes
    uint32_t phase[2] = {0u, 0u};
    // PAIR (16 antennas, scratch in ws): the detection tiles are per antenna OCTET, so a cell flagged on both octets of
    // its (range, Doppler) tile leads a group in two segments -- 26 % of all leaders at the benchmark density -- with the
    // same 16-channel snapshot in both.  A unit is then the segment PAIR of one tile:
    //   mark      both leader lists set the bit of their cells in a 2048-bit map (shared memory);
    //   enumerate the set bits, in cell order, become the queue of the tile loop (ordered compaction);
    //   scan      every distinct cell once; (grid index, TIE / GUARD bits, phase) go to a table indexed by the cell
    //             (per-CTA scratch in global memory, 20 KB that stay in L2);
    //   scatter   every leader of both segments copies its cell's row to its group (emit) and adds to the segment's
    //             velocity sums, thread t taking the leaders [t c, t c + c): fixed order, fixed sums.
    // (A first version scanned segment 0, then the cells of segment 1 not seen yet: fewer instructions but 2x the DRAM
    // traffic and 13 % slower -- by the time the second pass gathered its cells the tile had left L2.)
    uint2* wres = reinterpret_cast<uint2*>(ws + (size_t)blockIdx.x * PAIR_WS_BYTES);       // [cells] result table
    unsigned short* wqc = reinterpret_cast<unsigned short*>(wres + PAIR_CELLS);            // [cells] queue of cell indices
    const int pair_dmask = pair_td - 1, pair_dshift = 31 - __clz(pair_td);
    const int nunits = PAIR ? nsegs / 2 : nsegs;
    // leaders [t c, t c + c) of a segment through fn(index, leader word, key, row): four at a time, their words, then
    // their keys, then (WITH_ROW) their rows of the result table requested together -- three load latencies per four
    // leaders; fn itself runs in a rolled loop (it may inline the large emit)
    auto for_leaders = [&](int sg, auto with_row, auto fn) {
        constexpr bool WITH_ROW = decltype(with_row)::value;
        const int nl = p.det_nlead[sg];
        const size_t sb = (size_t)sg * p.seg_cap;
        const int c = (nl + ANG_THREADS - 1) / ANG_THREADS, i0 = tid * c;
#pragma unroll 1
        for (int j0 = 0; j0 < c; j0 += 4) {
            uint32_t lv[4], kv[4];
            uint2 rv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) lv[j] = (j0 + j < c && i0 + j0 + j < nl) ? __ldg(p.det_lead + sb + i0 + j0 + j) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = (j0 + j < c && i0 + j0 + j < nl) ? __ldg(p.det_key + sb + (lv[j] & 0xFFFFu)) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rv[j] = make_uint2(0u, 0u);
                if (WITH_ROW && j0 + j < c && i0 + j0 + j < nl) {
                    int a_, r_, d_;
                    rs_split_key(kv[j], a_, r_, d_);
                    rv[j] = __ldcg(wres + (((r_ & (pair_tr - 1)) << pair_dshift) | (d_ & pair_dmask)));
                }
            }
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                uint32_t l = lv[0], ky = kv[0];
                uint2 row = rv[0];
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    if (j == q) { l = lv[q]; ky = kv[q]; row = rv[q]; }
                if (j0 + j < c && i0 + j0 + j < nl) fn(i0 + j0 + j, l, ky, row);
            }
        }
    };
    auto cell_of = [&](uint32_t key) -> int {
        int a_, r_, d_;
        rs_split_key(key, a_, r_, d_);
        return ((r_ & (pair_tr - 1)) << pair_dshift) | (d_ & pair_dmask);
    };
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int seg = PAIR ? 2 * unit : unit;
    int n = 0, tile_r0 = 0, tile_d0 = 0;
    if constexpr (PAIR) {
        const int tile = (seg % p.nseg_per_frame) >> 1, ntd = p.D / pair_td;
        tile_r0 = (tile / ntd) * pair_tr;
        tile_d0 = (tile % ntd) * pair_td;
        if (tid < PAIR_CELLS / 32) cellmap[tid] = 0u;
        __syncthreads();
        for (int sg = seg; sg < seg + 2; ++sg)
            for_leaders(sg, std::false_type{}, [&](int, uint32_t, uint32_t key, uint2) {
                const int cell = cell_of(key);
                atomicOr(&cellmap[cell >> 5], 1u << (cell & 31));
            });
        __syncthreads();
        // thread t enumerates bits [16 t, 16 t + 16)
        const uint32_t bits = (cellmap[tid >> 1] >> (16 * (tid & 1))) & 0xFFFFu;
        const int mine = __popc(bits);
        int inc = mine;
#pragma unroll
        for (int o_ = 1; o_ < 32; o_ <<= 1) {
            const int t_ = __shfl_up_sync(0xffffffffu, inc, o_);
            if (lane >= o_) inc += t_;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int off = inc - mine;
#pragma unroll
        for (int w = 0; w < ANG_THREADS / 32; ++w) {
            if (w < wid) off += wsum[w];
            n += wsum[w];
        }
        for (uint32_t b = bits; b; b &= b - 1) __stcg(wqc + off++, (unsigned short)(16 * tid + __ffs(b) - 1));
        __syncthreads();                                             // the queue is complete
    } else {
        n = p.det_nlead[seg];
    }
    const bool seg_clean = p.det_nnear != nullptr && p.det_nnear[seg] == 0;
    double acc_ls[7] = {0, 0, 0, 0, 0, 0, 0};
    if (n > 0) {
        const int f = seg / p.nseg_per_frame;
        const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
        const float NEG = -3.0e38f;
        // this thread's row of a chunk: (tid / 8) * 256 + (tid % 8) * 16, k 0..7 at +0, k 8..15 at +128
        const uint32_t arow = (uint32_t)((tid >> 3) * 256 + (tid & 7) * 16);
        float* row = Lx + tid * LSTRIDE;
        const uint32_t a_base = s32(At), b_base = s32(Bt);
        // job j = grid pairs [32 j, 32 j + 32) of the current tile -> TMEM buffer j & 1 (E at +0, O at +32 of its 64 columns)
        auto issue_job = [&](int jb) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d0 = tmem_base + (uint32_t)((jb & 1) * 2 * NPH);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const int ac = (AP == 16 && c == 2) ? 0 : c;                     // AP 16: the third chunk multiplies hi again
                    umma(d0 + (uint32_t)(part * NPH), smem_desc(a_base + (uint32_t)((part * A_CHUNKS + ac) * A_CHUNK)),
                         smem_desc(b_base + (uint32_t)(((jb * 2 + part) * KC + c) * B_CHUNK)), c > 0 ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar[jb & 1])) : "memory");
        };
        // The chain leader -> key -> snapshot is three dependent global loads.  It is software pipelined over the tiles:
        // while tile t is scanned, the snapshot of tile t + 1 is in flight (its key arrived during tile t - 1) and so is the
        // leader of tile t + 2; its key is requested after the scan.
        float2 s[AP];
        const size_t segbase = (size_t)seg * p.seg_cap;
        // PAIR: slot i is the i-th queued cell; its "leader word" is the cell index, its "key" the (range, Doppler) key
        auto lead_of = [&](int i) -> uint32_t {
            if constexpr (PAIR) return i < n ? (uint32_t)__ldcg(wqc + i) : 0u;
            else return i < n ? p.det_lead[segbase + i] : (1u << 16);
        };
        auto key_of = [&](int i, uint32_t l) -> uint32_t {
            if constexpr (PAIR) return ((uint32_t)(tile_r0 + (int)(l >> pair_dshift)) << 12) | (uint32_t)(tile_d0 + (int)(l & pair_dmask));
            else return i < n ? p.det_key[segbase + (l & 0xFFFFu)] : 0u;
        };
        auto snapshot = [&](int i, uint32_t key) {
            if (i < n) {
                int a, r, d;
                rs_split_key(key, a, r, d);
                const float2* cell = frame + (size_t)r * M * p.D + d;
#pragma unroll
                for (int m = 0; m < AP; ++m) s[m] = (m < M) ? __ldg(cell + (size_t)m * p.D) : make_float2(0.f, 0.f);
            } else {
#pragma unroll
                for (int m = 0; m < AP; ++m) s[m] = make_float2(0.f, 0.f);
            }
        };
        // (the key of tile t + 2 is requested at the START of tile t's scan, from a leader that arrived a tile earlier: a
        // load issued at the end of the iteration made the first scoreboard wait of the next one a full L2 latency)
        uint32_t ld = lead_of(tid), ld1 = lead_of(ANG_THREADS + tid), ld2 = lead_of(2 * ANG_THREADS + tid), ld3 = 1u << 16;
        uint32_t key1 = 0u, key2 = 0u;
        snapshot(tid, key_of(tid, ld));
        key1 = key_of(ANG_THREADS + tid, ld1);
        for (int base = 0; base < n; base += ANG_THREADS) {
            const int i = base + tid;
            const bool valid = i < n;
            const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
            const int mult = (int)(ld >> 16);
            float yv = 0.f;
            {
                float rr0 = 0.f;
                if (valid) {
                    yv = atan2f(s[1].y * s[0].x - s[1].x * s[0].y, s[1].x * s[0].x + s[1].y * s[0].y);
                    if constexpr (!PAIR) emit_power<AP>(p, o, mult, s);
                }
#pragma unroll
                for (int m = 0; m < AP; ++m) rr0 = fmaf(s[m].x, s[m].x, fmaf(s[m].y, s[m].y, rr0));
                const float inv = rr0 > 0.f ? 1.f / rr0 : 0.f;       // |R_k| <= R_0: the normalised lags lie in [-1, 1]
                float le[K], lo_[K];
                // lags in packed f32x2 arithmetic: (xr, xi) += s[m+k] * (s[m].x, s[m].x), then += (s[m+k].y, -s[m+k].x) *
                // (s[m].y, s[m].y) -- per component the same fused multiply-adds in the same order as the scalar form (the
                // half swap, the sign and the broadcast are operand modifiers of FFMA2): half the instructions
#pragma unroll
                for (int k = 1; k < AP; ++k) {
                    unsigned long long acc = tc5::pk2f(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m + k < AP; ++m) {
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(tc5::pk2f(s[m + k].x, s[m + k].y)), "l"(tc5::pk2f(s[m].x, s[m].x)));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(tc5::pk2f(s[m + k].y, -s[m + k].x)), "l"(tc5::pk2f(s[m].y, s[m].y)));
                    }
                    float xr, xi;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(xr), "=f"(xi) : "l"(acc));
                    le[k - 1] = xr * inv;
                    lo_[k - 1] = xi * inv;
                    row[k - 1] = xi * inv;                            // the odd lags are needed again after the scan
                }
                le[K - 1] = 1.f;     // constant-one slot: the table holds 0 there for grid pairs, -16384 for the padding columns
                lo_[K - 1] = 0.f;
                // fp16 hi / lo split, packed along K, into this thread's rows of the E and O operands
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t hi[K / 2], lw[K / 2];
#pragma unroll
                    for (int j = 0; j < K / 2; ++j) {
                        if (part == 0) split_f16x2(le[2 * j], le[2 * j + 1], hi[j], lw[j]);
                        else split_f16x2(lo_[2 * j], lo_[2 * j + 1], hi[j], lw[j]);
                    }
                    unsigned char* A0 = At + part * A_CHUNKS * A_CHUNK + arow;
                    if (AP == 8) {
                        // chunk 0 = [hi(8) | lo(8)], chunk 1 = [hi(8) | 0]
                        *reinterpret_cast<uint4*>(A0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(A0 + 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        *reinterpret_cast<uint4*>(A0 + A_CHUNK) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(A0 + A_CHUNK + 128) = make_uint4(0u, 0u, 0u, 0u);
                    } else {
                        // chunk 0 = hi(16), chunk 1 = lo(16)
                        constexpr int H = K / 2;
                        *reinterpret_cast<uint4*>(A0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(A0 + 128) = make_uint4(hi[4 % H], hi[5 % H], hi[6 % H], hi[7 % H]);
                        *reinterpret_cast<uint4*>(A0 + A_CHUNK) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        *reinterpret_cast<uint4*>(A0 + A_CHUNK + 128) = make_uint4(lw[4 % H], lw[5 % H], lw[6 % H], lw[7 % H]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes of A -> the tensor core's async reads
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                             // A complete; both TMEM buffers drained by the last tile
            if (tid == 0) {
                issue_job(0);
                if (njobs > 1) issue_job(1);
            }
            // the next tile's snapshot is requested now and consumed after this tile's scan: its latency hides under the tracking
            snapshot(base + ANG_THREADS + tid, key1);
            key2 = key_of(base + 2 * ANG_THREADS + tid, ld2);
            ld3 = lead_of(base + 3 * ANG_THREADS + tid);
            // Two-level tracking.  The scan is bound by instruction issue, and a (best, runner-up, index) tracker costs 6
            // operations per grid pair.  Here a group of 8 pair maxima is reduced to its maximum with four 3-input maxima
            // and only THAT enters the tracker (best / runner-up over the group maxima, 5 operations per group); the group
            // that raises the best leaves its eight values in a stash (8 predicated moves).  After the scan the winner's
            // index inside its group and the runner-up inside the group come from the stash, once per cell: every value
            // outside the winning group is bounded by its own group maximum, so
            //     runner-up = max(second largest group maximum, second largest value of the winning group)
            // exactly as the flat tracker computes it.  ~3.1 instead of ~9 instructions per pair and cell.
            // Columns beyond the last grid pair need no mask: the table puts -TC_PAD on the constant-one K slot there.
            float g_best = NEG, g_second = NEG;
            int g_idx = 0, g_run = 0;
            float stash[8] = {NEG, NEG, NEG, NEG, NEG, NEG, NEG, NEG};
            for (int jb = 0; jb < njobs; ++jb) {
                const int b = jb & 1;
                {
                    uint32_t ok = 0;
                    asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                                 : "=r"(ok) : "r"(s32(&mbar[b])), "r"(phase[b]) : "memory");
                    if (!ok) {                                           // the clock (a hang becomes a trap) only on the slow path
                        const long long t0 = clock64();
                        while (!ok) {
                            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                                         : "=r"(ok) : "r"(s32(&mbar[b])), "r"(phase[b]) : "memory");
                            if (!ok && clock64() - t0 > 4000000000ll) __trap();
                        }
                    }
                    phase[b] ^= 1u;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tb = trow + (uint32_t)(b * 2 * NPH);
                // TMEM reads (64 B / clk / SM): groups of 8 pairs, the loads of group g + 1 in flight while group g is reduced
                float e[2][8], od[2][8];
                ld8(tb, e[0]);
                ld8(tb + (uint32_t)NPH, od[0]);
#pragma unroll
                for (int g = 0; g < NPH / 8; ++g) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (g + 1 < NPH / 8) {
                        ld8(tb + (uint32_t)(8 * (g + 1)), e[(g + 1) & 1]);
                        ld8(tb + (uint32_t)(NPH + 8 * (g + 1)), od[(g + 1) & 1]);
                    }
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = e[g & 1][j] + fabsf(od[g & 1][j]);
                    const float gm = fmaxf(fmaxf(fmaxf(v[0], v[1]), v[2]),
                                           fmaxf(fmaxf(fmaxf(v[3], v[4]), v[5]), fmaxf(v[6], v[7])));
                    const bool up = gm > g_best;                         // strict: the earliest group keeps exact ties
                    g_second = fmaxf(g_second, fminf(g_best, gm));
                    g_best = fmaxf(g_best, gm);
                    g_idx = up ? g_run : g_idx;
                    ++g_run;
#pragma unroll
                    for (int j = 0; j < 8; ++j) stash[j] = up ? v[j] : stash[j];
                }
                if (jb + 2 < njobs) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncthreads();                                     // every thread has read buffer b
                    if (tid == 0) issue_job(jb + 2);
                }
            }
            // inside the winning group: first index holding the maximum, and the largest of the other seven
            int in_idx = 7;
#pragma unroll
            for (int j = 6; j >= 0; --j) in_idx = (stash[j] == g_best) ? j : in_idx;
            float in_second = NEG;
#pragma unroll
            for (int j = 0; j < 8; ++j) in_second = fmaxf(in_second, j == in_idx ? NEG : stash[j]);
            float my_best = g_best, my_second = fmaxf(g_second, in_second);
            const int my_pair = 8 * g_idx + in_idx;
            if (valid) {
                // which side of the pair won: sign of the odd part, re-evaluated with the fp32 table
                const float4* st4 = reinterpret_cast<const float4*>(St + my_pair * LSTRIDE);
                const float4* rw4 = reinterpret_cast<const float4*>(row);
                float od = 0.f;
#pragma unroll
                for (int q = 0; q < AP / 4; ++q) {                      // slot AP - 1 is zero in both rows
                    const float4 x = rw4[q], t = st4[q];
                    od = fmaf(x.x, t.x, od); od = fmaf(x.y, t.y, od); od = fmaf(x.z, t.z, od); od = fmaf(x.w, t.w, od);
                }
                const bool middle = odd && my_pair == half;             // the middle angle has no partner
                const int bi = middle ? half : (od >= 0.f ? my_pair : G - 1 - my_pair);
                if (!middle) my_second = fmaxf(my_second, my_best - 2.f * fabsf(od));   // the losing side of the winning pair
                const float pnorm = 1.f + 2.f * my_best;
                uint8_t flags = 0;
                if (2.f * (my_best - my_second) <= p.tie_eps * fabsf(pnorm)) flags |= RS_FLAG_TIE;
                if (p.method == RS_METHOD_MUSIC) {
                    const float full = (float)M;
                    if (full - pnorm <= 1e-4f * full) flags |= RS_FLAG_GUARD;
                }
                if constexpr (PAIR) {
                    __stcg(wres + (ld & 0xFFFFu), make_uint2((uint32_t)bi | ((uint32_t)flags << 16), __float_as_uint(yv)));
                } else {
                    const int live = emit<true>(p, seg, i, o, mult, bi, Gd[bi], yv, flags, seg_clean);
                    if (ls_partials != nullptr) {
                        const double2 cs = Gc[bi];
                        const double c = cs.x, sn = cs.y, y = (double)yv, w = (double)live;
                        acc_ls[0] += w * c * c; acc_ls[1] += w * sn * sn; acc_ls[2] += w * c * sn;
                        acc_ls[3] += w * y * c; acc_ls[4] += w * y * sn; acc_ls[5] += w * y * y; acc_ls[6] += w;
                    }
                }
            }
            ld = ld1;
            ld1 = ld2;
            ld2 = ld3;
            key1 = key2;
        }
    }
    auto reduce_sums = [&](int sg) {
#pragma unroll
        for (int q = 0; q < 7; ++q) {
#pragma unroll
            for (int off = 16; off; off >>= 1) acc_ls[q] += __shfl_xor_sync(0xffffffffu, acc_ls[q], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 7; ++q) red[wid][q] = acc_ls[q];
        }
        __syncthreads();
        if (tid < 7) {
            double t = 0;
            for (int w = 0; w < ANG_THREADS / 32; ++w) t += red[w][tid];
            ls_partials[(size_t)sg * 8 + tid] = t;
        }
        __syncthreads();                                             // red is reused
    };
    if constexpr (PAIR) {
        __syncthreads();                                             // the table is complete
        for (int sg = seg; sg < seg + 2; ++sg) {
            const bool clean = p.det_nnear != nullptr && p.det_nnear[sg] == 0;
            const size_t sb = (size_t)sg * p.seg_cap;
#pragma unroll
            for (int q = 0; q < 7; ++q) acc_ls[q] = 0.0;
            for_leaders(sg, std::true_type{}, [&](int i, uint32_t l, uint32_t, uint2 rv) {
                const int bi = (int)(rv.x & 0xFFFFu);
                const float yv = __uint_as_float(rv.y);
                const int live = emit<true>(p, sg, i, sb + (l & 0xFFFFu), (int)(l >> 16), bi, Gd[bi], yv, (uint8_t)(rv.x >> 16), clean);
                if (ls_partials != nullptr) {
                    const double2 cs = Gc[bi];
                    const double c = cs.x, sn = cs.y, y = (double)yv, w = (double)live;
                    acc_ls[0] += w * c * c; acc_ls[1] += w * sn * sn; acc_ls[2] += w * c * sn;
                    acc_ls[3] += w * y * c; acc_ls[4] += w * y * sn; acc_ls[5] += w * y * y; acc_ls[6] += w;
                }
            });
            if (ls_partials != nullptr) reduce_sums(sg);
        }
        __syncthreads();                                             // table, queue and bit map are free again
    } else {
        if (ls_partials != nullptr) reduce_sums(seg);
    }
    }   // units
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ---- any A (used for A > 16): one warp evaluates one cell, lanes over antennas / grid points, snapshot in smem ----
struct CellResult {
    int aidx;
    float adeg, phase;
    uint32_t flags;
};

// Valid on lane 0.  `cell` points at antenna 0 of the cell; antennas are `stride` elements apart.
__device__ __forceinline__ CellResult large_cell(const AngleArgs& p, const float2* __restrict__ cell, size_t stride,
                                                 float2* s, int lane) {
    const int M = p.A;
    CellResult out{-1, 0.f, 0.f, 0u};
    __syncwarp();
    float e = 0.f;
    for (int m = lane; m < M; m += 32) {
        const float2 x = __ldg(cell + (size_t)m * stride);
        s[m] = x;
        e = fmaf(x.x, x.x, fmaf(x.y, x.y, e));
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    __syncwarp();
    out.phase = atan2f(s[1].y * s[0].x - s[1].x * s[0].y, s[1].x * s[0].x + s[1].y * s[0].y);
    if (p.method == RS_METHOD_ESPRIT) {
        // warp-cooperative sums in fp64
        double alpha = 0, gamma = 0, br = 0, bi = 0;
        for (int m = lane; m < M - 1; m += 32) {
            const double xr = s[m].x, xi = s[m].y, yr = s[m + 1].x, yi = s[m + 1].y;
            alpha += xr * xr + xi * xi; gamma += yr * yr + yi * yi;
            br += xr * yr + xi * yi;    bi += xr * yi - xi * yr;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            alpha += __shfl_xor_sync(0xffffffffu, alpha, off); gamma += __shfl_xor_sync(0xffffffffu, gamma, off);
            br += __shfl_xor_sync(0xffffffffu, br, off);       bi += __shfl_xor_sync(0xffffffffu, bi, off);
        }
        const double half = 0.5 * (alpha - gamma);
        const double lam = 0.5 * (alpha + gamma) + sqrt(half * half + br * br + bi * bi);
        double v0r, v0i, v1r, v1i;
        const double na = br * br + bi * bi + (lam - alpha) * (lam - alpha);
        const double nb = (lam - gamma) * (lam - gamma) + br * br + bi * bi;
        if (na >= nb) { v0r = br; v0i = bi; v1r = lam - alpha; v1i = 0; }
        else          { v0r = lam - gamma; v0i = 0; v1r = br; v1i = -bi; }
        double nr = 0, ni = 0;
        for (int m = lane; m < M - 2; m += 32) {
            const double x0r = s[m].x, x0i = s[m].y, x1r = s[m + 1].x, x1i = s[m + 1].y, x2r = s[m + 2].x, x2i = s[m + 2].y;
            const double ur = v0r * x0r - v0i * x0i + v1r * x1r - v1i * x1i;
            const double ui = v0r * x0i + v0i * x0r + v1r * x1i + v1i * x1r;
            const double wr = v0r * x1r - v0i * x1i + v1r * x2r - v1i * x2i;
            const double wi = v0r * x1i + v0i * x1r + v1r * x2i + v1i * x2r;
            nr += ur * wr + ui * wi;
            ni += ur * wi - ui * wr;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            nr += __shfl_xor_sync(0xffffffffu, nr, off);
            ni += __shfl_xor_sync(0xffffffffu, ni, off);
        }
        out.adeg = (float)(asin(atan2(ni, nr) * p.esprit_scale) * (180.0 / 3.14159265358979323846));
        return out;
    }
    float best = -1.f, second = -1.f;
    int bi = 0x7fffffff;
    for (int g = lane; g < p.G; g += 32) {
        float ar = 0.f, ai = 0.f;
        const float2* st = p.steer + g;
        for (int m = 0; m < M; ++m) {
            const float2 w = __ldg(st + (size_t)m * p.G);     // a_g[m]; accumulate conj(a) * s
            const float2 x = s[m];
            ar = fmaf(w.x, x.x, fmaf(w.y, x.y, ar));
            ai = fmaf(w.x, x.y, fmaf(-w.y, x.x, ai));
        }
        const float v = fmaf(ar, ar, ai * ai);
        if (v > best) { second = best; best = v; bi = g; }
        else if (v > second) second = v;
    }
    // warp arg-max with first-index tie break; second = best of the rest
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const float os = __shfl_xor_sync(0xffffffffu, second, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { second = fmaxf(best, os); best = ob; bi = oi; }
        else second = fmaxf(second, ob);
    }
    if ((best - second) <= p.tie_eps * best) out.flags |= RS_FLAG_TIE;
    if (p.method == RS_METHOD_MUSIC) {
        const float full = (float)M * e;
        if (full - best <= 1e-4f * full) out.flags |= RS_FLAG_GUARD;
    }
    out.aidx = bi;
    out.adeg = p.grid_deg[bi < p.G ? bi : 0];
    return out;
}

// warp per detection leader (no workspace)
__global__ void __launch_bounds__(ANG_THREADS) angles_large_kernel(AngleArgs p) {
    extern __shared__ float2 snap[];   // [warps][A]
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    if (n == 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int f = seg / p.nseg_per_frame;
    const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
    for (int i = wid; i < n; i += nw) {
        const uint32_t ld = p.det_lead[(size_t)seg * p.seg_cap + i];
        const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
        int a, r, d;
        rs_split_key(p.det_key[o], a, r, d);
        const CellResult c = large_cell(p, frame + (size_t)r * p.A * p.D + d, (size_t)p.D, snap + wid * p.A, lane);
        if (lane == 0) emit(p, seg, i, o, (int)(ld >> 16), c.aidx, c.adeg, c.phase, (uint8_t)c.flags);
    }
}

// With many antennas a range-Doppler cell is flagged on many of them (17 detections per cell at A = 192), and their
// snapshot -- hence angle and phase -- is the same.  Three passes evaluate every distinct cell ONCE per frame:
// mark the cells that carry a detection, evaluate the marked cells (warp per cell), copy the results to the lists.
__global__ void __launch_bounds__(256) mark_cells_kernel(AngleArgs p, uint8_t* __restrict__ mark) {
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    const int f = seg / p.nseg_per_frame;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t ld = p.det_lead[(size_t)seg * p.seg_cap + i];
        int a, r, d;
        rs_split_key(p.det_key[(size_t)seg * p.seg_cap + (ld & 0xFFFFu)], a, r, d);
        mark[((size_t)f * p.R + r) * p.D + d] = 1;
    }
}

__global__ void __launch_bounds__(ANG_THREADS) eval_cells_kernel(AngleArgs p, const uint8_t* __restrict__ mark,
                                                                  CellResult* __restrict__ cells, long long ncells) {
    extern __shared__ float2 snap[];   // [warps][A]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long per_frame = (long long)p.R * p.D;
    for (long long c = (long long)blockIdx.x * nw + wid; c < ncells; c += (long long)gridDim.x * nw) {
        if (!mark[c]) continue;
        const long long f = c / per_frame;
        const long long rd = c - f * per_frame;
        const int r = (int)(rd / p.D), d = (int)(rd - (long long)r * p.D);
        const float2* cell = p.rds + ((size_t)f * p.R + r) * p.A * p.D + d;
        const CellResult res = large_cell(p, cell, (size_t)p.D, snap + wid * p.A, lane);
        if (lane == 0) cells[c] = res;
    }
}

// ESPRIT on the marked cells as ONE streaming pass, a thread per cell (large_cell's closed form, re-arranged after an ncu
// capture of the warp-per-cell eval_cells_kernel at 512 x 256 x 192: 1055 warp instructions per cell -- cross-lane
// reductions, the eigen-solve and fp64 asin(atan2()) replicated on 32 lanes -- the L1 pipe at 81 % from one 32-byte
// sector per 8-byte element).  The closed form needs only three lag sums of the snapshot:
//     T0 = sum |x_m|^2,   L1 = sum_{m < M-1} conj(x_m) x_{m+1},   L2 = sum_{m < M-2} conj(x_m) x_{m+2}
//   alpha = T0 - |x_{M-1}|^2, gamma = T0 - |x_0|^2, b = L1                      (the 2 x 2 Hermitian problem, :195-221)
//   n = sum_{m < M-2} conj(u_m) u_{m+1},  u_m = v0 x_m + v1 x_{m+1}
//     = |v0|^2 (L1 - conj(x_{M-2}) x_{M-1}) + conj(v0) v1 L2 + conj(v1) v0 (T0 - |x_0|^2 - |x_{M-1}|^2) + |v1|^2 (L1 - conj(x_0) x_1)
// so thread t of a CTA walks the antennas of cell c0 + t once (consecutive cells are consecutive Doppler bins: every
// load instruction of a warp is one contiguous 256-byte run), accumulates the three sums in fp64 and finishes the cell
// on its own: no shuffles, no shared memory, the snapshot read exactly once -- HBM bound.
__global__ void __launch_bounds__(ANG_THREADS) eval_cells_esprit_kernel(AngleArgs p, const uint8_t* __restrict__ mark,
                                                                         CellResult* __restrict__ cells, long long ncells) {
    const int M = p.A;
    const long long per_frame = (long long)p.R * p.D;
    for (long long c = (long long)blockIdx.x * ANG_THREADS + threadIdx.x; c < ncells; c += (long long)gridDim.x * ANG_THREADS) {
        if (!mark[c]) continue;
        const long long f = c / per_frame;
        const long long rd = c - f * per_frame;
        const int r = (int)(rd / p.D), d = (int)(rd - (long long)r * p.D);
        const float2* cell = p.rds + ((size_t)f * p.R + r) * M * p.D + d;
        double t0 = 0, l1r = 0, l1i = 0, l2r = 0, l2i = 0;
        double p1x = 0, p1y = 0, p2x = 0, p2y = 0;              // x_{m-1}, x_{m-2}
        double x0x = 0, x0y = 0, x1x = 0, x1y = 0;              // x_0, x_1
        float2 f0 = make_float2(0.f, 0.f), f1 = make_float2(0.f, 0.f);
        for (int m0 = 0; m0 < M; m0 += 8) {
            float2 xs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xs[j] = m0 + j < M ? __ldg(cell + (size_t)(m0 + j) * p.D) : make_float2(0.f, 0.f);
            if (m0 == 0) { f0 = xs[0]; f1 = xs[1]; x0x = xs[0].x; x0y = xs[0].y; x1x = xs[1].x; x1y = xs[1].y; }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (m0 + j < M) {
                    const double xx = xs[j].x, xy = xs[j].y;
                    t0 += xx * xx + xy * xy;
                    l1r += p1x * xx + p1y * xy;  l1i += p1x * xy - p1y * xx;        // conj(x_{m-1}) x_m  (zero for m = 0)
                    l2r += p2x * xx + p2y * xy;  l2i += p2x * xy - p2y * xx;        // conj(x_{m-2}) x_m
                    p2x = p1x; p2y = p1y; p1x = xx; p1y = xy;
                }
            }
        }
        // after the loop p1 = x_{M-1}, p2 = x_{M-2}
        const double e0 = x0x * x0x + x0y * x0y, eL = p1x * p1x + p1y * p1y;
        const double alpha = t0 - eL, gamma = t0 - e0, br = l1r, bi = l1i;
        const double half = 0.5 * (alpha - gamma);
        const double lam = 0.5 * (alpha + gamma) + sqrt(half * half + br * br + bi * bi);
        double v0r, v0i, v1r, v1i;
        const double na = br * br + bi * bi + (lam - alpha) * (lam - alpha);
        const double nb = (lam - gamma) * (lam - gamma) + br * br + bi * bi;
        if (na >= nb) { v0r = br; v0i = bi; v1r = lam - alpha; v1i = 0; }
        else          { v0r = lam - gamma; v0i = 0; v1r = br; v1i = -bi; }
        // the four data sums of n
        const double sar = l1r - (p2x * p1x + p2y * p1y), sai = l1i - (p2x * p1y - p2y * p1x);   // L1 - conj(x_{M-2}) x_{M-1}
        const double sbr = l1r - (x0x * x1x + x0y * x1y), sbi = l1i - (x0x * x1y - x0y * x1x);   // L1 - conj(x_0) x_1
        const double e1 = t0 - e0 - eL;
        const double a00 = v0r * v0r + v0i * v0i, a11 = v1r * v1r + v1i * v1i;
        const double cr = v0r * v1r + v0i * v1i, ci = v0r * v1i - v0i * v1r;                     // conj(v0) v1
        // n = a00 Sa + (conj(v0) v1) L2 + conj(conj(v0) v1) e1 + a11 Sb
        const double nr = a00 * sar + (cr * l2r - ci * l2i) + cr * e1 + a11 * sbr;
        const double ni = a00 * sai + (cr * l2i + ci * l2r) - ci * e1 + a11 * sbi;
        CellResult out{-1, 0.f, 0.f, 0u};
        out.phase = atan2f(f1.y * f0.x - f1.x * f0.y, f1.x * f0.x + f1.y * f0.y);
        out.adeg = (float)(asin(atan2(ni, nr) * p.esprit_scale) * (180.0 / 3.14159265358979323846));
        cells[c] = out;
    }
}

// ---- A > 16, MUSIC / beamforming: the steering scan as a dense contraction on the 5th-generation tensor cores ----------
// With many antennas y_g = a_g^H s over the whole grid IS a GEMM: [cells x 2A] . [2A x 2G] (real form; at A = 192, G = 181
// it is 384 x 362 per cell, 35 k complex multiply-adds, against 1.4 k for the 8-channel lag scan).  One CTA = 128 threads =
// 128 consecutive Doppler cells of one range row = the 128 TMEM lanes of a UMMA tile, so the snapshot gathers are perfectly
// coalesced (thread t reads antenna m of cell d0 + t: 1 KB rows).  Per tile:
//   pre-pass   |s|^2 of every cell -> scale 1 / |s| (the fp16 operands need a per-cell scale: a strong target is 10^4 x the noise)
//   per N-half (96 grid points = 192 accumulator columns: Re y, Im y interleaved), per chunk of 8 antennas:
//     every thread splits its scaled [Re s (8) | Im s (8)] into fp16 hi / lo and writes its row of the two A operands
//     (K-major, no swizzle), the CTA copies the chunk's [192 x 16] hi / lo B operands (radar_slam_b200/tables.py:
//     steer_tc_table, L2 resident), one thread issues three tcgen05.mma 128 x 192 x 16 (hi.hi + lo.hi + hi.lo, fp32
//     accumulation in TMEM) and commits to the stage's mbarrier; two stages, so the tensor core works on chunk c while
//     the threads prepare chunk c + 1 (the next chunk's snapshot loads are in flight across the barrier);
//   epilogue  tcgen05.ld: every thread owns the whole accumulator row of its cell: P_g = Re^2 + Im^2, two-level argmax /
//     runner-up tracking as in angles_tc5_kernel (groups of 8 grid points), carried over the halves.
// Every cell of the tile is evaluated (at these channel counts every cell carries detections); results go to the
// CellResult table the scatter kernel reads.  The TIE band is at least 2e-5: 1152 products per output in fp32.
namespace tc5 {
constexpr int SG_N = 192;                                                  // accumulator columns per N-half
constexpr uint32_t IDESC_SG = (1u << 4) | ((uint32_t)(SG_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ void umma_sg(uint32_t tmem_d, unsigned long long da, unsigned long long db, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC_SG), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait_trap(unsigned long long* bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    if (!ok) {
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
            if (!ok && clock64() - t0 > 4000000000ll) __trap();
        }
    }
}
}  // namespace tc5

__global__ void __launch_bounds__(ANG_THREADS, 2) music_tc_kernel(AngleArgs p, const uint4* __restrict__ btab, int nhalves,
                                                               int nchunks, CellResult* __restrict__ cells, long long ntiles) {
    using namespace tc5;
    constexpr int A_BYTES = 128 * 32, B_BYTES = SG_N * 32, A_STAGE = 2 * A_BYTES, B_STAGE = 2 * B_BYTES, SG_TMEM = 256;
    constexpr int B_STAGES = 4;
    extern __shared__ __align__(128) unsigned char sgsm[];                   // [2][A_hi, A_lo] then [4][B_hi, B_lo]
    __shared__ unsigned long long mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, wid = tid >> 5;
    const int M = p.A, G = p.G;
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "n"(SG_TMEM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t trow = tmem_base + ((uint32_t)(wid * 32) << 16);
    uint32_t phase[2] = {0u, 0u};
    const uint32_t arow = (uint32_t)((tid >> 3) * 256 + (tid & 7) * 16);
    const uint32_t sm_base = s32(sgsm);
    const float NEG = -3.0e38f;
    const float tie_eps = fmaxf(p.tie_eps, 2e-5f);
    const int tiles_d = p.D / 128;
    const bool exact = (M & 7) == 0;
    const size_t row8 = (size_t)8 * p.D * sizeof(float2);
    uint32_t roff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) roff[j] = (uint32_t)(j * p.D * (int)sizeof(float2));
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long fr = t / tiles_d;                                    // f * R + r
        const int d = (int)(t - fr * tiles_d) * 128 + tid;
        const float2* cell = p.rds + (size_t)fr * M * p.D + d;
        // pre-pass: energy of the cell (and the two antennas of the inter-antenna phase)
        float e = 0.f;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
        for (int m0 = 0; m0 < M; m0 += 32) {                                 // 32 rows in flight per thread
            float2 x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = m0 + j < M ? __ldg(cell + (size_t)(m0 + j) * p.D) : make_float2(0.f, 0.f);
            if (m0 == 0) { s0 = x[0]; s1 = x[1]; }
#pragma unroll
            for (int j = 0; j < 32; ++j) e = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, e));
        }
        const float scale = e > 0.f ? 1.f / sqrtf(e) : 0.f;
        // rows 8 c .. 8 c + 7 of this thread's cell: one base pointer per chunk, the eight row offsets are kernel constants
        auto load_chunk = [&](int c, float2 (&x)[8]) {
            const char* base = reinterpret_cast<const char*>(cell) + (size_t)c * row8;
            if (exact) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = __ldg(reinterpret_cast<const float2*>(base + roff[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    x[j] = 8 * c + j < M ? __ldg(reinterpret_cast<const float2*>(base + roff[j])) : make_float2(0.f, 0.f);
            }
        };
        float g_best = NEG, g_second = NEG;
        int g_idx = 0, g_run = 0;
        float stash[8] = {NEG, NEG, NEG, NEG, NEG, NEG, NEG, NEG};
        // B operands: a four-stage ring filled with cp.async two chunks ahead (stage (q + 2) & 3 was last read by the MMAs
        // of chunk q - 2, whose commit this thread has waited for by then); snapshot rows two chunks ahead in registers
        const int nseq = nhalves * nchunks;
        auto fetch_b = [&](int q) {
            if (q < nseq) {
                const uint4* src = btab + (size_t)q * (B_STAGE / 16);
                const uint32_t dst = sm_base + (uint32_t)(2 * A_STAGE + (q & (B_STAGES - 1)) * B_STAGE);
#pragma unroll
                for (int i = 0; i < B_STAGE / 16 / ANG_THREADS; ++i)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)((tid + i * ANG_THREADS) * 16)),
                                 "l"(src + tid + i * ANG_THREADS) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fetch_b(0);
        fetch_b(1);
        float2 nx[8], nx2[8];
        load_chunk(0, nx);
        load_chunk(nchunks > 1 ? 1 : 0, nx2);
        for (int h = 0; h < nhalves; ++h) {
            for (int c = 0; c < nchunks; ++c) {
                const int st = c & 1, q = h * nchunks + c;
                if (c >= 2) {                                                // the MMAs that read this stage two chunks ago
                    mbar_wait_trap(&mbar[st], phase[st]);
                    phase[st] ^= 1u;
                }
                unsigned char* sa = sgsm + (size_t)st * A_STAGE;
                {
                    uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        split_f16x2(nx[2 * j].x * scale, nx[2 * j + 1].x * scale, rh[j], rl[j]);
                        split_f16x2(nx[2 * j].y * scale, nx[2 * j + 1].y * scale, ih[j], il[j]);
                    }
                    *reinterpret_cast<uint4*>(sa + arow) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
                    *reinterpret_cast<uint4*>(sa + arow + 128) = make_uint4(ih[0], ih[1], ih[2], ih[3]);
                    *reinterpret_cast<uint4*>(sa + A_BYTES + arow) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
                    *reinterpret_cast<uint4*>(sa + A_BYTES + arow + 128) = make_uint4(il[0], il[1], il[2], il[3]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) nx[j] = nx2[j];
                if (q + 2 < nseq) load_chunk(c + 2 < nchunks ? c + 2 : c + 2 - nchunks, nx2);   // chunk q + 2 (wraps into the next half)
                asm volatile("cp.async.wait_group 1;" ::: "memory");         // this thread's part of B(q) has landed
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = sm_base + (uint32_t)(st * A_STAGE), a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = sm_base + (uint32_t)(2 * A_STAGE + (q & (B_STAGES - 1)) * B_STAGE), b_lo = b_hi + B_BYTES;
                    umma_sg(tmem_base, smem_desc(a_hi), smem_desc(b_hi), c > 0 ? 1u : 0u);
                    umma_sg(tmem_base, smem_desc(a_lo), smem_desc(b_hi), 1u);
                    umma_sg(tmem_base, smem_desc(a_hi), smem_desc(b_lo), 1u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar[st])) : "memory");
                }
                fetch_b(q + 2);
            }
            // the last two chunks' commits (in issue order): all MMAs of this half are done
            if (nchunks >= 2) {
                const int st = (nchunks - 2) & 1;
                mbar_wait_trap(&mbar[st], phase[st]);
                phase[st] ^= 1u;
            }
            {
                const int st = (nchunks - 1) & 1;
                mbar_wait_trap(&mbar[st], phase[st]);
                phase[st] ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // epilogue: 12 groups of 8 grid points (16 columns), the next group's read in flight
            float acc[2][16];
            ld16(trow, acc[0]);
#pragma unroll
            for (int g = 0; g < SG_N / 16; ++g) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g + 1 < SG_N / 16) ld16(trow + (uint32_t)(16 * (g + 1)), acc[(g + 1) & 1]);
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(acc[g & 1][2 * j], acc[g & 1][2 * j], acc[g & 1][2 * j + 1] * acc[g & 1][2 * j + 1]);
                const float gm = fmaxf(fmaxf(fmaxf(v[0], v[1]), v[2]), fmaxf(fmaxf(fmaxf(v[3], v[4]), v[5]), fmaxf(v[6], v[7])));
                const bool up = gm > g_best;
                g_second = fmaxf(g_second, fminf(g_best, gm));
                g_best = fmaxf(g_best, gm);
                g_idx = up ? g_run : g_idx;
                ++g_run;
#pragma unroll
                for (int j = 0; j < 8; ++j) stash[j] = up ? v[j] : stash[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the next half's first MMA comes after a barrier
        }
        int in_idx = 7;
#pragma unroll
        for (int j = 6; j >= 0; --j) in_idx = (stash[j] == g_best) ? j : in_idx;
        float in_second = NEG;
#pragma unroll
        for (int j = 0; j < 8; ++j) in_second = fmaxf(in_second, j == in_idx ? NEG : stash[j]);
        const float best = g_best, second = fmaxf(g_second, in_second);
        int bi = 8 * g_idx + in_idx;
        bi = bi < G ? bi : G - 1;
        CellResult out{bi, p.grid_deg[bi], 0.f, 0u};
        out.phase = atan2f(s1.y * s0.x - s1.x * s0.y, s1.x * s0.x + s1.y * s0.y);
        if ((best - second) <= tie_eps * best) out.flags |= RS_FLAG_TIE;
        if (p.method == RS_METHOD_MUSIC) {
            const float full = (float)M * (e * scale * scale);
            if (full - best <= 1e-4f * full) out.flags |= RS_FLAG_GUARD;
        }
        cells[(size_t)fr * p.D + d] = out;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(SG_TMEM) : "memory");
}

__global__ void __launch_bounds__(256) scatter_cells_kernel(AngleArgs p, const CellResult* __restrict__ cells) {
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    const int f = seg / p.nseg_per_frame;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t ld = p.det_lead[(size_t)seg * p.seg_cap + i];
        const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
        int a, r, d;
        rs_split_key(p.det_key[o], a, r, d);
        const CellResult c = cells[((size_t)f * p.R + r) * p.D + d];
        emit(p, seg, i, o, (int)(ld >> 16), c.aidx, c.adeg, c.phase, (uint8_t)c.flags);
    }
}

// ---- fp64 helpers for the legacy adapters -----------------------------------------------------
__global__ void signatures_f64_kernel(const float2* __restrict__ rds, const uint32_t* __restrict__ keys,
                                      const int32_t* __restrict__ frames, int n, double2* __restrict__ out, int R, int D,
                                      int A) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    int a, r, d;
    rs_split_key(keys[i], a, r, d);
    const float2* cell = rds + ((size_t)frames[i] * R + r) * A * D + d;
    double e = 0;
    for (int m = lane; m < A; m += 32) {
        const float2 x = cell[(size_t)m * D];
        e += (double)x.x * x.x + (double)x.y * x.y;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    // angle_estimation.py:86-88: divide by sqrt(power) only when power > 0
    const double sc = e > 0 ? sqrt(e) : 1.0;
    for (int m = lane; m < A; m += 32) {
        const float2 x = cell[(size_t)m * D];
        out[(size_t)i * A + m] = make_double2((double)x.x / sc, (double)x.y / sc);
    }
}

__global__ void spectra_f64_kernel(const double2* __restrict__ sig, const double2* __restrict__ steer, int method, int A,
                                   int G, double* __restrict__ out, int32_t* __restrict__ aidx) {
    extern __shared__ double2 sm_s[];   // [A]
    __shared__ double red_v[256];
    __shared__ int red_i[256];
    const int i = blockIdx.x;
    for (int m = threadIdx.x; m < A; m += blockDim.x) sm_s[m] = sig[(size_t)i * A + m];
    __syncthreads();
    // eigenvectors of s s^H do not depend on |s|: MUSIC sees the unit-energy snapshot
    double energy = 0;
    for (int m = 0; m < A; ++m) energy += sm_s[m].x * sm_s[m].x + sm_s[m].y * sm_s[m].y;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double ar = 0, ai = 0;
        for (int m = 0; m < A; ++m) {
            const double2 w = steer[(size_t)m * G + g];
            const double2 x = sm_s[m];
            ar += w.x * x.x + w.y * x.y;
            ai += w.x * x.y - w.y * x.x;
        }
        double v = ar * ar + ai * ai;
        if (method == RS_METHOD_MUSIC) {
            // a^H E_n E_n^H a = M - |a^H s|^2 for a unit-energy snapshot; guard at angle_estimation.py:149
            // (a zero snapshot gives eigh(0) = I: the noise subspace is M-1 unit vectors, den = M-1)
            const double den = fabs((double)A - (energy > 0 ? v / energy : 1.0));
            v = den > 1e-12 ? 1.0 / den : 0.0;
        }
        out[(size_t)i * G + g] = v;
        // np.argmax (angle_estimation.py:173): a NaN is the maximum, the first NaN wins -- encoded as +inf here, where
        // the lowest index already wins ties
        if (v != v) v = INFINITY;
        if (v > best) { best = v; bi = g; }
    }
    red_v[threadIdx.x] = best;
    red_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x >> 1; s; s >>= 1) {
        if (threadIdx.x < s) {
            const double ov = red_v[threadIdx.x + s];
            const int oi = red_i[threadIdx.x + s];
            if (ov > red_v[threadIdx.x] || (ov == red_v[threadIdx.x] && oi < red_i[threadIdx.x])) {
                red_v[threadIdx.x] = ov;
                red_i[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) aidx[i] = red_i[0] < G ? red_i[0] : 0;
}

}  // namespace

// |X|^2 of every entry, gathered from the RDS, for the paths whose scan kernel does not write it on the way
__global__ void __launch_bounds__(256) det_power_kernel(AngleArgs p) {
    const int seg = blockIdx.x;
    const int n = p.det_nlead[seg];
    const int f = seg / p.nseg_per_frame;
    const float2* frame = p.rds + (size_t)f * p.R * p.D * p.A;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t ld = p.det_lead[(size_t)seg * p.seg_cap + i];
        const size_t o = (size_t)seg * p.seg_cap + (ld & 0xFFFFu);
        const int k = (int)(ld >> 16);
        for (int e = 0; e < k; ++e) {
            int a, r, d;
            rs_split_key(p.det_key[o + e], a, r, d);
            const float2 x = __ldg(frame + ((size_t)r * p.A + a) * p.D + d);
            p.det_power[o + e] = fmaf(x.x, x.x, x.y * x.y);
        }
    }
}

extern "C" int rs_angles(const void* rds, const float* scan_table, int scan_stride, const void* steer,
                         const float* grid_deg, int G, int method, float tie_eps, double esprit_scale,
                         const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead, uint8_t* det_flags,
                         int32_t* det_aidx,
                         float* det_adeg, float* det_phase, int seg_cap, int nseg_per_frame, int F, int R, int D, int A,
                         const double* grid_cs, double* ls_partials, int grid_symmetric, int32_t* det_ntie,
                         int32_t* det_tielist, const float* mma_table, int mma_tiles, void* cell_ws, const void* tc_table,
                         int tc_halves, float* det_power_out, const int32_t* det_nnear, void* stream) {
    RS_CHECK_ARG(rds && det_key && det_lead && det_nlead && det_flags && det_aidx && det_adeg && det_phase,
                 "rs_angles: null pointer");
    RS_CHECK_ARG(ls_partials == nullptr || grid_cs != nullptr, "rs_angles: ls_partials needs grid_cs");
    RS_CHECK_ARG(method >= 0 && method <= 2, "rs_angles: unknown method %d", method);
    RS_CHECK_ARG(A >= 2 && A <= RS_MAX_ANTENNAS, "rs_angles: need 2 <= A <= %d", RS_MAX_ANTENNAS);
    RS_CHECK_ARG(F > 0 && R > 0 && D > 0 && seg_cap > 0 && nseg_per_frame > 0, "rs_angles: bad dims");
    const bool scan = method != RS_METHOD_ESPRIT;
    RS_CHECK_ARG(!scan || (G > 0 && grid_deg), "rs_angles: grid required");
    AngleArgs p{(const float2*)rds, scan_table, scan_stride, (const float2*)steer, grid_deg, G, method, tie_eps,
                esprit_scale, det_key, det_lead, det_nlead, det_ntie, det_tielist, det_flags, det_aidx, det_adeg, det_phase, seg_cap,
                nseg_per_frame,
                R, D, A, det_power_out, det_nnear};
    const long long blocks = (long long)F * nseg_per_frame;
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_angles: too many segments");
    cudaStream_t st = (cudaStream_t)stream;
    if (det_ntie) cudaMemsetAsync(det_ntie, 0, sizeof(int32_t) * (size_t)blocks, st);
    if (A <= 16) {
        RS_CHECK_ARG(!scan || scan_table, "rs_angles: scan_table required for A <= 16");
        const int ap = A <= 2 ? 2 : A <= 4 ? 4 : A <= 8 ? 8 : 16;
        int need_stride = 2 * (ap - 1);
        need_stride = (need_stride + 3) & ~3;
        RS_CHECK_ARG(!scan || scan_stride == need_stride, "rs_angles: scan_stride must be %d for A=%d", need_stride, A);
        const int rows = (scan && grid_symmetric) ? (G + 1) / 2 : G;
        const size_t smem = scan ? (size_t)rows * scan_stride * sizeof(float) : 0;
        if (smem > (size_t)rs_smem_optin_limit()) {
            rs_set_error("rs_angles: grid table needs %zu B of shared memory", smem);
            return RS_ECAPACITY;
        }
        if (!scan) {
            RS_CHECK_ARG(ls_partials == nullptr, "rs_angles: ls_partials is only produced by the grid methods");
#define LAUNCH_SMALL(AP) angles_small_kernel<AP><<<(unsigned)blocks, ANG_THREADS, 0, st>>>(p)
            if (ap == 2) { LAUNCH_SMALL(2); }
            else if (ap == 4) { LAUNCH_SMALL(4); }
            else if (ap == 8) { LAUNCH_SMALL(8); }
            else { LAUNCH_SMALL(16); }
#undef LAUNCH_SMALL
        } else {
#define LAUNCH_SCAN(AP, ND)                                                                                              \
    do {                                                                                                                 \
        if (grid_symmetric) {                                                                                            \
            cudaFuncSetAttribute(angles_scan_kernel<AP, ND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            angles_scan_kernel<AP, ND, true><<<(unsigned)blocks, ANG_THREADS, smem, st>>>(p, grid_cs, ls_partials);      \
        } else {                                                                                                         \
            cudaFuncSetAttribute(angles_scan_kernel<AP, ND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            angles_scan_kernel<AP, ND, false><<<(unsigned)blocks, ANG_THREADS, smem, st>>>(p, grid_cs, ls_partials);     \
        }                                                                                                                \
    } while (0)
            // RS_ANGLES_TC: 1 / 0 force the tcgen05 / TMEM scan on / off.  Default: on.  With the two-level tracking, the
            // persistent CTAs and the packed lag arithmetic it beats the mma.sync scan at both widths: 1.12 vs 1.56 ms per
            // 300 frames of 256 x 128 x 16, 1.04 vs 1.26 ms per 1000 frames of 256 x 128 x 8 (profiles/angles_bench.py)
            const char* tc_env = getenv("RS_ANGLES_TC");
            const bool use_tc = tc_env ? atoi(tc_env) == 1 : true;
            if (tc_table != nullptr && tc_halves > 0 && grid_symmetric && (ap == 8 || ap == 16) && use_tc) {
                RS_CHECK_ARG(tc_halves == ((G + 1) / 2 + tc5::NPH - 1) / tc5::NPH, "rs_angles: tc_halves must be ceil(ceil(G/2)/32)");
                const int kc = ap == 8 ? 2 : 3;
                size_t sm = (size_t)tc_halves * 2 * kc * tc5::NPH * 32 + 2 * 2 * 128 * 32 +
                            (size_t)(128 + (G + 1) / 2) * (ap + 4) * sizeof(float) + 16 + (size_t)G * 20;
                // at most four CTAs (4 x 128 TMEM columns) may be resident on an SM: pad the request so that a fifth never fits
                const size_t floor4 = (size_t)rs_smem_optin_limit() / 5 + 1024;
                if (sm < floor4) sm = floor4;
                if (sm <= (size_t)rs_smem_optin_limit()) {
                    // persistent: as many CTAs as fit (4 per SM by TMEM columns, fewer when the tables of a fine grid are large)
                    const size_t per_sm = (size_t)rs_smem_optin_limit() / (sm + 1024);
                    long long resident = (long long)rs_sm_count() * (long long)(per_sm < 4 ? (per_sm < 1 ? 1 : per_sm) : 4);
                    // pair mode (see the kernel): 16 antennas in two octet segments per tile, scratch in cell_ws
                    unsigned char* ws = nullptr;
                    int rmask = 1, dshift = 1;                                   // tile rows / Doppler bins
                    const char* dd_env = getenv("RS_ANGLES_DEDUP");
                    if (A == 16 && cell_ws != nullptr && det_power_out == nullptr && !(dd_env && atoi(dd_env) == 0)) {
                        int tr = 0, td = 0, nt = 0;
                        rs_detect_tiling(R, D, A, &tr, &td, &nt);
                        const bool pow2 = tr > 0 && td > 0 && (tr & (tr - 1)) == 0 && (td & (td - 1)) == 0;
                        if (pow2 && nt == nseg_per_frame && (nt & 1) == 0 && tr * td <= tc5::PAIR_CELLS && D % td == 0 && R % tr == 0) {
                            ws = (unsigned char*)cell_ws;
                            rmask = tr;
                            dshift = td;
                            const long long cap = RS_ANGLES_WS_BYTES / tc5::PAIR_WS_BYTES;
                            if (resident > cap) resident = cap;
                        }
                    }
                    const long long units = ws ? blocks / 2 : blocks;
                    const unsigned tc_grid = (unsigned)(units < resident ? units : resident);
#define LAUNCH_TC5(AP_, PAIR_)                                                                                          \
    do {                                                                                                                \
        cudaFuncSetAttribute(angles_tc5_kernel<AP_, PAIR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);      \
        angles_tc5_kernel<AP_, PAIR_><<<tc_grid, ANG_THREADS, sm, st>>>(p, (const uint4*)tc_table, tc_halves, grid_cs,  \
                                                                        ls_partials, (int)blocks, ws, rmask, dshift);   \
    } while (0)
                    if (ap == 8) LAUNCH_TC5(8, false);
                    else if (ws != nullptr) LAUNCH_TC5(16, true);
                    else LAUNCH_TC5(16, false);
#undef LAUNCH_TC5
                    RS_CHECK_LAUNCH("rs_angles(tcgen05)");
                    return RS_OK;
                }
            }
            const char* mma_env = getenv("RS_ANGLES_MMA");  // 0: force the CUDA-core scan
            if (mma_table != nullptr && grid_symmetric && (ap == 8 || ap == 16) && !(mma_env && atoi(mma_env) == 0)) {
                RS_CHECK_ARG(mma_tiles == ((G + 1) / 2 + 7) / 8, "rs_angles: mma_tiles must be ceil(ceil(G/2)/8)");
                size_t sm = ((size_t)mma_tiles * (ap / 8) * 192 + (size_t)(ANG_THREADS / 32) * 32 * (2 * ap + 8)) *
                            sizeof(float);
                // RS_ANGLES_CTAS: resident CTAs per SM (by padding the shared-memory request).  Fewer than the 8 that fill the
                // register file leave room for the CTAs of the fp64 recheck kernels, which FramePipeline.process runs on a
                // second stream beside this scan
                const char* cta_env = getenv("RS_ANGLES_CTAS");
                const int ctas = cta_env ? atoi(cta_env) : 0;
                if (ctas >= 1 && ctas < 8) {
                    const size_t pad = (size_t)rs_smem_optin_limit() / (ctas + 1) + 1024;
                    if (sm < pad) sm = pad;
                }
                if (sm <= (size_t)rs_smem_optin_limit()) {
                    // occupancy beats per-warp ILP here (measured): 64 registers / 8 CTAs per SM at A <= 8
                    // (9 / 10 CTAs at 56 / 48 registers spill more than the extra warps hide: 1.49 / 1.64 ms against 1.26)
#define LAUNCH_MMA(KERN)                                                                                       \
    do {                                                                                                       \
        cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                     \
        KERN<<<(unsigned)blocks, ANG_THREADS, sm, st>>>(p, (const uint32_t*)mma_table, mma_tiles, grid_cs, ls_partials); \
    } while (0)
                    if (ap == 8) {
                        if (det_power_out) LAUNCH_MMA((angles_mma_kernel<8, 8, true>));
                        else LAUNCH_MMA((angles_mma_kernel<8, 8>));
                    } else {
                        if (det_power_out) LAUNCH_MMA((angles_mma_kernel<16, 5, true>));
                        else LAUNCH_MMA((angles_mma_kernel<16, 5>));
                    }
#undef LAUNCH_MMA
                    RS_CHECK_LAUNCH("rs_angles(mma)");
                    return RS_OK;
                }
            }
            const char* nd_env = getenv("RS_SCAN_ND");      // tuning knob (detections carried per thread)
            const int nd8 = nd_env ? atoi(nd_env) : 2;      // measured on B200: ND=2 4.12 ms, 3: 4.49, 4: 4.53 per 1k frames
            if (ap == 2) LAUNCH_SCAN(2, 4);
            else if (ap == 4) LAUNCH_SCAN(4, 4);
            else if (ap == 8) {
                if (nd8 == 2) LAUNCH_SCAN(8, 2);
                else if (nd8 == 3) LAUNCH_SCAN(8, 3);
                else LAUNCH_SCAN(8, 4);
            } else LAUNCH_SCAN(16, 2);
#undef LAUNCH_SCAN
        }
    } else {
        RS_CHECK_ARG(ls_partials == nullptr, "rs_angles: ls_partials is not produced for A > 16 (use rs_velocity_ls)");
        RS_CHECK_ARG(!scan || steer, "rs_angles: steer table required for A > 16");
        const size_t smem = (size_t)(ANG_THREADS / 32) * A * sizeof(float2);
        if (cell_ws != nullptr) {
            // workspace: CellResult [F*R*D] followed by the mark bytes [F*R*D]
            const long long ncells = (long long)F * R * D;
            CellResult* cells = (CellResult*)cell_ws;
            uint8_t* mark = (uint8_t*)(cells + ncells);
            cudaMemsetAsync(mark, 0, (size_t)ncells, st);
            mark_cells_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, mark);
            const long long want = (ncells + ANG_THREADS / 32 - 1) / (ANG_THREADS / 32);
            const long long cap = (long long)rs_sm_count() * 64;
            const char* mtc_env = getenv("RS_MUSIC_TC");
            const int sg_halves = (G + 95) / 96;
            if (method != RS_METHOD_ESPRIT && tc_table != nullptr && tc_halves == sg_halves && D % 128 == 0 &&
                !(mtc_env && atoi(mtc_env) == 0)) {
                // steering GEMM on the tensor cores: every cell of the batch, 128-cell tiles, two CTAs (2 x 256 TMEM columns) per SM
                const long long ntiles = ncells / 128;
                const size_t smem_sg = 2 * (2 * 128 * 32) + 4 * (2 * (size_t)tc5::SG_N * 32);
                const long long resident = 2ll * rs_sm_count();
                cudaFuncSetAttribute(music_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sg);
                music_tc_kernel<<<(unsigned)(ntiles < resident ? ntiles : resident), ANG_THREADS, smem_sg, st>>>(
                    p, (const uint4*)tc_table, sg_halves, (A + 7) / 8, cells, ntiles);
            } else if (method == RS_METHOD_ESPRIT && A >= 3 && !getenv("RS_ESPRIT_LEGACY")) {
                const long long wantc = (ncells + ANG_THREADS - 1) / ANG_THREADS;
                eval_cells_esprit_kernel<<<(unsigned)(wantc < cap ? wantc : cap), ANG_THREADS, 0, st>>>(p, mark, cells, ncells);
            } else {
                eval_cells_kernel<<<(unsigned)(want < cap ? want : cap), ANG_THREADS, smem, st>>>(p, mark, cells, ncells);
            }
            scatter_cells_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, cells);
        } else {
            angles_large_kernel<<<(unsigned)blocks, ANG_THREADS, smem, st>>>(p);
        }
    }
    RS_CHECK_LAUNCH("rs_angles");
    if (det_power_out) {             // only the tensor-core scans write the powers on the way
        det_power_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
        RS_CHECK_LAUNCH("rs_angles(power)");
    }
    return RS_OK;
}

extern "C" int rs_detection_power(const void* rds, const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead,
                                  float* det_power, int seg_cap, int nseg_per_frame, int F, int R, int D, int A, void* stream) {
    RS_CHECK_ARG(rds && det_key && det_lead && det_nlead && det_power, "rs_detection_power: null pointer");
    RS_CHECK_ARG(F > 0 && R > 0 && D > 0 && A > 0 && seg_cap > 0 && nseg_per_frame > 0, "rs_detection_power: bad dims");
    const long long blocks = (long long)F * nseg_per_frame;
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_detection_power: too many segments");
    AngleArgs p{};
    p.rds = (const float2*)rds;
    p.det_key = det_key;
    p.det_lead = det_lead;
    p.det_nlead = det_nlead;
    p.det_power = det_power;
    p.seg_cap = seg_cap;
    p.nseg_per_frame = nseg_per_frame;
    p.R = R;
    p.D = D;
    p.A = A;
    det_power_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    RS_CHECK_LAUNCH("rs_detection_power");
    return RS_OK;
}

extern "C" int rs_signatures_f64(const void* rds, const uint32_t* keys, const int32_t* frames, int n, void* out, int F,
                                 int R, int D, int A, void* stream) {
    RS_CHECK_ARG(rds && keys && frames && out && n >= 0 && F > 0, "rs_signatures_f64: bad args");
    if (n == 0) return RS_OK;
    const int wpb = 8;
    signatures_f64_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        (const float2*)rds, keys, frames, n, (double2*)out, R, D, A);
    RS_CHECK_LAUNCH("rs_signatures_f64");
    return RS_OK;
}

extern "C" int rs_spectra_f64(const void* sig128, const void* steer128, int method, int n, int A, int G, double* out,
                              int32_t* aidx, void* stream) {
    RS_CHECK_ARG(sig128 && steer128 && out && aidx && n >= 0 && A > 0 && G > 0, "rs_spectra_f64: bad args");
    RS_CHECK_ARG(method == RS_METHOD_MUSIC || method == RS_METHOD_BEAMFORMING, "rs_spectra_f64: bad method");
    if (n == 0) return RS_OK;
    spectra_f64_kernel<<<n, 256, (size_t)A * sizeof(double2), (cudaStream_t)stream>>>(
        (const double2*)sig128, (const double2*)steer128, method, A, G, out, aidx);
    RS_CHECK_LAUNCH("rs_spectra_f64");
    return RS_OK;
}
