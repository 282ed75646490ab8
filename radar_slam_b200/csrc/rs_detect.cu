// Peak detection on the RDS rds[F][R][A][D] (rs_common.cuh).
//
// Replaces SignalPreprocessor.extract_range_doppler_peaks (dechirp.py:215-278).  The reference
// compares dB values; 10 log10(p + 1e-12) is monotone in p, so the local-maximum test and the
// threshold are evaluated on the linear power p = |X|^2 (host converts the threshold in fp64).
// scipy.ndimage.maximum_filter(size=3) with its default 'reflect' boundary duplicates the edge
// row/column, i.e. out-of-range neighbours are ignored; ties count as maxima (== test).
//
// A CTA owns a tile of TR range bins x TD Doppler bins x AC antennas (+1 halo), computes the
// power plane into shared memory, walks columns with a 3-row sliding window in registers, and
// compacts the hits of the tile into its own fixed-capacity segment with a block-wide prefix sum
// (deterministic order, no atomics).
//
// Detections of the same range-Doppler cell on different antennas share one snapshot, hence one
// angle and one inter-antenna phase.  The fast kernel therefore emits a cell's detections
// contiguously and records one LEADER per cell -- det_lead = position | (multiplicity << 16) -- so
// the angle stage evaluates every distinct cell once (-28 % work at 8 channels, -48 % at 16).
#include <cmath>
#include <cstdlib>
#include "rs_common.cuh"
#include "rs_detect_fused.cuh"

namespace {

constexpr int DET_THREADS = 256;
constexpr int DET_MAX_COLS_PER_THREAD = 8;

struct Tiling {
    int TR, TD, AC, ntr, ntd, nac, ntiles;
};

static Tiling make_tiling(int R, int D, int A) {
    Tiling t;
    t.AC = A < 8 ? A : 8;
    t.TD = D < 128 ? D : 128;
    // keep TD * AC <= threads * cols-per-thread
    while (t.TD * t.AC > DET_THREADS * DET_MAX_COLS_PER_THREAD) t.TD /= 2;
    t.TR = R < 16 ? R : 16;
    t.ntr = (R + t.TR - 1) / t.TR;
    t.ntd = (D + t.TD - 1) / t.TD;
    t.nac = (A + t.AC - 1) / t.AC;
    t.ntiles = t.ntr * t.ntd * t.nac;
    return t;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = blockDim.x >> 5;
        int w = lane < nw ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) warp_sums[lane] = winc - w;
        if (lane == nw - 1) *total = winc;
    }
    __syncthreads();
    return warp_sums[wid] + inc - v;
}

struct DetOut {
    uint32_t* key;
    float* power;
    uint8_t* flags;
    uint32_t* lead;
    int32_t* count;
    int32_t* nlead;
    int32_t* overflow;
    int32_t* nnear;      // per segment: entries carrying RS_FLAG_NEARMAX (lets the fp64 recheck skip clean segments)
    float* psum;         // per segment: sum of |X|^2 over the tile (frame noise level for the recheck's error bound)
    int seg_cap;
};

// block-wide sums of an int and a float into shared scalars (integer sum is order independent; the float sum is
// reduced in a fixed shuffle/warp order, so both are deterministic)
__device__ __forceinline__ void block_sums(int iv, float fv, int* s_i, float* s_f) {
    __shared__ float wf[DET_THREADS / 32];
    __shared__ int wi[DET_THREADS / 32];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        iv += __shfl_xor_sync(0xffffffffu, iv, off);
        fv += __shfl_xor_sync(0xffffffffu, fv, off);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { wi[wid] = iv; wf[wid] = fv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ti = 0;
        float tf = 0.f;
        for (int w = 0; w < DET_THREADS / 32; ++w) { ti += wi[w]; tf += wf[w]; }
        *s_i = ti;
        *s_f = tf;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// generic kernel: any A / D / R (every detection is its own leader)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DET_THREADS)
detect_kernel(const float2* __restrict__ rds, const uint8_t* __restrict__ gate, float thr, float eps, DetOut out, int R,
              int D, int A, Tiling tl) {
    extern __shared__ float pw[];   // [(TR+2)][(TD+2)][AC]
    __shared__ int warp_sums[DET_THREADS / 32];
    __shared__ int total_s;

    const int tile = blockIdx.x % tl.ntiles;
    const int f = blockIdx.x / tl.ntiles;
    const int ia = tile % tl.nac;
    const int id = (tile / tl.nac) % tl.ntd;
    const int ir = tile / (tl.nac * tl.ntd);
    const int r0 = ir * tl.TR, d0 = id * tl.TD, a0 = ia * tl.AC;
    const int TR = tl.TR, TD = tl.TD, AC = tl.AC;
    const int W = (TD + 2) * AC;                 // floats per tile row
    const float2* frame = rds + (size_t)f * R * D * A;

    // ---- power plane (with halo); out-of-range neighbours get -1 so they never win
    const int n_el = (TR + 2) * W;
    for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
        const int rr = i / W;
        const int rem = i - rr * W;
        const int dd = rem / AC, ac = rem - dd * AC;
        const int r = r0 - 1 + rr, d = d0 - 1 + dd, a = a0 + ac;
        float p = -1.f;
        if (r >= 0 && r < R && d >= 0 && d < D && a < A) {
            const float2 x = __ldg(frame + ((size_t)r * A + a) * D + d);
            p = fmaf(x.x, x.x, x.y * x.y);
        }
        pw[i] = p;
    }
    __syncthreads();

    // ---- column walk: masks of hits / near-ties per column, kept in registers
    const int ncols = TD * AC;
    uint32_t hit[DET_MAX_COLS_PER_THREAD], near[DET_MAX_COLS_PER_THREAD], cand[DET_MAX_COLS_PER_THREAD];
    int my_count = 0, my_near = 0;
    float my_psum = 0.f;
#pragma unroll
    for (int q = 0; q < DET_MAX_COLS_PER_THREAD; ++q) {
        hit[q] = 0u;
        near[q] = 0u;
        cand[q] = 0u;
        const int col = threadIdx.x + q * DET_THREADS;
        if (col < ncols) {
            const int dd = col / AC + 1, ac = col - (col / AC) * AC;
            const int d = d0 + dd - 1, a = a0 + ac;
            if (d < D && a < A) {
                const float* base = pw + dd * AC + ac;
                float l0 = base[-AC], c0 = base[0], rt0 = base[AC];
                float h_prev = fmaxf(fmaxf(l0, c0), rt0);
                float l1 = base[W - AC], c1 = base[W], rt1 = base[W + AC];
                for (int rr = 1; rr <= TR; ++rr) {
                    const float* nx = base + (rr + 1) * W;
                    const float l2 = nx[-AC], c2 = nx[0], rt2 = nx[AC];
                    const float h_next = fmaxf(fmaxf(l2, c2), rt2);
                    const int r = r0 + rr - 1;
                    if (r < R) {
                        const float m = fmaxf(fmaxf(h_prev, h_next), fmaxf(l1, rt1));
                        const int cls = rs_classify(c1, m, thr, eps);
                        if (cls && gate[r]) {
                            hit[q] |= 1u << (rr - 1);
                            ++my_count;
                            if (cls & 2) { near[q] |= 1u << (rr - 1); ++my_near; }
                            if (cls & 4) cand[q] |= 1u << (rr - 1);
                        }
                        my_psum += c1;
                    }
                    h_prev = fmaxf(fmaxf(l1, c1), rt1);
                    l1 = l2; c1 = c2; rt1 = rt2;
                }
            }
        }
    }

    // ---- deterministic compaction into this tile's segment
    __shared__ int near_s;
    __shared__ float psum_s;
    block_sums(my_near, my_psum, &near_s, &psum_s);
    const int offset = block_exclusive_scan(my_count, warp_sums, &total_s);
    const int total = total_s;
    const size_t seg = (size_t)blockIdx.x;
    if (threadIdx.x == 0) {
        const int n = total < out.seg_cap ? total : out.seg_cap;
        out.count[seg] = n;
        out.nlead[seg] = n;
        if (total > out.seg_cap) out.overflow[f] = 1;
        if (out.nnear) out.nnear[seg] = near_s;
        if (out.psum) out.psum[seg] = psum_s;
    }
    int pos = offset;
#pragma unroll
    for (int q = 0; q < DET_MAX_COLS_PER_THREAD; ++q) {
        uint32_t m = hit[q];
        if (m) {
            const int col = threadIdx.x + q * DET_THREADS;
            const int dd = col / AC + 1, ac = col - (col / AC) * AC;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (pos < out.seg_cap) {
                    const size_t o = seg * out.seg_cap + pos;
                    out.key[o] = rs_make_key(a0 + ac, r0 + b, d0 + dd - 1);
                    out.power[o] = pw[(b + 1) * W + dd * AC + ac];
                    out.flags[o] = ((cand[q] >> b) & 1u) ? (RS_FLAG_NEARMAX | RS_FLAG_DROPPED)
                                                        : ((near[q] >> b) & 1u) ? RS_FLAG_NEARMAX : 0;
                    out.lead[o] = (uint32_t)pos | (1u << 16);
                }
                ++pos;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path: antenna octets (A % 8 == 0), tile = 16 range bins x 128 Doppler bins x 8 antennas.
// 256 threads; thread = (Doppler bin, half of the tile's rows) and owns ALL 8 antennas of its cells, so
// the detections of one cell are emitted together.  A tile row is one contiguous 8 KB run of the RDS (8 antenna rows
// of 128 cells); a warp loads 8 antennas x 4 consecutive cells per step and transposes into the power plane, which is
// [row][doppler + 1][8] floats with the two antenna quads of a cell XOR-swizzled by bit 2 of the Doppler
// index, which makes both the float2 stores of the load phase and the LDS.128 reads of the walk
// conflict-free.
// ---------------------------------------------------------------------------------------------
constexpr int A8_TR = 16, A8_TD = 128, A8_AC = 8, A8_W = (A8_TD + 2) * A8_AC, A8_HALF = A8_TR / 2;

__device__ __forceinline__ int a8_off(int rr, int ddp, int quad) {          // float offset of (row, doppler+1, quad)
    return rr * A8_W + ddp * A8_AC + ((quad ^ ((ddp >> 2) & 1)) << 2);
}

// Tail of detect_a8_kernel: per-thread masks (byte per row: antenna mask; rows rbase+1 .. rbase+8 of the tile, Doppler
// column ddp) -> the tile's segment.  power_at(rr, j): |X|^2 of row rr (1-based), antenna j.
template <class PowerAt>
__device__ __forceinline__ void a8_emit(const uint32_t (&hit)[2], const uint32_t (&near)[2], const uint32_t (&cand)[2],
                                        int my_near, float my_psum, const DetOut& out, int f, int r0, int d0, int a0,
                                        int rbase, int ddp, PowerAt power_at) {
    __shared__ int warp_sums[DET_THREADS / 32];
    __shared__ int total_s;
    const int tid = threadIdx.x;
    // entries and leaders of this thread, scanned together (entries <= 2^14 per tile)
    const int n_ent = __popc(hit[0]) + __popc(hit[1]);
    int n_lead = 0;
#pragma unroll
    for (int i = 0; i < A8_HALF; ++i) n_lead += ((hit[i >> 2] >> ((i & 3) * 8)) & 0xFFu) != 0u;
    const int packed = block_exclusive_scan(n_ent | (n_lead << 16), warp_sums, &total_s);
    const int total = total_s & 0xFFFF, total_lead = total_s >> 16;
    const size_t seg = (size_t)blockIdx.x;
    int pos = packed & 0xFFFF, lpos = packed >> 16;
    if (tid == 0) {
        out.count[seg] = total < out.seg_cap ? total : out.seg_cap;
        if (total > out.seg_cap) out.overflow[f] = 1;
    }
    {
        __shared__ int near_s;
        __shared__ float psum_s;
        block_sums(my_near, my_psum, &near_s, &psum_s);
        if (tid == 0) {
            if (out.nnear) out.nnear[seg] = near_s;
            if (out.psum) out.psum[seg] = psum_s;
        }
    }
    // a leader is kept only if all of its cell's entries fit; because entries are emitted in order, the kept
    // leaders are a prefix of the leader list -> nlead = number of leaders whose last entry fits
    int kept = 0;
#pragma unroll
    for (int i = 0; i < A8_HALF; ++i) {
        uint32_t m = (hit[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
        if (!m) continue;
        const uint32_t nb = (near[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
        const uint32_t cb = (cand[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
        const int k = __popc(m);
        const int rr = rbase + 1 + i;
        if (pos + k <= out.seg_cap) {
            out.lead[seg * out.seg_cap + lpos] = (uint32_t)pos | ((uint32_t)k << 16);
            ++kept;
        }
        ++lpos;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            if (pos < out.seg_cap) {
                const size_t o = seg * out.seg_cap + pos;
                out.key[o] = rs_make_key(a0 + j, r0 + rr - 1, d0 + ddp - 1);
                if (out.power) out.power[o] = power_at(rr, j);
                out.flags[o] = ((cb >> j) & 1u) ? (RS_FLAG_NEARMAX | RS_FLAG_DROPPED) : ((nb >> j) & 1u) ? RS_FLAG_NEARMAX : 0;
            }
            ++pos;
        }
    }
    // number of kept leaders: all of them unless the segment overflowed (then a block sum, uniform branch)
    if (total <= out.seg_cap) {
        if (tid == 0) out.nlead[seg] = total_lead;
    } else {
        __syncthreads();
        block_exclusive_scan(kept, warp_sums, &total_s);
        if (tid == 0) out.nlead[seg] = total_s;
    }
}


__global__ void __launch_bounds__(DET_THREADS, 3)
detect_a8_kernel(const float2* __restrict__ rds, const uint8_t* __restrict__ gate, float thr, float eps, DetOut out, int R,
                 int D, int A, Tiling tl) {
    extern __shared__ float pw[];   // [(TR+2)][W]

    const int tile = blockIdx.x % tl.ntiles;
    const int f = blockIdx.x / tl.ntiles;
    const int ia = tile % tl.nac;
    const int id = (tile / tl.nac) % tl.ntd;
    const int ir = tile / (tl.nac * tl.ntd);
    const int r0 = ir * A8_TR, d0 = id * A8_TD, a0 = ia * A8_AC;
    const float2* frame = rds + (size_t)f * R * D * A;
    const int tid = threadIdx.x;

    // ---- halo columns (only present when the frame is wider than one tile) and out-of-range fill
    for (int i = tid; i < (A8_TR + 2) * 2 * A8_AC; i += DET_THREADS) {
        const int rr = i / (2 * A8_AC), rem = i - rr * 2 * A8_AC;
        const int side = rem / A8_AC, ac = rem - side * A8_AC;
        const int r = r0 - 1 + rr, d = side ? d0 + A8_TD : d0 - 1;
        float p = -1.f;
        if (r >= 0 && r < R && d >= 0 && d < D) {
            const float2 x = __ldg(frame + ((size_t)r * A + a0 + ac) * D + d);
            p = fmaf(x.x, x.x, x.y * x.y);
        }
        pw[a8_off(rr, side ? A8_TD + 1 : 0, ac >> 2) + (ac & 3)] = p;
    }
    // ---- interior: a tile row is [8 antennas][128 cells] contiguous.  lane = (antenna a, cell pair j): one float4 load
    // brings two consecutive cells of antenna a (a warp reads eight 64-byte runs); antenna-neighbour lanes then swap one
    // power each, so that every lane holds (one cell, antennas 2m and 2m+1) and the plane is filled with conflict-free
    // 64-bit stores.  2 loads per thread per row; row-invariant address parts hoisted, three rows in flight.
    {
        const int lane = tid & 31, wid = tid >> 5;
        const int la = lane >> 2, lj = lane & 3;
        const size_t row_f4 = (size_t)D * A / 2;                                    // float4 per RDS range row
        const float4* src[2];
        float* dst[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int cb = (wid + 8 * k) * 8 + 2 * lj;                              // first cell of this lane's pair
            src[k] = reinterpret_cast<const float4*>(frame + ((size_t)(a0 + la)) * D + d0 + cb);
            const int cell = cb + (la & 1);                                         // even antennas keep the first cell
            dst[k] = pw + a8_off(0, cell + 1, la >> 2) + (la & 2);
        }
        // The six row groups are software pipelined through two register buffers: the loads of group g + 1 are in
        // flight while group g is squared and stored, so the CTA waits for HBM once instead of six times.
        constexpr int NR = (A8_TR + 2) / 6;
        float4 v[2][NR][2];
        auto fetch = [&](int g, float4 (&w)[NR][2]) {
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                const int r = r0 - 1 + g * NR + q;
                const bool ok = r >= 0 && r < R;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    w[q][k] = ok ? __ldg(src[k] + (size_t)r * row_f4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0, v[0]);
#pragma unroll
        for (int g = 0; g < 6; ++g) {
            if (g + 1 < 6) fetch(g + 1, v[(g + 1) & 1]);
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                const int rr = g * NR + q, r = r0 - 1 + rr;
                const bool ok = r >= 0 && r < R;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float4 x = v[g & 1][q][k];
                    const float p0 = ok ? fmaf(x.x, x.x, x.y * x.y) : -1.f;   // cell cb
                    const float p1 = ok ? fmaf(x.z, x.z, x.w * x.w) : -1.f;   // cell cb + 1
                    const float got = __shfl_xor_sync(0xffffffffu, (la & 1) ? p0 : p1, 4);            // antenna a ^ 1
                    const float2 pr = (la & 1) ? make_float2(got, p1) : make_float2(p0, got);
                    *reinterpret_cast<float2*>(dst[k] + rr * A8_W) = pr;
                }
            }
        }
    }
    __syncthreads();

    // ---- walk: thread = (doppler bin, row half); 8 antennas per cell
    const int ddp = (tid & (A8_TD - 1)) + 1, rh = tid >> 7;
    const int rbase = rh * A8_HALF;                  // tile rows rbase+1 .. rbase+8 (1-based incl. halo)
    uint32_t hit[2] = {0u, 0u}, near[2] = {0u, 0u}, cand[2] = {0u, 0u}, unc[2] = {0u, 0u};  // byte per row: antenna mask
    int my_near = 0;
    float my_psum = 0.f;
    const float band = 2.f * eps;
    // the two antenna quads of the cell are walked one after the other (half the live registers of an
    // 8-wide walk, so three CTAs fit per SM); their hit bits land in the same per-row antenna mask
#pragma unroll
    for (int quad = 0; quad < 2; ++quad) {
        auto ld4 = [&](int rr, int dp, float (&o)[4]) {
            const float4 v0 = *reinterpret_cast<const float4*>(pw + a8_off(rr, dp, quad));
            o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v0.w;
        };
        float l1[4], c1[4], q1[4], hp[4];
        {
            float l0[4], c0[4], q0[4];
            ld4(rbase, ddp - 1, l0); ld4(rbase, ddp, c0); ld4(rbase, ddp + 1, q0);
#pragma unroll
            for (int j = 0; j < 4; ++j) hp[j] = fmaxf(fmaxf(l0[j], c0[j]), q0[j]);
        }
        ld4(rbase + 1, ddp - 1, l1); ld4(rbase + 1, ddp, c1); ld4(rbase + 1, ddp + 1, q1);
#pragma unroll
        for (int i = 0; i < A8_HALF; ++i) {
            const int rr = rbase + 1 + i;
            float l2[4], c2[4], q2[4];
            ld4(rr + 1, ddp - 1, l2); ld4(rr + 1, ddp, c2); ld4(rr + 1, ddp + 1, q2);
            const int r = r0 + rr - 1;
            const bool row_ok = r < R && gate[r < R ? r : 0];
            uint32_t hm = 0u, um = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float hn = fmaxf(fmaxf(l2[j], c2[j]), q2[j]);
                const float m = fmaxf(fmaxf(hp[j], hn), fmaxf(l1[j], q1[j]));
                const float c = c1[j];
                my_psum += c;
                // three zones from two scaled copies of c: surely a detection (margin > 2 eps to the best neighbour and to
                // the threshold), surely not one, and the thin band in between where the exact rule of classify() decides
                const float cu = fmaf(c, band, c), cl = fmaf(c, -band, c);
                const bool maybe = row_ok && cu >= m && cu > thr;
                const bool sure = cl >= m && cl > thr;
                hm |= maybe ? (1u << j) : 0u;
                um |= (maybe && !sure) ? (1u << j) : 0u;          // ~1e-5 of the cells: settled after the walk
                hp[j] = fmaxf(fmaxf(l1[j], c), q1[j]);
                l1[j] = l2[j]; c1[j] = c2[j]; q1[j] = q2[j];
            }
            const int sh = (i & 3) * 8 + quad * 4;
            hit[i >> 2] |= hm << sh;
            unc[i >> 2] |= um << sh;
        }
    }
    // cells inside the band: the exact rule on the nine powers, re-read from the tile
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        uint32_t u = unc[w];
        while (u) {
            const int b = __ffs(u) - 1;
            u &= u - 1;
            const int i = w * 4 + (b >> 3), j = b & 7;
            const int rr = rbase + 1 + i;
            auto at = [&](int r_, int d_) { return pw[a8_off(r_, d_, j >> 2) + (j & 3)]; };
            const float c = at(rr, ddp);
            float m = fmaxf(fmaxf(at(rr - 1, ddp - 1), at(rr - 1, ddp)), at(rr - 1, ddp + 1));
            m = fmaxf(m, fmaxf(at(rr, ddp - 1), at(rr, ddp + 1)));
            m = fmaxf(m, fmaxf(fmaxf(at(rr + 1, ddp - 1), at(rr + 1, ddp)), at(rr + 1, ddp + 1)));
            const int cls = rs_classify(c, m, thr, eps);
            if (cls == 0) hit[w] &= ~(1u << b);
            if (cls & 2) { near[w] |= 1u << b; ++my_near; }
            if (cls & 4) cand[w] |= 1u << b;
        }
    }
    a8_emit(hit, near, cand, my_near, my_psum, out, f, r0, d0, a0, rbase, ddp,
            [&](int rr, int j) { return pw[a8_off(rr, ddp, j >> 2) + (j & 3)]; });
}

// ---------------------------------------------------------------------------------------------
// Compaction of the hit masks the fused 2-D FFT kernel wrote (rs_detect_fused.cuh, rs_fft2d_ws.cu): the same segments,
// order, leaders and counters as detect_a8_kernel, from 2 KB of mask words per tile instead of the tile's 144 KB of RDS.
// Same thread <-> cell assignment as the walk above: thread = (Doppler bin, half of the 16 rows) = one row group of the
// masks, all 8 antennas of the octet.  Byte (d & 3) of antenna j's word is the thread's 8-row column on that antenna; an
// 8 x 8 bit transpose turns the eight columns into the 64-bit mask "byte i = antennas that flagged row i", whose set bits
// in ascending order ARE the thread's entries in list order.  The list is assembled in shared memory and leaves as
// coalesced rows (a thread's ~6 entries sit at consecutive positions: direct stores would touch 32 sectors per
// instruction).  det_power (optional) is gathered from the RDS for the entries only.
// ---------------------------------------------------------------------------------------------
constexpr int EMIT_STAGE = 2048;          // staged entries per tile (a tile of the benchmark scene holds ~1 500); the rest go direct

__device__ __forceinline__ unsigned long long transpose8x8(unsigned long long x) {     // bit (8 r + c) <-> bit (8 c + r)
    unsigned long long t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x ^= t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x ^= t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x ^= t ^ (t << 28);
    return x;
}

__global__ void __launch_bounds__(DET_THREADS)
compact_masks_kernel(const float2* __restrict__ rds, FusedDetectMasks fd, DetOut out, int R, int D, int A, Tiling tl) {
    __shared__ int warp_sums[DET_THREADS / 32];
    __shared__ int total_s;
    __shared__ uint32_t s_key[EMIT_STAGE], s_lead[EMIT_STAGE];
    __shared__ __align__(16) uint8_t s_flags[EMIT_STAGE];
    const int tile = blockIdx.x % tl.ntiles;
    const int f = blockIdx.x / tl.ntiles;
    const int ia = tile % tl.nac;
    const int ir = tile / tl.nac;                    // ntd == 1: a tile spans the Doppler axis
    const int a0 = ia * A8_AC;
    const int tid = threadIdx.x;
    const int d = tid & (A8_TD - 1), rh = tid >> 7;
    const int row0 = ir * A8_TR + rh * A8_HALF;      // first of this thread's 8 range bins
    const int sh = 8 * (d & 3);
    unsigned long long cols = 0ull, ncols = 0ull, ccols = 0ull;        // byte j: the 8-row column of antenna a0 + j
    float my_psum = 0.f;
    bool any_near = false;
#pragma unroll
    for (int j = 0; j < A8_AC; ++j) {
        const size_t gi = ((size_t)f * A + a0 + j) * FD_GROUPS + 2 * ir + rh;
        const uint32_t w = __ldg(fd.hit + gi * FD_WORDS + (d >> 2));
        cols |= (unsigned long long)((w >> sh) & 0xFFu) << (8 * j);
        const float2 rec = __ldg(fd.rec + gi);
        if (__float_as_int(rec.y)) {                 // warp-uniform: this row group has cells inside the guard band
            any_near = true;
            ncols |= (unsigned long long)((__ldg(fd.near + gi * FD_WORDS + (d >> 2)) >> sh) & 0xFFu) << (8 * j);
            ccols |= (unsigned long long)((__ldg(fd.cand + gi * FD_WORDS + (d >> 2)) >> sh) & 0xFFu) << (8 * j);
        }
        if (d == 0) my_psum += rec.x;                // threads 0 and 128: the 16 row-group sums of the tile, fixed order
    }
    const unsigned long long m64 = transpose8x8(cols);                 // byte i: antennas that flagged row row0 + i
    unsigned long long n64 = 0ull, c64 = 0ull;
    if (any_near) { n64 = transpose8x8(ncols); c64 = transpose8x8(ccols); }
    // entries and leaders (= rows with a non-zero byte) of this thread, scanned together (entries <= 2^14 per tile)
    unsigned long long nz = m64 | (m64 >> 4);
    nz |= nz >> 2;
    nz |= nz >> 1;
    nz &= 0x0101010101010101ull;
    const int packed = block_exclusive_scan(__popcll(m64) | (__popcll(nz) << 16), warp_sums, &total_s);
    const int total = total_s & 0xFFFF, total_lead = total_s >> 16;
    const size_t seg = (size_t)blockIdx.x;
    int pos = packed & 0xFFFF, lpos = packed >> 16;
    if (tid == 0) {
        out.count[seg] = total < out.seg_cap ? total : out.seg_cap;
        if (total > out.seg_cap) out.overflow[f] = 1;
    }
    {
        __shared__ int near_s;
        __shared__ float psum_s;
        block_sums(__popcll(n64), my_psum, &near_s, &psum_s);
        if (tid == 0) {
            if (out.nnear) out.nnear[seg] = near_s;
            if (out.psum) out.psum[seg] = psum_s;
        }
    }
    const float2* frame = rds + (size_t)f * R * D * A;
    uint32_t* gk = out.key + seg * out.seg_cap;
    uint32_t* gl = out.lead + seg * out.seg_cap;
    uint8_t* gf = out.flags + seg * out.seg_cap;
    const uint32_t key0 = rs_make_key(a0, row0, d);
    int kept = 0, last_row = -1;
    unsigned long long m = m64;
    while (m) {                                      // set bits in ascending order = rows ascending, antennas ascending
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int i = b >> 3, j = b & 7;
        if (i != last_row) {                         // first entry of a cell: its leader
            last_row = i;
            const int k = __popc((uint32_t)(m64 >> (8 * i)) & 0xFFu);
            if (pos + k <= out.seg_cap) {            // a leader is kept only if all of its cell's entries fit
                const uint32_t lv = (uint32_t)pos | ((uint32_t)k << 16);
                if (lpos < EMIT_STAGE) s_lead[lpos] = lv;
                else gl[lpos] = lv;
                ++kept;
            }
            ++lpos;
        }
        if (pos < out.seg_cap) {
            const uint32_t kv = key0 + ((uint32_t)j << 24) + ((uint32_t)i << 12);
            const uint8_t fv = ((c64 >> b) & 1ull) ? (RS_FLAG_NEARMAX | RS_FLAG_DROPPED) : ((n64 >> b) & 1ull) ? RS_FLAG_NEARMAX : 0;
            if (pos < EMIT_STAGE) {
                s_key[pos] = kv;
                s_flags[pos] = fv;
            } else {
                gk[pos] = kv;
                gf[pos] = fv;
            }
            if (out.power) {
                const float2 x = __ldg(frame + ((size_t)(row0 + i) * A + a0 + j) * D + d);
                out.power[seg * out.seg_cap + pos] = fmaf(x.x, x.x, x.y * x.y);
            }
        }
        ++pos;
    }
    __syncthreads();
    {
        const int n = min(min(total, out.seg_cap), EMIT_STAGE), nl = min(total_lead, EMIT_STAGE);
        for (int i = tid; i < n; i += DET_THREADS) gk[i] = s_key[i];
        for (int i = tid; i < nl; i += DET_THREADS) gl[i] = s_lead[i];      // slots beyond the kept leaders are never read
        if (out.seg_cap % 4 == 0) {
            for (int i = tid; i < (n + 3) / 4; i += DET_THREADS)            // the tail bytes of the last word lie beyond `count`
                reinterpret_cast<uint32_t*>(gf)[i] = reinterpret_cast<const uint32_t*>(s_flags)[i];
        } else {
            for (int i = tid; i < n; i += DET_THREADS) gf[i] = s_flags[i];
        }
    }
    // number of kept leaders: all of them unless the segment overflowed (then a block sum, uniform branch)
    if (total <= out.seg_cap) {
        if (tid == 0) out.nlead[seg] = total_lead;
    } else {
        __syncthreads();
        block_exclusive_scan(kept, warp_sums, &total_s);
        if (tid == 0) out.nlead[seg] = total_s;
    }
}

}  // namespace

extern "C" int rs_detect_tiling(int R, int D, int A, int* tile_r, int* tile_d, int* ntiles) {
    RS_CHECK_ARG(R > 0 && D > 0 && A > 0 && R <= RS_MAX_RANGE_BINS && D <= RS_MAX_DOPPLER_BINS && A <= RS_MAX_ANTENNAS,
                 "rs_detect_tiling: bad dims");
    Tiling t = make_tiling(R, D, A);
    if (tile_r) *tile_r = t.TR;
    if (tile_d) *tile_d = t.TD;
    if (ntiles) *ntiles = t.ntiles;
    return RS_OK;
}

// frames [f0, f0 + nf) of a batch through the stand-alone kernels (the overflow flags are cleared by the caller)
static int launch_detect(const float2* rds, const uint8_t* gate, float thr, float eps, DetOut out, int f0, int nf, int R, int D,
                         int A, cudaStream_t stream, const char* what) {
    Tiling t = make_tiling(R, D, A);
    const long long blocks = (long long)nf * t.ntiles;
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_detect: too many blocks");
    if (nf <= 0) return RS_OK;
    rds += (size_t)f0 * R * D * A;
    const size_t s0 = (size_t)f0 * t.ntiles, e0 = s0 * out.seg_cap;
    out.key += e0;
    if (out.power) out.power += e0;
    out.flags += e0;
    out.lead += e0;
    out.count += s0;
    out.nlead += s0;
    out.overflow += f0;
    if (out.nnear) out.nnear += s0;
    if (out.psum) out.psum += s0;
    if (A % 8 == 0 && D % A8_TD == 0 && R % A8_TR == 0 && t.TD == A8_TD && t.AC == A8_AC && t.TR == A8_TR) {
        const size_t smem = (size_t)(A8_TR + 2) * A8_W * sizeof(float);
        cudaFuncSetAttribute(detect_a8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        detect_a8_kernel<<<(unsigned)blocks, DET_THREADS, smem, stream>>>(rds, gate, thr, eps, out, R, D, A, t);
        RS_CHECK_LAUNCH(what);
        return RS_OK;
    }
    const size_t smem = (size_t)(t.TR + 2) * (t.TD + 2) * t.AC * sizeof(float);
    if (smem > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_detect: tile needs %zu B of shared memory", smem);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    detect_kernel<<<(unsigned)blocks, DET_THREADS, smem, stream>>>(rds, gate, thr, eps, out, R, D, A, t);
    RS_CHECK_LAUNCH(what);
    return RS_OK;
}

extern "C" int rs_detect(const void* rds, const uint8_t* range_gate, float thr_power, float det_eps, uint32_t* det_key,
                         float* det_power, uint8_t* det_flags, uint32_t* det_lead, int32_t* det_count,
                         int32_t* det_nlead, int32_t* det_overflow, int32_t* det_nnear, float* det_psum, int seg_cap,
                         int F, int R, int D, int A, void* stream) {
    RS_CHECK_ARG(rds && range_gate && det_key && det_power && det_flags && det_lead && det_count && det_nlead &&
                     det_overflow,
                 "rs_detect: null pointer");
    RS_CHECK_ARG(F > 0 && R > 0 && D > 0 && A > 0 && R <= RS_MAX_RANGE_BINS && D <= RS_MAX_DOPPLER_BINS &&
                     A <= RS_MAX_ANTENNAS && seg_cap > 0 && seg_cap <= 65535,
                 "rs_detect: bad dims (seg_cap must be in 1..65535)");
    DetOut out{det_key, det_power, det_flags, det_lead, det_count, det_nlead, det_overflow, det_nnear, det_psum, seg_cap};
    cudaMemsetAsync(det_overflow, 0, sizeof(int32_t) * F, (cudaStream_t)stream);
    return launch_detect((const float2*)rds, range_gate, thr_power, det_eps, out, 0, F, R, D, A, (cudaStream_t)stream, "rs_detect");
}

// ---------------------------------------------------------------------------------------------
// 2-D FFT + detection in one call (rows a3-a8): the detection rides in the Doppler phase of the persistent FFT kernel
// ---------------------------------------------------------------------------------------------
int rs_range_doppler_fft_impl(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                              void* mid_ws, void* rds, int F, int A, int C_total, int chirp0, int C_used, int S,
                              int dc_removal, void* stream, FusedDetectReq* req);

extern "C" long long rs_fused_detect_ws_bytes(int F, int A) {
    return (F > 0 && A > 0) ? (long long)((size_t)F * A * FD_BYTES_PER_PLANE) : 0;
}

namespace {
struct SideDetectCtx {
    const float2* rds;
    const uint8_t* gate;
    float thr, eps;
    DetOut out;
    int R, D, A, rc;
};
void side_detect_hook(void* ctx_, int f0, int nf, cudaStream_t st) {
    SideDetectCtx* c = (SideDetectCtx*)ctx_;
    c->rc = launch_detect(c->rds, c->gate, c->thr, c->eps, c->out, f0, nf, c->R, c->D, c->A, st, "rs_range_doppler_detect(side)");
}
}  // namespace

extern "C" int rs_range_doppler_detect(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                                       void* mid_ws, void* rds, void* fused_ws, int F, int A, int C_total, int chirp0,
                                       int C_used, int S, int dc_removal, const uint8_t* range_gate, float thr_power,
                                       float det_eps, uint32_t* det_key, float* det_power, uint8_t* det_flags,
                                       uint32_t* det_lead, int32_t* det_count, int32_t* det_nlead, int32_t* det_overflow,
                                       int32_t* det_nnear, float* det_psum, int seg_cap, void* stream) {
    RS_CHECK_ARG(cube && table && twiddle_s && twiddle_c && rds && range_gate && det_key && det_flags && det_lead &&
                     det_count && det_nlead && det_overflow,
                 "rs_range_doppler_detect: null pointer");
    RS_CHECK_ARG(F > 0 && A > 0 && A <= RS_MAX_ANTENNAS && S > 0 && S <= RS_MAX_RANGE_BINS && C_used > 0 &&
                     C_used <= RS_MAX_DOPPLER_BINS && chirp0 >= 0 && chirp0 + C_used <= C_total && seg_cap > 0 && seg_cap <= 65535,
                 "rs_range_doppler_detect: bad dims (seg_cap must be in 1..65535)");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = S, D = C_used;
    DetOut out{det_key, det_power, det_flags, det_lead, det_count, det_nlead, det_overflow, det_nnear, det_psum, seg_cap};
    cudaMemsetAsync(det_overflow, 0, sizeof(int32_t) * F, st);
    const char* env = getenv("RS_FUSED_DETECT");           // 0: the two stages one after the other
    Tiling t = make_tiling(R, D, A);
    const bool fused = fused_ws && S == 256 && C_used == 128 && A % 8 == 0 && t.TD == A8_TD && t.AC == A8_AC && t.TR == A8_TR &&
                       !(env && atoi(env) == 0);
    if (!fused) {
        RS_CHECK_ARG(det_power, "rs_range_doppler_detect: det_power is optional only on the fused path");
        int rc = rs_range_doppler_fft_impl(cube, table, twiddle_s, twiddle_c, mid_ws, rds, F, A, C_total, chirp0, C_used, S,
                                           dc_removal, stream, nullptr);
        if (rc != RS_OK) return rc;
        return launch_detect((const float2*)rds, range_gate, thr_power, det_eps, out, 0, F, R, D, A, st, "rs_range_doppler_detect");
    }
    const size_t planes = (size_t)F * A;
    uint32_t* words = (uint32_t*)fused_ws;
    FusedDetectReq req;
    req.masks.hit = words;
    req.masks.near = words + planes * FD_GROUPS * FD_WORDS;
    req.masks.cand = words + 2 * planes * FD_GROUPS * FD_WORDS;
    req.masks.rec = (float2*)(words + 3 * planes * FD_GROUPS * FD_WORDS);
    req.masks.gate = range_gate;
    req.masks.thr = thr_power;
    req.masks.thrn = nextafterf(thr_power, INFINITY);
    req.masks.eps = det_eps;
    { const char* dbg = getenv("RS_FD_DBG"); req.masks.dbg = dbg ? atoi(dbg) : 0; }
    SideDetectCtx ctx{(const float2*)rds, range_gate, thr_power, det_eps, out, R, D, A, RS_OK};
    req.side_hook = side_detect_hook;
    req.ctx = &ctx;
    req.frames_masked = 0;
    int rc = rs_range_doppler_fft_impl(cube, table, twiddle_s, twiddle_c, mid_ws, rds, F, A, C_total, chirp0, C_used, S,
                                       dc_removal, stream, &req);
    if (rc != RS_OK) return rc;
    if (ctx.rc != RS_OK) return ctx.rc;
    const int Fm = req.frames_masked;
    const char* nc = getenv("RS_FD_NO_COMPACT");          // timing probe only: leave the lists unwritten
    if (Fm > 0 && !(nc && atoi(nc) == 1)) {
        compact_masks_kernel<<<(unsigned)((long long)Fm * t.ntiles), DET_THREADS, 0, st>>>((const float2*)rds, req.masks, out, R, D, A, t);
        RS_CHECK_LAUNCH("rs_range_doppler_detect(compact)");
    }
    if (Fm == 0)          // the fused kernel was not available: the side hook did not run either
        return launch_detect((const float2*)rds, range_gate, thr_power, det_eps, out, 0, F, R, D, A, st, "rs_range_doppler_detect");
    return RS_OK;
}
