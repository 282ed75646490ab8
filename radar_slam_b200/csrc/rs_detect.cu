// Peak detection on the cell-major RDS.
//
// Replaces SignalPreprocessor.extract_range_doppler_peaks (dechirp.py:215-278).  The reference
// compares dB values; 10 log10(p + 1e-12) is monotone in p, so the local-maximum test and the
// threshold are evaluated on the linear power p = |X|^2 (host converts the threshold in fp64).
// scipy.ndimage.maximum_filter(size=3) with its default 'reflect' boundary duplicates the edge
// row/column, i.e. out-of-range neighbours are ignored; ties count as maxima (== test).
//
// A CTA owns a tile of TR range bins x TD Doppler bins x AC antennas (+1 halo), computes the
// power plane into shared memory, walks every (doppler, antenna) column with a 3-row sliding
// window in registers, and compacts the hits of the tile into its own fixed-capacity segment with
// a block-wide prefix sum (deterministic order, no atomics).
#include "rs_common.cuh"

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

__global__ void __launch_bounds__(DET_THREADS)
detect_kernel(const float2* __restrict__ rds, const uint8_t* __restrict__ gate, float thr, float eps,
              uint32_t* __restrict__ det_key, float* __restrict__ det_power, uint8_t* __restrict__ det_flags,
              int32_t* __restrict__ det_count, int32_t* __restrict__ det_overflow, int seg_cap, int R, int D, int A,
              Tiling tl) {
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
            const float2 x = __ldg(frame + ((size_t)r * D + d) * A + a);
            p = fmaf(x.x, x.x, x.y * x.y);
        }
        pw[i] = p;
    }
    __syncthreads();

    // ---- column walk: masks of hits / near-ties per column, kept in registers
    const int ncols = TD * AC;
    uint32_t hit[DET_MAX_COLS_PER_THREAD], near[DET_MAX_COLS_PER_THREAD];
    int my_count = 0;
#pragma unroll
    for (int q = 0; q < DET_MAX_COLS_PER_THREAD; ++q) {
        hit[q] = 0u;
        near[q] = 0u;
        const int col = threadIdx.x + q * DET_THREADS;
        if (col < ncols) {
            const int dd = col / AC + 1, ac = col - (col / AC) * AC;
            const int d = d0 + dd - 1, a = a0 + ac;
            if (d < D && a < A) {
                const float* base = pw + dd * AC + ac;
                // h = max over the 3 horizontal neighbours of a row; l/c/r of the current row kept separately
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
                        const bool is_max = c1 >= m;
                        const bool above = c1 > thr;
                        if (is_max && above && gate[r]) {
                            hit[q] |= 1u << (rr - 1);
                            ++my_count;
                            if ((c1 - m) <= eps * c1 || (c1 - thr) <= eps * fabsf(thr)) near[q] |= 1u << (rr - 1);
                        }
                    }
                    h_prev = fmaxf(fmaxf(l1, c1), rt1);
                    l1 = l2; c1 = c2; rt1 = rt2;
                }
            }
        }
    }

    // ---- deterministic compaction into this tile's segment
    const int offset = block_exclusive_scan(my_count, warp_sums, &total_s);
    const int total = total_s;
    const size_t seg = (size_t)blockIdx.x;
    if (threadIdx.x == 0) {
        det_count[seg] = total < seg_cap ? total : seg_cap;
        if (total > seg_cap) det_overflow[f] = 1;
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
                if (pos < seg_cap) {
                    const size_t o = seg * seg_cap + pos;
                    det_key[o] = rs_make_key(a0 + ac, r0 + b, d0 + dd - 1);
                    det_power[o] = pw[(b + 1) * W + dd * AC + ac];
                    det_flags[o] = (near[q] >> b) & 1u ? RS_FLAG_NEARMAX : 0;
                }
                ++pos;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path: antenna chunks of 8 (A % 8 == 0), tile = TR range bins x 128 Doppler bins x 8 antennas.
// 256 threads; thread = (doppler bin, antenna quad).  Loads are float4 (two antennas), the power plane
// is [row][doppler+1][8] floats so a thread reads its 4 antennas and both Doppler neighbours with three
// LDS.128 per row step; ~20 instructions per cell instead of ~150 in the generic kernel.
// ---------------------------------------------------------------------------------------------
template <int TR>
__global__ void __launch_bounds__(DET_THREADS)
detect_a8_kernel(const float2* __restrict__ rds, const uint8_t* __restrict__ gate, float thr, float eps,
                 uint32_t* __restrict__ det_key, float* __restrict__ det_power, uint8_t* __restrict__ det_flags,
                 int32_t* __restrict__ det_count, int32_t* __restrict__ det_overflow, int seg_cap, int R, int D, int A,
                 Tiling tl) {
    constexpr int TD = 128, AC = 8, W = (TD + 2) * AC;
    extern __shared__ float pw[];   // [(TR+2)][W]
    __shared__ int warp_sums[DET_THREADS / 32];
    __shared__ int total_s;

    const int tile = blockIdx.x % tl.ntiles;
    const int f = blockIdx.x / tl.ntiles;
    const int ia = tile % tl.nac;
    const int id = (tile / tl.nac) % tl.ntd;
    const int ir = tile / (tl.nac * tl.ntd);
    const int r0 = ir * TR, d0 = id * TD, a0 = ia * AC;
    const float2* frame = rds + (size_t)f * R * D * A;
    const int tid = threadIdx.x;

    // ---- halo columns (only when the frame is wider than one tile) and out-of-range fill
    for (int i = tid; i < (TR + 2) * 2 * AC; i += DET_THREADS) {
        const int rr = i / (2 * AC), rem = i - rr * 2 * AC;
        const int side = rem / AC, ac = rem - side * AC;
        const int r = r0 - 1 + rr, d = side ? d0 + TD : d0 - 1;
        float p = -1.f;
        if (r >= 0 && r < R && d >= 0 && d < D) {
            const float2 x = __ldg(frame + ((size_t)r * D + d) * A + a0 + ac);
            p = fmaf(x.x, x.x, x.y * x.y);
        }
        pw[rr * W + (side ? (TD + 1) * AC : 0) + ac] = p;
    }
    // ---- interior: each row is TD cells x 4 float4 (= 2 antennas each); 2 float4 per thread per row
#pragma unroll 3
    for (int rr = 0; rr < TR + 2; ++rr) {
        const int r = r0 - 1 + rr;
        const bool ok = r >= 0 && r < R;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i4 = tid + DET_THREADS * k;
            const int cell = i4 >> 2, part = i4 & 3;
            float2 p2 = make_float2(-1.f, -1.f);
            if (ok) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(frame + ((size_t)r * D + d0 + cell) * A + a0) + part);
                p2.x = fmaf(v.x, v.x, v.y * v.y);
                p2.y = fmaf(v.z, v.z, v.w * v.w);
            }
            *reinterpret_cast<float2*>(pw + rr * W + (cell + 1) * AC + part * 2) = p2;
        }
    }
    __syncthreads();

    // ---- column walk: thread = (doppler bin dd, antenna quad aq)
    const int dd = (tid >> 1) + 1, aq = (tid & 1) * 4;
    uint32_t hit[4] = {0u, 0u, 0u, 0u}, near[4] = {0u, 0u, 0u, 0u};
    int my_count = 0;
    {
        const float* base = pw + dd * AC + aq;
        auto ld4 = [&](const float* p, float (&o)[4]) {
            const float4 v = *reinterpret_cast<const float4*>(p);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        };
        float l0[4], c0[4], q0[4], l1[4], c1[4], q1[4], hp[4];
        ld4(base - AC, l0); ld4(base, c0); ld4(base + AC, q0);
        ld4(base + W - AC, l1); ld4(base + W, c1); ld4(base + W + AC, q1);
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = fmaxf(fmaxf(l0[j], c0[j]), q0[j]);
#pragma unroll
        for (int rr = 1; rr <= TR; ++rr) {
            float l2[4], c2[4], q2[4];
            const float* nx = base + (rr + 1) * W;
            ld4(nx - AC, l2); ld4(nx, c2); ld4(nx + AC, q2);
            const int r = r0 + rr - 1;
            const bool row_ok = r < R && gate[r < R ? r : 0];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float hn = fmaxf(fmaxf(l2[j], c2[j]), q2[j]);
                const float m = fmaxf(fmaxf(hp[j], hn), fmaxf(l1[j], q1[j]));
                const float c = c1[j];
                if (row_ok && c >= m && c > thr) {
                    hit[j] |= 1u << (rr - 1);
                    ++my_count;
                    if ((c - m) <= eps * c || (c - thr) <= eps * fabsf(thr)) near[j] |= 1u << (rr - 1);
                }
                hp[j] = fmaxf(fmaxf(l1[j], c), q1[j]);
                l1[j] = l2[j]; c1[j] = c2[j]; q1[j] = q2[j];
            }
        }
    }

    const int offset = block_exclusive_scan(my_count, warp_sums, &total_s);
    const int total = total_s;
    const size_t seg = (size_t)blockIdx.x;
    if (tid == 0) {
        det_count[seg] = total < seg_cap ? total : seg_cap;
        if (total > seg_cap) det_overflow[f] = 1;
    }
    int pos = offset;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t m = hit[j];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (pos < seg_cap) {
                const size_t o = seg * seg_cap + pos;
                det_key[o] = rs_make_key(a0 + aq + j, r0 + b, d0 + dd - 1);
                det_power[o] = pw[(b + 1) * W + dd * AC + aq + j];
                det_flags[o] = (near[j] >> b) & 1u ? RS_FLAG_NEARMAX : 0;
            }
            ++pos;
        }
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

extern "C" int rs_detect(const void* rds, const uint8_t* range_gate, float thr_power, float det_eps, uint32_t* det_key,
                         float* det_power, uint8_t* det_flags, int32_t* det_count, int32_t* det_overflow, int seg_cap,
                         int F, int R, int D, int A, void* stream) {
    RS_CHECK_ARG(rds && range_gate && det_key && det_power && det_flags && det_count && det_overflow,
                 "rs_detect: null pointer");
    RS_CHECK_ARG(F > 0 && R > 0 && D > 0 && A > 0 && R <= RS_MAX_RANGE_BINS && D <= RS_MAX_DOPPLER_BINS &&
                     A <= RS_MAX_ANTENNAS && seg_cap > 0,
                 "rs_detect: bad dims");
    Tiling t = make_tiling(R, D, A);
    const size_t smem = (size_t)(t.TR + 2) * (t.TD + 2) * t.AC * sizeof(float);
    if (smem > (size_t)rs_smem_optin_limit()) {
        rs_set_error("rs_detect: tile needs %zu B of shared memory", smem);
        return RS_ECAPACITY;
    }
    cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long blocks = (long long)F * t.ntiles;
    RS_CHECK_ARG(blocks < (1ll << 31), "rs_detect: too many blocks");
    cudaMemsetAsync(det_overflow, 0, sizeof(int32_t) * F, (cudaStream_t)stream);
    if (A % 8 == 0 && D % 128 == 0 && t.TD == 128 && t.AC == 8 && t.TR == 16 && R % 16 == 0) {
        cudaFuncSetAttribute(detect_a8_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        detect_a8_kernel<16><<<(unsigned)blocks, DET_THREADS, smem, (cudaStream_t)stream>>>(
            (const float2*)rds, range_gate, thr_power, det_eps, det_key, det_power, det_flags, det_count, det_overflow,
            seg_cap, R, D, A, t);
        RS_CHECK_LAUNCH("rs_detect(a8)");
        return RS_OK;
    }
    detect_kernel<<<(unsigned)blocks, DET_THREADS, smem, (cudaStream_t)stream>>>(
        (const float2*)rds, range_gate, thr_power, det_eps, det_key, det_power, det_flags, det_count, det_overflow,
        seg_cap, R, D, A, t);
    RS_CHECK_LAUNCH("rs_detect");
    return RS_OK;
}
