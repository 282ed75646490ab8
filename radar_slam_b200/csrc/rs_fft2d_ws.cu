// K12 v2: the fused dechirp + window + range FFT + Doppler FFT of 256 x 128 planes as a PERSISTENT, WARP-SPECIALISED
// 4-CTA cluster kernel fed by TMA (dechirp.py:143-213).
//
// Round 1's cluster kernel (rs_fft2d.cu) ran the two phases of a plane one after the other in every CTA: raw chirps
// only flowed from HBM during the range phase, RDS rows only during the Doppler phase, and a cluster barrier stood
// between them (0.50 of the HBM peak; long-scoreboard / barrier / MEMBAR stalls).  Here the two phases of
// CONSECUTIVE planes run side by side on every SM:
//
//   producer        one thread of the range group streams the CTA's 32 chirps of plane n+1 with cp.async.bulk.tensor
//                   (TMA, evict-first) into a two-slot ring of 16-chirp stages (32 KB each), full/empty mbarriers
//   range group     8 warps: stage -> registers (conj(ref)*window column held in registers) -> radix-16 pass 1,
//                   written back IN PLACE in the stage through a per-chirp rotation -> radix-16 pass 2 -> range bin p
//                   of chirp c goes to M[n&1][p mod 64][c] of the CTA that owns range bins [64r, 64r+64)
//                   (st.shared::cluster, 128-byte runs along chirps), then one release-arrive per warp on the four
//                   owners' "plane full" mbarriers
//   Doppler group   8 warps: wait for "plane full" of buffer n&1, each warp owns 8 range rows: radix-16 pass 1 in
//                   place (rotation keeps pass 2 conflict free), radix-8 pass 2, Doppler fftshift, rows stored either
//                   straight from registers (default) or written back to the row and shipped with a bulk (TMA) store
//                   (RS_K12_STORE=tma), then one release-arrive per warp on the four CTAs' "buffer empty" mbarriers
//
// The plane quarter M is double buffered (2 x 68 KB), so the range group of plane n+1 never waits for the Doppler
// group of plane n, a slow CTA of the cluster is absorbed by the buffer instead of a barrier, and loads and stores
// are in flight all the time.  No __syncthreads and no barrier.cluster inside the plane loop: one 256-thread named
// barrier per stage in the range group, __syncwarp in the Doppler group, mbarriers across CTAs.
// HBM traffic stays at the algorithmic minimum, 8 B in + 8 B out per cell.
// (The default variant streams 8-chirp stages through a four-slot ring, two independent range subgroups, and hands the
// range bins over with st.async on a transaction mbarrier; the list above describes variant 0.  RS_K12_VARIANT.)
//
// DETECT (rs_range_doppler_detect): the Doppler group also runs the peak detection of dechirp.py:235-263 on the plane it
// holds -- see the comment on the kernel template and rs_detect_fused.cuh -- so the RDS is never read by a detection pass.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <cmath>
#include "rs_common.cuh"
#include "rs_detect_fused.cuh"
#include "rs_fft_pow2.cuh"

namespace ws {

constexpr int S = 256, C = 128, NC = 4;
constexpr int CPC = C / NC;             // chirps one CTA transforms per plane
constexpr int MAXSTAGE = 4;             // stages per plane == ring slots: CPC / CB with CB = 16 or 8 chirps per stage
constexpr int ROWS = S / NC;            // range bins owned by one CTA
constexpr int MP = C + 8;               // row pitch of M (complex): 8 (mod 16)
constexpr int R_THREADS = 256, D_THREADS = 256;
constexpr int THREADS = R_THREADS + D_THREADS;       // 16 warps x 128 registers fill the register file

struct __align__(128) Smem {
    float2 stage[CPC * S];              // TMA destination: CPC / CB slots of CB chirps, range passes run in place
    float2 M[2][ROWS * MP];             // this CTA's range bins x all chirps, double buffered over planes
    float2 tw1s[S];                     // range inter-pass twiddles w_S^{k1 n2}
    float2 tw1c[C];                     // Doppler inter-pass twiddles
    float halo[2][2][C];                // DETECT: |X|^2 of range rows -1 / ROWS (the neighbour CTAs' edge rows), per M buffer
    unsigned long long full_ld[MAXSTAGE], empty_ld[MAXSTAGE], full_M[2], empty_M[2], halo_full[2];
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {     // release at cluster scope
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error must end as a trapped launch (RS_ECUDA), never as a hung GPU.  Waits are acquire.cta:
// what they guard lives in THIS CTA's shared memory (written by TMA, st.async or a peer's st.shared::cluster), so the
// L1 invalidation an acquire.cluster would add (CCTL.IVALL) buys nothing.
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    if (mbar_try(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(b, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// "I no longer read it" needs no cluster-scope release (nothing the peer must see was written): the default
// release.cta arrive is what costs no MEMBAR.ALL.GPU
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// remote store that completes bytes on the destination CTA's transaction barrier: hand-over without any fence
__device__ __forceinline__ void st_async(uint32_t addr, float2 v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "f"(v.x), "f"(v.y), "r"(mbar) : "memory");
}
// shared-memory accesses by 32-bit shared-space address: the swizzled layouts below form every address with at most one
// XOR of an immediate, which the compiler cannot derive from generic-pointer index arithmetic
__device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 ::"r"(addr), "f"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void sts1(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ void st_cluster(uint32_t addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar,
                                            unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(s32(dst)), "l"(map), "r"(x), "r"(y), "r"(s32(bar)), "l"(policy) : "memory");
}
// one finished Doppler row (1 KB, 16-byte aligned in M) -> its place in the RDS
__device__ __forceinline__ void bulk_store_row(float2* dst, uint32_t src) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src), "n"(C * sizeof(float2)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr unsigned long long EVICT_FIRST = 0x12F0000000000000ull;       // createpolicy fractional evict_first, 1.0

// XFER: how range bins reach their owner: 0 = st.shared::cluster + one release.cluster arrive per warp,
//       1 = st.async completing bytes on the owner's transaction barrier (no fence anywhere).
// PACKED: butterflies with f32x2 adds (FADD2).
// DETECT: the Doppler group also runs the |X|^2 local-maximum / threshold test of dechirp.py:235-263 on the plane it
//       holds and writes hit masks (rs_detect_fused.cuh) -- the RDS is not read again by a detection pass.  The powers of
//       a finished Doppler row replace the row in M; the two edge rows of every CTA also go to the neighbour CTAs' halo
//       rows with st.async (512 B each, completing on the neighbour's halo barrier); warp w then walks its own 8 rows:
//       lane = 4 Doppler columns (left / right neighbours by shuffle), 3-row sliding window of horizontal maxima, the
//       three-zone classification of rs_detect.cu with the threshold folded into the neighbour maximum.
template <bool STORE_TMA, int XFER, bool PACKED, int RG, int CB, bool DETECT = false>
__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(THREADS, 1)
fft2d_ws_kernel(const __grid_constant__ CUtensorMap map_cube, const float2* __restrict__ table, const float2* __restrict__ tw_s_g, const float2* __restrict__ tw_c_g,
                float2* __restrict__ rds, int A, int C_total, int chirp0, int dc_removal, int nplanes, const FusedDetectMasks fd) {
    static_assert(!(DETECT && STORE_TMA), "the fused detection reuses the rows a bulk store would still be reading");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr int NSTAGE = CPC / CB;
    constexpr uint32_t STAGE_BYTES = CB * S * sizeof(float2);
    const int tid = threadIdx.x;
    uint32_t q;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
    const int cid = blockIdx.x / NC, ncl = gridDim.x / NC;

    for (int i = tid; i < S; i += THREADS) {
        const int k1 = i >> 4, t = i & 15;
        sm.tw1s[i] = tw_s_g[(k1 * t) & (S - 1)];
    }
    for (int i = tid; i < C; i += THREADS) {
        const int k1 = i >> 3, t = i & 7;
        sm.tw1c[i] = tw_c_g[(k1 * t) & (C - 1)];
    }
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&sm.full_ld[s], 1);
            mbar_init(&sm.empty_ld[s], R_THREADS / RG / 32);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.full_M[b], XFER == 0 ? NC * (R_THREADS / 32) : 1);
            mbar_init(&sm.empty_M[b], NC * (D_THREADS / 32));
            mbar_init(&sm.halo_full[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (DETECT)
        for (int i = tid; i < 4 * C; i += THREADS) (&sm.halo[0][0][0])[i] = -1.f;      // rows outside the plane never win
    __syncthreads();
    cluster_sync_all();                       // every CTA's barriers exist before the first remote arrive / store

    const int warp = tid >> 5, lane = tid & 31;
    if (warp < R_THREADS / 32) {
        // =============================================================== range group
        // RG subgroups of 8 / RG warps; a subgroup owns the stages s = g (mod RG) of every plane (its own TMA requests,
        // its own named barrier), so with RG = 2 the halves drift apart and fill each other's latencies
        constexpr int RGT = R_THREADS / RG;                   // threads of a subgroup
        constexpr int NI = CB * 16 / RGT;                     // items per thread and pass
        constexpr int HS1 = RGT / 16, HS2 = RGT / CB;         // item strides of the slow index in pass 1 / pass 2
        static_assert(NI >= 1 && NI * RGT == CB * 16 && NSTAGE % RG == 0 && NSTAGE <= MAXSTAGE, "range item geometry");
        const int g = tid / RGT, tg = tid - g * RGT;
        const int t1 = tg & 15, hi = tg >> 4;                 // pass 1: fast-time column, chirp hi (+ HS1 i) of the stage
        const int row2 = tg & (CB - 1), k1b = tg / CB;        // pass 2: lanes along chirps, k1 = k1b (+ HS2 i)
        float2 tabv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tabv[j] = table[t1 + 16 * j];
        // remote bases of row `k1b` of M[0], this thread's chirp column of stage 0, in the four owners
        uint32_t mbase[NC], bbase[NC];
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            mbase[o] = mapa(s32(&sm.M[0][0]), o) + (uint32_t)((k1b * MP + q * CPC + row2) * sizeof(float2));
            bbase[o] = mapa(s32(&sm.full_M[0]), o);
        }
        const int chirp_row0 = chirp0 + (int)q * CPC;
        if (tg == 0 && cid < nplanes) {
#pragma unroll 1
            for (int s = g; s < NSTAGE; s += RG) {
                mbar_expect_tx(&sm.full_ld[s], STAGE_BYTES);
                tma_load_2d(sm.stage + s * CB * S, &map_cube, 0, cid * C_total + chirp_row0 + s * CB, &sm.full_ld[s], EVICT_FIRST);
            }
        }
        int it = 0;
        for (int plane = cid; plane < nplanes; plane += ncl, ++it) {
            const int b = it & 1;
            const uint32_t boff = (uint32_t)(b * ROWS * MP * sizeof(float2));
#pragma unroll 1
            for (int s = g; s < NSTAGE; s += RG) {
                float2* st = sm.stage + s * CB * S;
                mbar_wait(&sm.full_ld[s], it & 1);
#pragma unroll 1
                for (int i = 0; i < NI; ++i) {
                    const int row1 = hi + HS1 * i;
                    const uint32_t rowa = s32(st) + (uint32_t)(row1 * S * sizeof(float2));
                    float2 v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = lds2(rowa + (uint32_t)((t1 + 16 * j) * sizeof(float2)));
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = PACKED ? pow2::cmulp(v[j], tabv[j]) : cmul(v[j], tabv[j]);
                    if (PACKED) pow2::dftp<16>(v); else pow2::dft<16>(v);
                    __syncwarp();                                       // the chirp is read before it is rewritten
                    // Y[k1][t1] -> slot 16 k1 + (t1 ^ rot), rot = chirp (^ 8 for odd k1 when a half-warp of pass 2 covers
                    // two k1): pass-1 writes and pass-2 reads are both conflict free in the unpadded 2 KB chirp rows
                    const uint32_t ye = rowa + (uint32_t)(((t1 ^ row1) & 15) * sizeof(float2));
                    const uint32_t yo = CB == 8 ? rowa + (uint32_t)(((t1 ^ row1 ^ 8) & 15) * sizeof(float2)) : ye;
#pragma unroll
                    for (int k1 = 0; k1 < 16; ++k1) sts2(((k1 & 1) ? yo : ye) + (uint32_t)(16 * k1 * sizeof(float2)), v[k1]);
                }
                named_bar_sync(1 + g, RGT);
#pragma unroll 1
                for (int i = 0; i < NI; ++i) {
                    const int k1r = k1b + HS2 * i;
                    float2 u[16];
                    {
                        const int rot = (row2 ^ (CB == 8 ? 8 * (k1r & 1) : 0)) & 15;
                        const uint32_t yb = s32(st) + (uint32_t)((row2 * S + 16 * k1r + rot) * sizeof(float2));
#pragma unroll
                        for (int n2 = 0; n2 < 16; ++n2) u[n2] = lds2(yb ^ (uint32_t)(n2 * sizeof(float2)));
                    }
                    if (i == NI - 1) {
                        // the stage is consumed: order the generic accesses before the next TMA write, release the slot,
                        // and thread 0 of the subgroup requests the same stage of the next plane once all its warps did
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.empty_ld[s]);
                        if (tg == 0 && plane + ncl < nplanes) {
                            mbar_wait(&sm.empty_ld[s], it & 1);
                            mbar_expect_tx(&sm.full_ld[s], STAGE_BYTES);
                            tma_load_2d(st, &map_cube, 0, (plane + ncl) * C_total + chirp_row0 + s * CB, &sm.full_ld[s], EVICT_FIRST);
                        }
                    }
                    const float2* twr = sm.tw1s + k1r * 16;
#pragma unroll
                    for (int n2 = 1; n2 < 16; ++n2) u[n2] = PACKED ? pow2::cmulp(u[n2], twr[n2]) : cmul(u[n2], twr[n2]);
                    if (PACKED) pow2::dftp<16>(u); else pow2::dft<16>(u);
                    if (s == g && i == 0) mbar_wait(&sm.empty_M[b], ((it >> 1) & 1) ^ 1);   // the owners' Doppler groups left buffer b
                    const uint32_t roff = boff + (uint32_t)((i * HS2 * MP + s * CB) * sizeof(float2));
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        // k = k1r + 16 k2, range fftshift p = (k + S/2) mod S = k1r + 16 ((k2 + 8) mod 16)
                        const int k2s = (k2 + 8) & 15;
                        const int owner = k2s >> 2, pl16 = k2s & 3;                   // p = 64 owner + 16 pl16 + k1r
                        float2 val = u[k2];
                        if (k2 == 0 && dc_removal && k1r == 0) val = make_float2(0.f, 0.f);   // mean removal, dechirp.py:120
                        const uint32_t dst = mbase[owner] + roff + (uint32_t)(pl16 * 16 * MP * sizeof(float2));
                        if (XFER == 0) st_cluster(dst, val);
                        else st_async(dst, val, bbase[owner] + (uint32_t)(b * 8));
                    }
                }
            }
            if (XFER == 0) {
                __syncwarp();
                if (lane < NC) mbar_arrive_cluster(mapa(s32(&sm.full_M[b]), lane));
            }
        }
    } else {
        // =============================================================== Doppler group
        const int w = warp - R_THREADS / 32;
        const int td = lane & 7;
        float2 twd[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) twd[k1] = sm.tw1c[k1 * 8 + td];
        const int k1p = lane & 15;
        // DETECT: where this warp's edge row goes (row 0 -> the upper halo of CTA q - 1, row ROWS - 1 -> the lower halo of
        // CTA q + 1), the range gate of this thread's 8 rows as a mask over its 32 cells
        const bool send_dn = DETECT && w == 0 && q > 0, send_up = DETECT && w == D_THREADS / 32 - 1 && q < NC - 1;
        uint32_t halo_dst = 0, halo_bar = 0, gatew = 0;
        if (DETECT) {
            if (send_dn) { halo_dst = mapa(s32(&sm.halo[0][1][0]), q - 1); halo_bar = mapa(s32(&sm.halo_full[0]), q - 1); }
            if (send_up) { halo_dst = mapa(s32(&sm.halo[0][0][0]), q + 1); halo_bar = mapa(s32(&sm.halo_full[0]), q + 1); }
#pragma unroll
            for (int rl = 0; rl < 8; ++rl) gatew |= fd.gate[q * ROWS + 8 * w + rl] ? (0x01010101u << rl) : 0u;
        }
        const uint32_t halo_bytes = (uint32_t)(((q > 0) + (q < NC - 1)) * C * sizeof(float));
        int it = 0;
        for (int plane = cid; plane < nplanes; plane += ncl, ++it) {
            const int b = it & 1;
            const int f = plane / A, a = plane - f * A;
            float2* Mb = sm.M[b];
            if (XFER == 1 && w == 0 && lane == 0) mbar_expect_tx(&sm.full_M[b], ROWS * C * sizeof(float2));
            if (DETECT && w == 1 && lane == 0) mbar_expect_tx(&sm.halo_full[b], halo_bytes);
            mbar_wait(&sm.full_M[b], (it >> 1) & 1);
            // pass 1, in place: radix 16 over chirps td + 8 j; V[k1][td] -> slot 8 k1 + (td ^ (k1 >> 1))
            const uint32_t Ma = s32(Mb);
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                const uint32_t ma = Ma + (uint32_t)(((8 * w + 4 * i + (lane >> 3)) * MP + td) * sizeof(float2));
                float2 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = lds2(ma + (uint32_t)(8 * j * sizeof(float2)));
                __syncwarp();
                if (PACKED) pow2::dftp<16>(v); else pow2::dft<16>(v);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1)
                    sts2((ma ^ (uint32_t)((k1 >> 1) * sizeof(float2))) + (uint32_t)(8 * k1 * sizeof(float2)),
                         (k1 == 0) ? v[0] : PACKED ? pow2::cmulp(v[k1], twd[k1]) : cmul(v[k1], twd[k1]));
            }
            __syncwarp();
            // pass 2: radix 8, Doppler fftshift
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                // DETECT: the last warp starts with the CTA's last row, so that both edge rows are under way first
                const int ii = (DETECT && w == D_THREADS / 32 - 1) ? 3 - i : i;
                const int row = 8 * w + 2 * ii + (lane >> 4);
                const uint32_t rowa = Ma + (uint32_t)(row * MP * sizeof(float2));
                const uint32_t ub = rowa + (uint32_t)((8 * k1p + (k1p >> 1)) * sizeof(float2));
                float2 u[8];
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) u[n2] = lds2(ub ^ (uint32_t)(n2 * sizeof(float2)));
                if (PACKED) pow2::dftp<8>(u); else pow2::dft<8>(u);
                const int p = q * ROWS + row;
                // Doppler fftshift: bin k1p + 16 k2 goes to (k1p + 16 k2 + 64) mod 128; k1p < 16, so no runtime wrap
                float2* dst = rds + (((size_t)f * S + p) * A + a) * C + k1p;
                if (!STORE_TMA) {
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) __stcs(dst + ((16 * k2 + C / 2) & (C - 1)), u[k2]);
                } else {
                    __syncwarp();                                       // both rows are read before they are rewritten
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2)
                        sts2(rowa + (uint32_t)((k1p + ((16 * k2 + C / 2) & (C - 1))) * sizeof(float2)), u[k2]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (k1p == 0) bulk_store_row(dst, rowa);
                }
                if (DETECT) {
                    // |X|^2 in Doppler order replaces the row (its first 512 B); the CTA's edge rows also go to the neighbours
                    float pw8[8];
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) pw8[k2] = fmaf(u[k2].x, u[k2].x, u[k2].y * u[k2].y);
                    __syncwarp();                                       // both rows are read before they are rewritten
                    if (!(fd.dbg & 4)) {
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2)
                        sts1(rowa + (uint32_t)((k1p + ((16 * k2 + C / 2) & (C - 1))) * sizeof(float)), pw8[k2]);
                    }
                    if ((send_dn && row == 0) || (send_up && row == ROWS - 1)) {
                        const uint32_t hd = halo_dst + (uint32_t)(b * 2 * C * sizeof(float)), hb = halo_bar + (uint32_t)(b * 8);
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2)
                            st_async_f32(hd + (uint32_t)((k1p + ((16 * k2 + C / 2) & (C - 1))) * sizeof(float)), pw8[k2], hb);
                    }
                }
            }
            if (STORE_TMA) {
                if (k1p == 0) asm volatile("cp.async.bulk.commit_group;\n cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (DETECT) {
                __syncwarp();                                           // this warp's 8 power rows are in place
                auto rowaddr = [&](int rl) -> uint32_t {                // rl = -1 .. 8 relative to this warp's first row
                    const int row = 8 * w + rl;
                    if (row < 0) return s32(&sm.halo[b][0][0]);
                    if (row >= ROWS) return s32(&sm.halo[b][1][0]);
                    return Ma + (uint32_t)(row * MP * sizeof(float2));
                };
                // a row of 4 powers per lane with the edge columns of the neighbouring lanes: horizontal maxima
                struct Row { float c[4], L, R, h[4]; };
                auto ld_row = [&](int rl) { return lds4(rowaddr(rl) + (uint32_t)(lane * 16)); };
                auto fin_row = [&](const float4& v) {                   // neighbour columns by shuffle, horizontal maxima
                    Row r;
                    r.c[0] = v.x; r.c[1] = v.y; r.c[2] = v.z; r.c[3] = v.w;
                    r.L = __shfl_up_sync(0xffffffffu, v.w, 1);
                    r.R = __shfl_down_sync(0xffffffffu, v.x, 1);
                    if (lane == 0) r.L = -1.f;                          // Doppler bins -1 and C: outside the plane
                    if (lane == 31) r.R = -1.f;
                    r.h[0] = max3(r.L, v.x, v.y); r.h[1] = max3(v.x, v.y, v.z);
                    r.h[2] = max3(v.y, v.z, v.w); r.h[3] = max3(v.z, v.w, r.R);
                    return r;
                };
                const float band = 2.f * fd.eps;
                uint32_t hitw = 0u, surew = 0u;
                float2 psum2 = make_float2(0.f, 0.f);
                // bit <- (x >= m): one FSETP + one predicated LOP3 (the compiler's SEL + 3-input add form costs 2.5 per test)
                auto set_if_ge = [](uint32_t& w, float x, float m, uint32_t bit) {
                    asm("{\n .reg .pred p;\n setp.ge.f32 p, %1, %2;\n @p or.b32 %0, %0, %3;\n}" : "+r"(w) : "f"(x), "f"(m), "r"(bit));
                };
                auto cells = [&](const Row& ra, const Row& rb, const Row& rc, const int rl) {
                    // the two scaled copies of the centre powers and the power sum in packed f32x2 arithmetic (FFMA2 / FADD2:
                    // IEEE-identical to the scalar fmaf, half the instructions)
                    float2 cu2[2], cl2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const unsigned long long cc = pow2::pk2(rb.c[2 * h], rb.c[2 * h + 1]);
                        unsigned long long u, l, ps = pow2::pk2(psum2.x, psum2.y);
                        asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(u) : "l"(cc), "l"(pow2::pk2(band, band)));
                        asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(l) : "l"(cc), "l"(pow2::pk2(-band, -band)));
                        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ps) : "l"(cc));
                        cu2[h] = pow2::up2(u);
                        cl2[h] = pow2::up2(l);
                        psum2 = pow2::up2(ps);
                    }
                    const float cu[4] = {cu2[0].x, cu2[0].y, cu2[1].x, cu2[1].y}, cl[4] = {cl2[0].x, cl2[0].y, cl2[1].x, cl2[1].y};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float left = j == 0 ? rb.L : rb.c[j - 1], right = j == 3 ? rb.R : rb.c[j + 1];
                        // best neighbour, with "above the threshold" folded in: p > thr <=> p >= thrn
                        const float m2 = max3(ra.h[j], rc.h[j], max3(left, right, fd.thrn));
                        set_if_ge(hitw, cu[j], m2, 1u << (8 * j + rl));      // surely or maybe a detection
                        set_if_ge(surew, cl[j], m2, 1u << (8 * j + rl));     // surely one
                    }
                };
                // The walk is latency, not issue: its LDS.128 -> shuffle -> FMNMX3 chains ran at 0.3 IPC.  All eight rows of the
                // warp are requested at once, so the shared-memory latency is paid once and only the shuffles stay in the chains.
                float4 v[8];
#pragma unroll
                for (int rl = 0; rl < 8; ++rl) v[rl] = ld_row(rl);
                if (!(fd.dbg & 1)) {
                    // rows 1 .. 6 need this warp's rows only: no waiting for anybody
                    Row ra = fin_row(v[0]), rb = fin_row(v[1]);
#pragma unroll
                    for (int rl = 1; rl < 7; ++rl) {
                        const Row rc = fin_row(v[rl + 1]);
                        cells(ra, rb, rc, rl);
                        ra = rb;
                        rb = rc;
                    }
                }
                // rows 0 and 7 also need the last / first row of the neighbouring warp (one 64-thread barrier per pair of
                // warps) or of the neighbouring CTA (the halo rows)
                if (!(fd.dbg & 2)) {
                if (w > 0) named_bar_sync(2 + w, 64);
                if (w < D_THREADS / 32 - 1) named_bar_sync(3 + w, 64);
                if (w == 0 || w == D_THREADS / 32 - 1) mbar_wait(&sm.halo_full[b], (it >> 1) & 1);
                }
                if (!(fd.dbg & 1)) {
                    const float4 vlo = ld_row(-1), vhi = ld_row(8);
                    cells(fin_row(vlo), fin_row(v[0]), fin_row(v[1]), 0);
                    cells(fin_row(v[6]), fin_row(v[7]), fin_row(vhi), 7);
                }
                hitw &= gatew;
                uint32_t uncw = hitw & ~surew, nearw = 0u, candw = 0u;  // ~1e-5 of the cells: inside the 2 eps band
                const bool anyunc = __any_sync(0xffffffffu, uncw != 0u);
                if (anyunc) {
                    while (uncw) {                                      // the exact rule on the nine powers, re-read from M
                        const int bi = __ffs(uncw) - 1;
                        uncw &= uncw - 1;
                        const int rl = bi & 7, d = 4 * lane + (bi >> 3);
                        auto at = [&](int r_, int d_) {
                            return (d_ < 0 || d_ >= C) ? -1.f : lds1(rowaddr(r_) + (uint32_t)(d_ * sizeof(float)));
                        };
                        const float c = at(rl, d);
                        float m = max3(at(rl - 1, d - 1), at(rl - 1, d), at(rl - 1, d + 1));
                        m = fmaxf(m, fmaxf(at(rl, d - 1), at(rl, d + 1)));
                        m = fmaxf(m, max3(at(rl + 1, d - 1), at(rl + 1, d), at(rl + 1, d + 1)));
                        const int cls = rs_classify(c, m, fd.thr, fd.eps);
                        if (cls == 0) hitw &= ~(1u << bi);
                        if (cls & 2) nearw |= 1u << bi;
                        if (cls & 4) candw |= 1u << bi;
                    }
                }
                float psum = psum2.x + psum2.y;
#pragma unroll
                for (int off = 16; off; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
                const size_t gi = (size_t)plane * FD_GROUPS + q * (ROWS / 8) + w;
                fd.hit[gi * FD_WORDS + lane] = hitw;
                if (anyunc) {
                    fd.near[gi * FD_WORDS + lane] = nearw;
                    fd.cand[gi * FD_WORDS + lane] = candw;
                }
                if (lane == 0) fd.rec[gi] = make_float2(psum, __int_as_float(anyunc ? 1 : 0));
            }
            __syncwarp();
            if (lane < NC) mbar_arrive_remote(mapa(s32(&sm.empty_M[b]), lane));
        }
        if (STORE_TMA && k1p == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    cluster_sync_all();                       // no CTA leaves while a peer may still store to it or arrive on it
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int NVARIANT = 5;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

}  // namespace ws

// Returns 1 when the kernel was launched, 0 when this device / shape cannot take it (the caller falls back), < 0 on error.
// fd != nullptr: the default variant with the fused detection (hit masks of every plane into fd's buffers).
int rs_fft2d_ws_launch(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c, void* rds, int F,
                       int A, int C_total, int chirp0, int dc_removal, int store_tma, int variant, const FusedDetectMasks* fd,
                       cudaStream_t stream) {
    using namespace ws;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 0;
    if (((uintptr_t)cube & 15) || ((uintptr_t)rds & 15)) return 0;
    const long long rows_in = (long long)F * A * C_total;
    if (rows_in >= (1ll << 31) || (long long)F * S >= (1ll << 31)) return 0;
    if (sizeof(Smem) > (size_t)rs_smem_optin_limit()) return 0;

    if (variant < 0 || variant >= NVARIANT || fd) variant = 4;
    if (fd) store_tma = 0;
    const int cb = variant == 4 ? 8 : 16;                  // chirps per stage = rows of the TMA box
    CUtensorMap map_cube;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)S, (cuuint64_t)rows_in};
        cuuint64_t gstr[1] = {(cuuint64_t)S * sizeof(float2)};
        cuuint32_t box[2] = {(cuuint32_t)S, (cuuint32_t)cb};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&map_cube, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(cube), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 0;
    }
    typedef void (*Kern)(const CUtensorMap, const float2*, const float2*, const float2*, float2*, int, int, int, int, int,
                         const FusedDetectMasks);
    // the measured variants: [bulk store][variant]: 0 = release-arrive hand-over, scalar, 1 subgroup (the first version),
    // 1 = st.async, scalar, 1 subgroup; 2 = st.async, packed, 1 subgroup; 3 = st.async, packed, 2 subgroups;
    // 4 = as 3 with 8-chirp stages: four ring slots, two per subgroup (a slot is refilled 3/4 of a plane ahead)
    static const Kern kerns[2][NVARIANT] = {
        {fft2d_ws_kernel<false, 0, false, 1, 16>, fft2d_ws_kernel<false, 1, false, 1, 16>, fft2d_ws_kernel<false, 1, true, 1, 16>,
         fft2d_ws_kernel<false, 1, true, 2, 16>, fft2d_ws_kernel<false, 1, true, 2, 8>},
        {fft2d_ws_kernel<true, 0, false, 1, 16>, fft2d_ws_kernel<true, 1, false, 1, 16>, fft2d_ws_kernel<true, 1, true, 1, 16>,
         fft2d_ws_kernel<true, 1, true, 2, 16>, fft2d_ws_kernel<true, 1, true, 2, 8>}};
    Kern kern = fd ? (Kern)fft2d_ws_kernel<false, 1, true, 2, 8, true> : kerns[store_tma ? 1 : 0][variant];
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    static int max_clusters[3][NVARIANT] = {{-1, -1, -1, -1, -1}, {-1, -1, -1, -1, -1}, {-1, -1, -1, -1, -1}};
    int& mc = max_clusters[fd ? 2 : store_tma ? 1 : 0][variant];
    if (mc < 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(NC * 64);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = sizeof(Smem);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 0;
        }
        mc = n;
    }
    int ncl = mc;
    const char* env = getenv("RS_K12_CLUSTERS");           // tuning knob
    if (env && atoi(env) > 0) ncl = atoi(env);
    if (ncl <= 0) return 0;
    const int nplanes = F * A;
    if (ncl > nplanes) ncl = nplanes;
    kern<<<ncl * NC, THREADS, sizeof(Smem), stream>>>(map_cube, (const float2*)table, (const float2*)twiddle_s,
                                                       (const float2*)twiddle_c, (float2*)rds, A, C_total, chirp0, dc_removal,
                                                       nplanes, fd ? *fd : FusedDetectMasks{});
    if (cudaGetLastError() != cudaSuccess) return 0;
    return 1;
}

extern "C" int rs_fft2d_ws_max_clusters(void) {
    using namespace ws;
    auto kern = fft2d_ws_kernel<false, 1, true, 2, 8>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 64);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = sizeof(Smem);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}
