// Fused dechirp + window + range FFT + Doppler FFT of one (frame, antenna) plane in ONE kernel, the plane kept
// on chip in the shared memory of a thread-block cluster (dechirp.py:143-213).
//
// A 256 x 128 complex64 plane is 256 KB -- more than one SM's shared memory, so the two-kernel path
// (rs_range_fft -> mid -> rs_doppler_fft) writes and re-reads it through HBM: 32 B per cell for a transform whose
// input and output are 16 B per cell.  Here NC CTAs of a cluster share the plane through distributed shared memory:
//   range phase    CTA q reads chirps [q C/NC, (q+1) C/NC) from HBM (coalesced, streamed once), multiplies by
//                  conj(ref)*window, runs the two register passes of the range FFT (rs_fft_pow2.cuh) and stores
//                  range bin p of chirp c into M[p mod S/NC][c] of the CTA that OWNS range bins
//                  [r S/NC, (r+1) S/NC) -- a remote (DSMEM) store for the bins of the peer, lanes along chirps.
//   cluster.sync   every owner now holds all C chirps of its S/NC range bins.
//   Doppler phase  in place in M: pass 1 (radix CR1 in registers, inter-pass twiddle) writes each row back through
//                  a rotation that makes the pass-2 reads conflict free; pass 2 (radix CR2) stores
//                  rds[f][p][a][.] with the Doppler fftshift -- 128-byte runs.
// HBM traffic: 8 B in + 8 B out per cell, the algorithmic minimum of the 2-D transform.
#include <cooperative_groups.h>
#include <cstdlib>
#include "rs_common.cuh"
#include "rs_detect_fused.cuh"
#include "rs_fft_pow2.cuh"

namespace cg = cooperative_groups;

namespace {

template <int SR1, int SR2, int CR1, int CR2, int NC, int F2_THREADS>
struct Fused {
    static constexpr int S = SR1 * SR2, C = CR1 * CR2;
    static constexpr int CB = F2_THREADS / SR2;            // chirps per range batch: one pass-1 item per thread
    static constexpr int RP = pow2::Geo<SR1, SR2>::ROWP_RANGE;
    static constexpr int K1P = pow2::Geo<SR1, SR2>::K1P;
    static constexpr int ROWS = S / NC;                    // range bins owned by one CTA
    static constexpr int CPC = C / NC;                     // chirps one CTA transforms in the range phase
    static constexpr int MP = C + 8;                       // pitch of a row of M (complex): 8 (mod 16)
    static constexpr size_t SMEM = (size_t)(2 * S + C + CB * RP + ROWS * MP) * sizeof(float2);
    static_assert(CPC % CB == 0, "range batches must tile the CTA's chirps");
    static_assert(CB * SR1 == F2_THREADS, "one range pass-2 item per thread");
    static_assert(CR2 == 8 && CR1 == 16, "the in-place rotation below is laid out for a 16 x 8 Doppler split");
    static_assert((ROWS * CR2) % F2_THREADS == 0 && (ROWS * CR1) % F2_THREADS == 0 && F2_THREADS % CR2 == 0,
                  "uniform Doppler loops");
};

// PACKED: the f32x2 arithmetic of rs_fft_pow2.cuh -- the same operations in the same order as the warp-specialised kernel
// (rs_fft2d_ws.cu), so a batch split between the two kernels is bit-identical to either of them alone
template <int SR1, int SR2, int CR1, int CR2, int NC, int F2_THREADS, int MINB, bool PACKED = false>
__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(F2_THREADS, MINB)
fft2d_cluster_kernel(const float2* __restrict__ cube, const float2* __restrict__ table,
                     const float2* __restrict__ tw_s_g, const float2* __restrict__ tw_c_g, float2* __restrict__ rds,
                     int A, int C_total, int chirp0, int dc_removal) {
    using P = Fused<SR1, SR2, CR1, CR2, NC, F2_THREADS>;
    constexpr int S = P::S, C = P::C;
    extern __shared__ float2 sm[];
    float2* tabs = sm;                       // [S] conj(ref) * window
    float2* tw1s = tabs + S;                 // [SR1][SR2] range inter-pass twiddles
    float2* tw1c = tw1s + S;                 // [CR1][CR2] Doppler inter-pass twiddles
    float2* Y = tw1c + C;                    // [CB][RP]   range pass-1 -> pass-2 exchange
    float2* M = Y + P::CB * P::RP;           // [ROWS][MP] this CTA's range bins x all chirps
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int plane = blockIdx.x / NC;       // f * A + a
    const int f = plane / A, a = plane - f * A;
    const float2* src = cube + ((size_t)plane * C_total + chirp0 + q * P::CPC) * S;
    const int row1 = tid / SR2, t1 = tid - row1 * SR2;
    const int row2 = tid % P::CB, k1r = tid / P::CB;

    // The raw chirps are software pipelined: batch b + 1 is requested right after pass 1 of batch b has consumed its
    // registers (batch 0 before anything else), so its HBM latency runs under pass 2 / the table staging.
    float2 v[SR1];
    {
        const float2* x = src + (size_t)row1 * S + t1;
#pragma unroll
        for (int j = 0; j < SR1; ++j) v[j] = __ldcs(x + SR2 * j);                  // streamed once: evict-first
    }
    // "Every CTA of the cluster is resident" is all the first cluster barrier has to establish (no data is handed
    // over), so it is split: a relaxed arrive here, the wait right before the first remote store.
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    for (int i = tid; i < S; i += F2_THREADS) {
        tabs[i] = table[i];
        const int k1 = i / SR2, t = i - k1 * SR2;
        tw1s[i] = tw_s_g[(k1 * t) % S];
    }
    for (int i = tid; i < C; i += F2_THREADS) {
        const int k1 = i / CR2, t = i - k1 * CR2;
        tw1c[i] = tw_c_g[(k1 * t) % C];
    }
    __syncthreads();

    // ---------------- range phase
    // per-thread constants: pass 1 always handles fast-time column t1, so the dechirp table column lives in registers;
    // the inter-pass twiddles w_S^{k1r n2} of pass 2 are read from shared memory (two distinct addresses per warp:
    // broadcasts), which leaves the registers for the chirps in flight
    float2 tabv[SR1];
#pragma unroll
    for (int j = 0; j < SR1; ++j) tabv[j] = tabs[t1 + SR2 * j];
    const float2* twr = tw1s + k1r * SR2;
    for (int b0 = 0; b0 < P::CPC; b0 += P::CB) {
        {
#pragma unroll
            for (int j = 0; j < SR1; ++j) v[j] = PACKED ? pow2::cmulp(v[j], tabv[j]) : cmul(v[j], tabv[j]);
            if (PACKED) pow2::dftp<SR1>(v); else pow2::dft<SR1>(v);
            float2* y = Y + row1 * P::RP + t1;
#pragma unroll
            for (int k1 = 0; k1 < SR1; ++k1) y[k1 * P::K1P] = v[k1];
        }
        if (b0 + P::CB < P::CPC) {
            const float2* x = src + (size_t)(b0 + P::CB + row1) * S + t1;
#pragma unroll
            for (int j = 0; j < SR1; ++j) v[j] = __ldcs(x + SR2 * j);
        }
        __syncthreads();
        {
            const float2* y = Y + row2 * P::RP + k1r * P::K1P;                      // lanes along chirps
            float2 u[SR2];
#pragma unroll
            for (int n2 = 0; n2 < SR2; ++n2) u[n2] = (n2 == 0) ? y[0] : PACKED ? pow2::cmulp(y[n2], twr[n2]) : cmul(y[n2], twr[n2]);
            if (PACKED) pow2::dftp<SR2>(u); else pow2::dft<SR2>(u);
            if (b0 == 0) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // the peers' M exists
            const int chirp = q * P::CPC + b0 + row2;
#pragma unroll
            for (int k2 = 0; k2 < SR2; ++k2) {
                const int k = k1r + SR1 * k2;
                const float2 val = (dc_removal && k == 0) ? make_float2(0.f, 0.f) : u[k2];    // mean removal, dechirp.py:120
                const int p = (k + S / 2) & (S - 1);                                           // range fftshift
                const int owner = p / P::ROWS, pl = p - owner * P::ROWS;
                float2* dstM = cluster.map_shared_rank(M, owner);
                dstM[pl * P::MP + chirp] = val;
            }
        }
        __syncthreads();
    }
    // all chirps of this CTA's range bins have arrived

    cluster.sync();                          // all chirps of this CTA's range bins have arrived (release / acquire)

    // ---------------- Doppler phase, in place in M
    {
        const int td = tid % CR2;                                                   // this thread's column in every item
        float2 twd[CR1];
#pragma unroll
        for (int k1 = 0; k1 < CR1; ++k1) twd[k1] = tw1c[k1 * CR2 + td];
        for (int it = tid; it < P::ROWS * CR2; it += F2_THREADS) {
            const int row = it / CR2;                                               // the CR2 lanes of a row share a warp
            float2* m = M + row * P::MP;
            float2 v[CR1];
#pragma unroll
            for (int j = 0; j < CR1; ++j) v[j] = m[td + CR2 * j];
            __syncwarp();                                                           // the row is read before it is rewritten
            if (PACKED) pow2::dftp<CR1>(v); else pow2::dft<CR1>(v);
#pragma unroll
            for (int k1 = 0; k1 < CR1; ++k1)
                m[k1 * CR2 + ((td + (k1 >> 1)) & (CR2 - 1))] = (k1 == 0) ? v[0] : PACKED ? pow2::cmulp(v[k1], twd[k1]) : cmul(v[k1], twd[k1]);
        }
    }
    __syncthreads();
    for (int it = tid; it < P::ROWS * CR1; it += F2_THREADS) {
        const int row = it / CR1, k1 = it - row * CR1;
        const float2* m = M + row * P::MP + k1 * CR2;
        float2 u[CR2];
#pragma unroll
        for (int n2 = 0; n2 < CR2; ++n2) u[n2] = m[(n2 + (k1 >> 1)) & (CR2 - 1)];
        if (PACKED) pow2::dftp<CR2>(u); else pow2::dft<CR2>(u);
        const int p = q * P::ROWS + row;
        float2* dst = rds + (((size_t)f * S + p) * A + a) * C;
#pragma unroll
        for (int k2 = 0; k2 < CR2; ++k2) __stcs(dst + ((k1 + CR1 * k2 + C / 2) & (C - 1)), u[k2]);   // Doppler fftshift
    }
}

}  // namespace

extern "C" int rs_range_fft(const void* cube, const void* table, const void* twiddle_s, void* mid, int F, int A,
                            int C_total, int chirp0, int C_used, int S, int dc_removal, void* stream);
extern "C" int rs_doppler_fft(const void* mid, const void* twiddle_c, void* rds, int F, int A, int C, int S,
                              void* stream);

// side stream + fork / join events of the calling host thread and current device (created on first use, never freed)
struct ForkJoin {
    cudaStream_t side;
    cudaEvent_t fork, join;
};
static ForkJoin* fork_join() {
    static thread_local ForkJoin fj[16];
    static thread_local bool have[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!have[dev]) {
        if (cudaStreamCreateWithFlags(&fj[dev].side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&fj[dev].fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&fj[dev].join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        have[dev] = true;
    }
    return &fj[dev];
}

// K12 v2 (rs_fft2d_ws.cu): persistent warp-specialised cluster kernel fed by TMA.  1 = launched, 0 = not applicable here.
int rs_fft2d_ws_launch(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c, void* rds, int F,
                       int A, int C_total, int chirp0, int dc_removal, int store_tma, int variant, const FusedDetectMasks* fd,
                       cudaStream_t stream);

// req != nullptr: rs_range_doppler_detect asks for the detection fused into the persistent kernel (rs_detect_fused.cuh)
int rs_range_doppler_fft_impl(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                              void* mid_ws, void* rds, int F, int A, int C_total, int chirp0, int C_used, int S,
                              int dc_removal, void* stream, FusedDetectReq* req) {
    if (req) req->frames_masked = 0;
    RS_CHECK_ARG(cube && table && twiddle_s && twiddle_c && rds, "rs_range_doppler_fft: null pointer");
    RS_CHECK_ARG(F > 0 && A > 0 && A <= RS_MAX_ANTENNAS && S > 0 && C_used > 0 && chirp0 >= 0 && chirp0 + C_used <= C_total,
                 "rs_range_doppler_fft: bad dims");
    const char* env = getenv("RS_FUSED_FFT");          // 0: force the two-kernel path
    if (S == 256 && C_used == 128 && !(env && atoi(env) == 0)) {
        // RS_K12 = ws (default): the warp-specialised TMA kernel; v1: round 1's phase-serial cluster kernel below.
        // RS_K12_STORE = tma: Doppler rows leave through bulk (TMA) stores instead of st.global from registers.
        const char* k12 = getenv("RS_K12");
        if (!(k12 && k12[0] == 'v')) {
            const char* st = getenv("RS_K12_STORE");
            // RS_K12_VARIANT: 0 = st.shared::cluster + release arrives, scalar butterflies; 1 = st.async hand-over;
            // 2 = + f32x2 arithmetic; 3 = + two independent range subgroups; 4 (default) = + 8-chirp stages, 4 ring slots
            const char* vr = getenv("RS_K12_VARIANT");
            // The 4-CTA clusters of the persistent kernel fill 132 of the 148 SMs (a GPC of 18 SMs holds four clusters and
            // strands two SMs).  Pairs of stranded SMs can still host 2-CTA clusters, so the last RS_K12_SIDE permille
            // (default 60) of the frames go to round 1's kernel at NC = 2 on a forked stream, joined before returning:
            // 0.98 -> 0.93 ms per 1000 frames (profiles/k12_side_probe.py; a 4-CTA-cluster side kernel gains nothing).
            // With the fused detection (req) the side frames also run the stand-alone detection kernel on the side stream --
            // but the default there is NO side kernel: that call is made by FramePipeline.process, which runs the fp64
            // recheck of the previous launch set on a second stream at the same time, and the latency-bound recheck kernels
            // make better use of the 16 stranded SMs than 6 % of the FFT does (3.15 against 3.27 ms per 1000-frame step,
            // profiles/side_share_probe.sh).
            const char* sd = getenv("RS_K12_SIDE");
            int side = sd ? atoi(sd) : (req ? 0 : 60);
            int F_side = (F >= 64 && side > 0) ? (int)(((long long)F * side + 500) / 1000) : 0;
            if (F_side >= F) F_side = 0;
            const int F_main = F - F_side;
            cudaStream_t main_st = (cudaStream_t)stream;
            ForkJoin* fj = F_side ? fork_join() : nullptr;
            if (F_side && !fj) { F_side = 0; }
            if (F_side) {                                   // fork first: the side kernel must not wait for the main one
                cudaEventRecord(fj->fork, main_st);
                cudaStreamWaitEvent(fj->side, fj->fork, 0);
            }
            const int r = rs_fft2d_ws_launch(cube, table, twiddle_s, twiddle_c, rds, F_side ? F_main : F, A, C_total, chirp0,
                                             dc_removal, st && st[0] == 't', vr ? atoi(vr) : -1, req ? &req->masks : nullptr,
                                             main_st);
            if (r == 1 && req) req->frames_masked = F_side ? F_main : F;
            if (r == 1 && F_side) {
                using P = Fused<16, 16, 16, 8, 2, 512>;
                auto kern = fft2d_cluster_kernel<16, 16, 16, 8, 2, 512, 1, true>;       // packed: bit-identical to the main kernel
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM);
                const float2* cube_s = (const float2*)cube + (size_t)F_main * A * C_total * S;
                float2* rds_s = (float2*)rds + (size_t)F_main * S * A * C_used;
                kern<<<(unsigned)(F_side * A * 2), 512, P::SMEM, fj->side>>>(cube_s, (const float2*)table, (const float2*)twiddle_s,
                                                                              (const float2*)twiddle_c, rds_s, A, C_total, chirp0,
                                                                              dc_removal);
                const cudaError_t e = cudaGetLastError();
                if (req && req->side_hook) req->side_hook(req->ctx, F_main, F_side, fj->side);
                cudaEventRecord(fj->join, fj->side);
                cudaStreamWaitEvent(main_st, fj->join, 0);
                if (e != cudaSuccess) {
                    rs_set_error("rs_range_doppler_fft: side kernel: %s", cudaGetErrorString(e));
                    return RS_ECUDA;
                }
                return RS_OK;
            }
            if (r == 1) return RS_OK;
            const char* strict = getenv("RS_K12_STRICT");        // tests: no silent fall-back to the round-1 kernel
            if (strict && atoi(strict) == 1) {
                rs_set_error("rs_range_doppler_fft: the warp-specialised kernel could not be launched (%s)",
                             cudaGetErrorString(cudaGetLastError()));
                return RS_ECUDA;
            }
        }
        const char* nc_env = getenv("RS_FUSED_NC");    // tuning knob: CTAs per cluster (2, 4 or 8)
        const int nc = nc_env ? atoi(nc_env) : 4;       // measured on B200 (1k frames 256x128x8): 2: 1.63 ms, 4: 1.53, 8: 1.54
        int launched = 0;
#define LAUNCH_F2D(NC, THREADS, MINB)                                                                                   \
    do {                                                                                                                \
        using P = Fused<16, 16, 16, 8, NC, THREADS>;                                                                    \
        const long long ctas = (long long)F * A * NC;                                                                   \
        if (P::SMEM <= (size_t)rs_smem_optin_limit() && ctas < (1ll << 31)) {                                           \
            auto kern = fft2d_cluster_kernel<16, 16, 16, 8, NC, THREADS, MINB>;                                         \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM);                      \
            kern<<<(unsigned)ctas, THREADS, P::SMEM, (cudaStream_t)stream>>>(                                           \
                (const float2*)cube, (const float2*)table, (const float2*)twiddle_s, (const float2*)twiddle_c,          \
                (float2*)rds, A, C_total, chirp0, dc_removal);                                                          \
            /* a device that cannot co-schedule the cluster rejects the launch: fall back to the two-kernel path */   \
            if (cudaGetLastError() == cudaSuccess) launched = 1;                                                        \
        }                                                                                                               \
    } while (0)
        if (nc == 8) LAUNCH_F2D(8, 256, 3);
        else if (nc == 4) LAUNCH_F2D(4, 256, 2);
        else LAUNCH_F2D(2, 512, 1);
#undef LAUNCH_F2D
        if (launched) return RS_OK;
    }
    RS_CHECK_ARG(mid_ws != nullptr, "rs_range_doppler_fft: this shape needs the mid workspace");
    int rc = rs_range_fft(cube, table, twiddle_s, mid_ws, F, A, C_total, chirp0, C_used, S, dc_removal, stream);
    if (rc != RS_OK) return rc;
    return rs_doppler_fft(mid_ws, twiddle_c, rds, F, A, C_used, S, stream);
}

extern "C" int rs_range_doppler_fft(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                                    void* mid_ws, void* rds, int F, int A, int C_total, int chirp0, int C_used, int S,
                                    int dc_removal, void* stream) {
    return rs_range_doppler_fft_impl(cube, table, twiddle_s, twiddle_c, mid_ws, rds, F, A, C_total, chirp0, C_used, S,
                                     dc_removal, stream, nullptr);
}
