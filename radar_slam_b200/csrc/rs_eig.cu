// General-covariance MUSIC: Hermitian Jacobi eigendecomposition in registers + noise-subspace scan.
//
// The reference's music_spectrum (angle_estimation.py:109-154) builds R, calls eigh, sorts the eigenpairs in
// descending order, keeps E_n = V[:, num_sources:] and evaluates 1 / |a^H E_n E_n^H a| with a 1e-12 guard.
// In the reference R is always the rank-1 outer product of one snapshot, for which the batched path uses the
// closed form (rs_angles).  This kernel is the general path for an arbitrary Hermitian R (multi-snapshot or
// spatially smoothed covariances, num_sources > 1): AP lanes per matrix (32 / AP matrices per warp, AP = A padded to a
// power of two), lane j of a group owns column j of R and of V (2 * A complex registers), parallel cyclic Jacobi with the round-robin pair ordering -- the A/2 disjoint
// pairs of a round are rotated simultaneously, column updates exchange the two columns of a pair by shuffle,
// row updates are local to every lane -- then eigenvalue ranking by shuffle and the grid scan with lanes
// summing |a_g^H v_j|^2 over the noise eigenvectors.
#include "rs_common.cuh"

namespace {

constexpr int EIG_WARPS = 4;

template <int AP>   // AP = padded size (power of two >= A, <= 32)
__global__ void __launch_bounds__(EIG_WARPS * 32)
music_cov_kernel(const float2* __restrict__ cov, int n, int A, int num_sources, const float2* __restrict__ steer, int G,
                 int sweeps, float* __restrict__ eigvals, float2* __restrict__ eigvecs, float* __restrict__ spectrum,
                 int32_t* __restrict__ aidx) {
    // lane = column index inside the group of AP lanes that shares one matrix; all shuffles stay inside the group
    constexpr int GROUPS = 32 / AP;
    const int lane = (threadIdx.x & 31) % AP, wid = threadIdx.x >> 5;
    const int mat_raw = (blockIdx.x * EIG_WARPS + wid) * GROUPS + (threadIdx.x & 31) / AP;
    const bool mat_ok = mat_raw < n;
    const int mat = mat_ok ? mat_raw : n - 1;        // idle groups recompute the last matrix and write nothing
    const unsigned full = 0xffffffffu;
    // lane j < A holds column j:  r[i] = R[i][j],  v[i] = V[i][j]
    float2 r[AP], v[AP];
#pragma unroll
    for (int i = 0; i < AP; ++i) {
        r[i] = (lane < A && i < A) ? cov[((size_t)mat * A + i) * A + lane] : make_float2(0.f, 0.f);
        v[i] = make_float2(i == lane ? 1.f : 0.f, 0.f);
    }
    // padded rows/columns: identity block with a huge negative diagonal would disturb nothing because all their
    // off-diagonal entries are exactly zero (rotation angle 0); keep them at 0.

    // round-robin tournament on AP players: in round t, player AP-1 is fixed, the others rotate
    for (int sw = 0; sw < sweeps; ++sw) {
        for (int round = 0; round < AP - 1; ++round) {
            // partner of this lane in the current round
            int me = lane < AP ? lane : 0;
            int partner;
            if (me == AP - 1) partner = round;
            else if (me == round) partner = AP - 1;
            else partner = ((2 * round - me) % (AP - 1) + (AP - 1)) % (AP - 1);
            const int p = min(me, partner), q = max(me, partner);
            // rotation parameters from the 2x2 block [[R_pp, R_pq],[conj(R_pq), R_qq]]; lane p owns R_pp = r[p] and
            // R_qp = r[q] = conj(R_pq); lane q owns R_qq = r[q]
            float mine_diag = 0.f;
            float2 off = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < AP; ++i) {
                if (i == me) mine_diag = r[i].x;
                if (i == q && me == p) off = make_float2(r[i].x, -r[i].y);       // R_pq = conj(R_qp)
            }
            const float other_diag = __shfl_sync(full, mine_diag, partner, AP);
            float2 bpq;                                                          // R_pq, known to both lanes
            bpq.x = __shfl_sync(full, off.x, p, AP);
            bpq.y = __shfl_sync(full, off.y, p, AP);
            const float app = (me == p) ? mine_diag : other_diag;
            const float aqq = (me == p) ? other_diag : mine_diag;
            const float babs = sqrtf(bpq.x * bpq.x + bpq.y * bpq.y);
            float c = 1.f, s = 0.f;
            float2 ph = make_float2(1.f, 0.f);                                   // e^{i phi}, phi = arg R_pq
            if (babs > 1e-30f && me != partner && p < A && q < A) {
                const float tau = (aqq - app) / (2.f * babs);
                const float t = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
                c = rsqrtf(1.f + t * t);
                s = t * c;
                ph = make_float2(bpq.x / babs, bpq.y / babs);
            }
            // U restricted to (p,q): [[c, s e^{i phi}], [-s e^{-i phi}, c]];  R <- U^H R U,  V <- V U
            // ---- column update: col_p' = c col_p - s e^{-i phi} col_q ; col_q' = s e^{i phi} col_p + c col_q
            {
                const float2 w = (me == p) ? make_float2(-s * ph.x, s * ph.y)     // -s e^{-i phi}
                                           : make_float2(s * ph.x, s * ph.y);      //  s e^{+i phi}
#pragma unroll
                for (int i = 0; i < AP; ++i) {
                    const float2 orr = make_float2(__shfl_sync(full, r[i].x, partner, AP), __shfl_sync(full, r[i].y, partner, AP));
                    const float2 ovv = make_float2(__shfl_sync(full, v[i].x, partner, AP), __shfl_sync(full, v[i].y, partner, AP));
                    r[i] = make_float2(c * r[i].x + (w.x * orr.x - w.y * orr.y), c * r[i].y + (w.x * orr.y + w.y * orr.x));
                    v[i] = make_float2(c * v[i].x + (w.x * ovv.x - w.y * ovv.y), c * v[i].y + (w.x * ovv.y + w.y * ovv.x));
                }
            }
            // ---- row update, every lane, for every pair (pp, qq) of the round:
            //      row_pp' = c row_pp - s e^{+i phi} row_qq ; row_qq' = s e^{-i phi} row_pp + c row_qq
#pragma unroll
            for (int pp = 0; pp < AP; ++pp) {
                // lane pp broadcasts its (partner, c, s, ph) ; handle each pair once (pp < its partner)
                const int qq = __shfl_sync(full, partner, pp, AP);
                const float cc = __shfl_sync(full, c, pp, AP), ss = __shfl_sync(full, s, pp, AP);
                const float2 pph = make_float2(__shfl_sync(full, ph.x, pp, AP), __shfl_sync(full, ph.y, pp, AP));
                if (pp < qq) {
                    float2 rp = make_float2(0.f, 0.f), rq = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < AP; ++i) {
                        if (i == pp) rp = r[i];
                        if (i == qq) rq = r[i];
                    }
                    const float2 a1 = make_float2(ss * (pph.x * rq.x - pph.y * rq.y), ss * (pph.x * rq.y + pph.y * rq.x));   // s e^{i phi} rq
                    const float2 a2 = make_float2(ss * (pph.x * rp.x + pph.y * rp.y), ss * (pph.x * rp.y - pph.y * rp.x));   // s e^{-i phi} rp
                    const float2 np_ = make_float2(cc * rp.x - a1.x, cc * rp.y - a1.y);
                    const float2 nq_ = make_float2(a2.x + cc * rq.x, a2.y + cc * rq.y);
#pragma unroll
                    for (int i = 0; i < AP; ++i) {
                        if (i == pp) r[i] = np_;
                        if (i == qq) r[i] = nq_;
                    }
                }
            }
        }
    }
    // eigenvalue of lane j = R[j][j]; rank in descending order (ties by index) via shuffles
    float lam = 0.f;
#pragma unroll
    for (int i = 0; i < AP; ++i)
        if (i == lane) lam = r[i].x;
    int rank = 0;
    for (int j = 0; j < A; ++j) {
        const float lj = __shfl_sync(full, lam, j, AP);
        if (lane < A && (lj > lam || (lj == lam && j < lane))) ++rank;
    }
    if (lane < A && mat_ok) {
        if (eigvals) eigvals[(size_t)mat * A + rank] = lam;
        if (eigvecs) {
#pragma unroll
            for (int i = 0; i < AP; ++i)
                if (i < A) eigvecs[((size_t)mat * A + i) * A + rank] = v[i];     // column `rank` of V, descending order
        }
    }
    // MUSIC scan: den_g = sum over noise eigenvectors |a_g^H v_j|^2   (angle_estimation.py:143-152)
    if (steer == nullptr) return;
    const bool noise = lane < A && rank >= num_sources;
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < G; ++g) {
        float pr = 0.f, pi = 0.f;
        if (noise) {
#pragma unroll
            for (int i = 0; i < AP; ++i) {
                if (i < A) {
                    const float2 w = __ldg(steer + (size_t)i * G + g);      // conj(a_i) v_i
                    pr += w.x * v[i].x + w.y * v[i].y;
                    pi += w.x * v[i].y - w.y * v[i].x;
                }
            }
        }
        float den = pr * pr + pi * pi;
#pragma unroll
        for (int off = AP / 2; off; off >>= 1) den += __shfl_xor_sync(full, den, off, AP);
        const float val = den > 1e-12f ? 1.f / den : 0.f;
        if (lane == 0 && spectrum && mat_ok) spectrum[(size_t)mat * G + g] = val;
        if (val > best) { best = val; bi = g; }
    }
    if (lane == 0 && aidx && mat_ok) aidx[mat] = bi;
}

}  // namespace

extern "C" int rs_music_covariance(const void* cov64, int n, int A, int num_sources, const void* steer64, int G, int sweeps,
                                   float* eigvals, void* eigvecs64, float* spectrum, int32_t* aidx, void* stream) {
    RS_CHECK_ARG(cov64 && n >= 0 && A >= 2 && A <= 32, "rs_music_covariance: need 2 <= A <= 32");
    RS_CHECK_ARG(num_sources >= 0 && num_sources < A && sweeps > 0 && sweeps <= 64, "rs_music_covariance: bad num_sources / sweeps");
    RS_CHECK_ARG(steer64 == nullptr || G > 0, "rs_music_covariance: bad grid");
    if (n == 0) return RS_OK;
    const int ap = A <= 2 ? 2 : A <= 4 ? 4 : A <= 8 ? 8 : A <= 16 ? 16 : 32;
    const int per_block = EIG_WARPS * (32 / ap);
    const int blocks = (n + per_block - 1) / per_block;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_EIG(AP)                                                                                                  \
    music_cov_kernel<AP><<<blocks, EIG_WARPS * 32, 0, st>>>((const float2*)cov64, n, A, num_sources, (const float2*)steer64, \
                                                           G, sweeps, eigvals, (float2*)eigvecs64, spectrum, aidx)
    if (A <= 2) LAUNCH_EIG(2);
    else if (A <= 4) LAUNCH_EIG(4);
    else if (A <= 8) LAUNCH_EIG(8);
    else if (A <= 16) LAUNCH_EIG(16);
    else LAUNCH_EIG(32);
#undef LAUNCH_EIG
    RS_CHECK_LAUNCH("rs_music_covariance");
    return RS_OK;
}
