// Detection fused into the Doppler phase of the persistent 2-D FFT kernel (rs_fft2d_ws.cu), SURVEY.md section 8 row a8,
// dechirp.py:235-263: what the FFT kernel leaves behind for the compaction kernel (rs_detect.cu), and the classification
// rule both detection paths share.
//
// The FFT kernel holds a whole (frame, antenna) plane on chip when the Doppler pass finishes, so the |X|^2 local-maximum
// / threshold test runs there and the RDS is not read again: per plane it writes 32 row groups (8 range bins each) of
//   hit   uint32 [32 lanes]   lane l covers Doppler bins 4 l .. 4 l + 3; bit 8 j + rl <=> cell (range 8 g + rl, Doppler 4 l + j)
//                             is a detection or a near-miss candidate (the same set rs_detect's hit masks hold): byte j of
//                             the word is the 8-row column of Doppler bin 4 l + j
//   near / cand               the same layout for RS_FLAG_NEARMAX / near-miss candidates; written only for row groups whose
//                             record says so (~1 % of them)
//   rec   float2              x = sum of |X|^2 over the row group (noise level for the recheck bound), y = int: near / cand
//                             words were written
// 4 KB + 256 B per 256 KB plane instead of the 256 KB read of a separate detection pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FusedDetectMasks {
    uint32_t* hit;          // [planes][32 row groups][32]
    uint32_t* near;
    uint32_t* cand;
    float2* rec;            // [planes][32 row groups]
    const uint8_t* gate;    // [256] range gate
    float thr;              // power threshold (strict)
    float thrn;             // nextafterf(thr, +inf):  p > thr  <=>  p >= thrn
    float eps;              // det_eps
    int dbg;                // timing probes only (RS_FD_DBG): 1 skip the walk, 2 skip barriers / halo wait, 4 skip the power stores
};

// Request handed to the 2-D FFT dispatcher (rs_fft2d.cu) by rs_range_doppler_detect (rs_detect.cu): the mask buffers, and
// what to do with the frames that did NOT go through the fused kernel -- the dispatcher sends the last few frames of a
// batch to a side kernel on the SMs the 4-CTA clusters strand; side_hook runs the stand-alone detection for them on the
// same side stream.  frames_masked: frames [0, frames_masked) have masks when the call returns (0: fused kernel not used).
struct FusedDetectReq {
    FusedDetectMasks masks;
    void (*side_hook)(void* ctx, int f0, int nf, cudaStream_t stream);
    void* ctx;
    int frames_masked;
};

constexpr int FD_GROUPS = 32;       // row groups per 256-row plane
constexpr int FD_WORDS = 32;        // words per row group and mask
constexpr size_t FD_BYTES_PER_PLANE = (size_t)FD_GROUPS * (3 * FD_WORDS * sizeof(uint32_t) + sizeof(float2));

// 0: not a detection; 1: detection; 3: detection whose margin to the best neighbour / threshold is inside
// the fp32 guard band; 7: NOT a detection in fp32 but inside the band (a candidate the fp64 recheck may promote).
__device__ __forceinline__ int rs_classify(float c, float m, float thr, float eps) {
    // cheap reject (91 % of the cells): more than 2 eps below the best neighbour or the threshold
    const float cu = fmaf(c, 2.f * eps, c);
    if (cu < m || cu <= thr) return 0;
    const bool ge_m = c >= m, gt_t = c > thr;
    const bool near_m = fabsf(c - m) <= eps * fmaxf(c, m);
    const bool near_t = fabsf(c - thr) <= eps * fabsf(thr);
    if (ge_m && gt_t) return (near_m || near_t) ? 3 : 1;
    if ((ge_m || near_m) && (gt_t || near_t)) return 7;
    return 0;
}
