/*
 * radar_slam_b200.h -- C ABI of libradarslam_b200.so (sm_100a).
 *
 * Drop-in boundary for radar-slam's per-frame signal-processing hot path.  The reference has
 * no FFI: its boundary is four Python classes.  The Python modules under
 * src/ keep those classes and call the entry points below through ctypes with raw device
 * pointers (torch is only the allocator / stream owner).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host
 *   - complex64 = interleaved (re, im) float; complex128 = interleaved double
 *   - all entry points enqueue on `stream` (a cudaStream_t passed as void*) and return
 *     immediately; they never allocate, never synchronise and never throw
 *   - return value: RS_OK or a negative RS_E* code; rs_last_error() gives the message
 *   - re-entrant across streams; outputs are caller-allocated
 *
 * Device data layout (one batch of F independent frames)
 *   cube   complex64 [F][A][C][S]   raw frame, reference layout frame_signals[A, C, S]
 *                                   (dechirp.py:168-181)
 *   mid    complex64 [F][S][A][C]   range spectrum, range axis already fftshift-ed (private)
 *   rds    complex64 [F][S][A][C]   range-Doppler spectrum, both axes fftshift-ed, RANGE-MAJOR PLANES:
 *                                   for every range bin the A antenna rows of C Doppler cells follow each
 *                                   other (the row order of `mid`, so the Doppler FFT keeps rows in place and
 *                                   a 16-range-bin detection tile row is one contiguous run).  The A-channel
 *                                   snapshot of a cell (the "spatial signature", angle_estimation.py:83) is A
 *                                   elements at stride C.  The reference's rds[a, r, d] (dechirp.py:193-213)
 *                                   is rds_dev[r][a][d].
 *   detections: per (frame, tile) segments of `seg_cap` slots, `det_count[F*ntiles]` valid each
 *       det_key    uint32  (antenna << 24) | (range_bin << 12) | doppler_bin; ascending key order
 *                          is the reference's output order antenna -> range -> doppler
 *                          (dechirp.py:246-258)
 *       det_power  float   |X|^2 of the cell on that antenna
 *       det_flags  uint8   RS_FLAG_* bits
 *       det_aidx   int32   argmax index into the azimuth grid (-1 for ESPRIT)
 *       det_adeg   float   azimuth in degrees
 *       det_phase  float   angle(s[1] conj(s[0]))            (velocity_solver.py:136)
 *   vel    double [F][8]    v_x v_y v_z w_x w_y w_z success n_targets
 */
#ifndef RADAR_SLAM_B200_H
#define RADAR_SLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_OK            0
#define RS_EINVAL       -1   /* bad argument / unsupported size */
#define RS_ECUDA        -2   /* CUDA runtime error (message in rs_last_error) */
#define RS_ECAPACITY    -3   /* configuration needs more shared memory than one SM has */

#define RS_METHOD_MUSIC        0   /* angle_estimation.py:109-176 */
#define RS_METHOD_BEAMFORMING  1   /* angle_estimation.py:227-251, robust_angle_estimation.py:237-245 */
#define RS_METHOD_ESPRIT       2   /* angle_estimation.py:178-225 */

#define RS_FLAG_TIE        1   /* top-2 grid values closer than tie_eps: argmax not trustworthy in fp32 */
#define RS_FLAG_NEARMAX    2   /* cell within det_eps of a neighbour or of the threshold */
#define RS_FLAG_GUARD      4   /* MUSIC denominator inside the 1e-12 guard zone (angle_estimation.py:149) */
#define RS_FLAG_FIXED      8   /* angle decision re-evaluated in fp64 by rs_recheck_angles_f64 */
#define RS_FLAG_DROPPED   16   /* not a detection (near-miss candidate, or dropped by the fp64 recheck); consumers skip it */
#define RS_FLAG_DETFIXED  32   /* detection decision re-evaluated in fp64 by rs_recheck_detections_f64 */

#define RS_MAX_RANGE_BINS   4096
#define RS_MAX_DOPPLER_BINS 4096
#define RS_MAX_ANTENNAS     256

int rs_version(void);
const char* rs_last_error(void);

/* Number of detection tiles per frame and the tile shape rs_detect will use for [R, D, A]. */
int rs_detect_tiling(int R, int D, int A, int* tile_r, int* tile_d, int* ntiles);

/* (a1)  replaces SignalPreprocessor.process_chirp + the range half of fft2/fftshift
 *       (dechirp.py:143-166, 196-211).  x * table, FFT over fast time, bin 0 zeroed when
 *       dc_removal (== subtracting the mean, dechirp.py:120), range fftshift, stored transposed.
 *       table     complex64 [S]  conj(reference_chirp) * window, built in fp64 by the host
 *       twiddle_s complex64 [S]  exp(-2 pi i k / S), built in fp64 by the host
 *       chirp0, C_used: chirp_subset (dechirp.py:184-187); `cube` holds C_total chirps, `mid` C_used. */
int rs_range_fft(const void* cube, const void* table, const void* twiddle_s, void* mid,
                 int F, int A, int C_total, int chirp0, int C_used, int S, int dc_removal, void* stream);

/* (a2)  the Doppler half of fft2/fftshift (dechirp.py:208-211): FFT over slow time, Doppler
 *       fftshift, rows kept in place (rds row = mid row).   twiddle_c complex64 [C] */
int rs_doppler_fft(const void* mid, const void* twiddle_c, void* rds, int F, int A, int C, int S, void* stream);

/* (a1+a2) the whole 2-D transform in one call.  For power-of-two shapes whose plane fits the shared memory of a
 *       thread-block cluster (S = 256, C_used = 128) one fused kernel keeps the plane on chip: 8 B in + 8 B out per
 *       cell instead of the 32 B of the two-kernel path; other shapes run rs_range_fft + rs_doppler_fft through
 *       mid_ws (complex64 [F][S][A][C_used]; may be NULL when the fused kernel applies). */
int rs_range_doppler_fft(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                         void* mid_ws, void* rds, int F, int A, int C_total, int chirp0, int C_used, int S,
                         int dc_removal, void* stream);

/*       Clusters of the persistent fused kernel that can be co-resident on the current device (0: the device cannot
 *       run it and rs_range_doppler_fft uses the other paths; < 0: CUDA error).  Diagnostic for bench / tests. */
int rs_fft2d_ws_max_clusters(void);

/* (b)   replaces extract_range_doppler_peaks (dechirp.py:215-278): |X|^2, 3x3 local maximum per
 *       antenna plane (scipy maximum_filter 'reflect' == ignore out-of-range neighbours, ties
 *       count), strict threshold, range gate, per-tile compaction by ballot/prefix sums.
 *       thr_power   p > thr_power  <=>  10 log10(p + 1e-12) > threshold_db  (host computes in fp64)
 *       range_gate  uint8 [R], 1 where min_range <= range_bins_m[i] <= max_range
 *       det_eps     relative guard band for RS_FLAG_NEARMAX
 *       det_overflow int32 [F], set to 1 when a tile had more than seg_cap detections
 *       det_lead   uint32 [F*ntiles*seg_cap]; det_nlead int32 [F*ntiles]: one LEADER per distinct
 *                   range-Doppler cell of a segment, position | (multiplicity << 16).  A cell flagged on k
 *                   antennas occupies k consecutive detection slots that share one snapshot, so the angle
 *                   stage evaluates each leader once.   seg_cap <= 65535.
 *       det_nnear   optional int32 [F*ntiles]: RS_FLAG_NEARMAX entries per segment (recheck skips clean segments)
 *       det_psum    optional float [F*ntiles]: sum of |X|^2 over the tile (frame noise level, used by the
 *                   recheck's error bound) */
int rs_detect(const void* rds, const uint8_t* range_gate, float thr_power, float det_eps,
              uint32_t* det_key, float* det_power, uint8_t* det_flags, uint32_t* det_lead,
              int32_t* det_count, int32_t* det_nlead, int32_t* det_overflow, int32_t* det_nnear, float* det_psum,
              int seg_cap, int F, int R, int D, int A, void* stream);

/* (a1+a2+b) rs_range_doppler_fft + rs_detect in one call (dechirp.py:193-263).  For 256 x 128 planes with A % 8 == 0
 *       the detection rides in the Doppler phase of the persistent 2-D FFT kernel: the (frame, antenna) plane is on chip
 *       when the Doppler pass finishes, so |X|^2, the 3x3 local maximum, the threshold and the range gate are evaluated
 *       there and only hit masks (4 KB per 256 KB plane) go to fused_ws; a small kernel turns the masks of an antenna
 *       octet into exactly the segments / order / leaders / counters rs_detect writes.  The RDS is not read by a
 *       detection pass (8 B per cell less HBM traffic).  Other shapes, or fused_ws == NULL, run the two stages one after
 *       the other with identical results.
 *       fused_ws   workspace of rs_fused_detect_ws_bytes(F, A) bytes (16-byte aligned), contents undefined on return
 *       det_power  may be NULL on the fused path (the kernel keeps no |X|^2 and nothing on the path to the velocity reads
 *                  it): rs_detection_power gathers the values on demand, or rs_angles writes them from the snapshots it
 *                  holds anyway (det_power_out) */
long long rs_fused_detect_ws_bytes(int F, int A);
/*       det_power of every listed entry, gathered from the RDS the lists came from (bit-identical to rs_detect's values) */
int rs_detection_power(const void* rds, const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead,
                       float* det_power, int seg_cap, int nseg_per_frame, int F, int R, int D, int A, void* stream);
int rs_range_doppler_detect(const void* cube, const void* table, const void* twiddle_s, const void* twiddle_c,
                            void* mid_ws, void* rds, void* fused_ws, int F, int A, int C_total, int chirp0, int C_used,
                            int S, int dc_removal, const uint8_t* range_gate, float thr_power, float det_eps,
                            uint32_t* det_key, float* det_power, uint8_t* det_flags, uint32_t* det_lead,
                            int32_t* det_count, int32_t* det_nlead, int32_t* det_overflow, int32_t* det_nnear,
                            float* det_psum, int seg_cap, void* stream);

/* (c)   replaces AngleEstimator.process_targets (angle_estimation.py:253-309) for every detection:
 *       snapshot gather, (rank-1) MUSIC / beamforming scan over the azimuth grid with first-index
 *       argmax, or the ESPRIT closed form; also the inter-antenna phase the solver uses.
 *       scan_table  float [G][scan_stride]  (cos k phi_g, sin k phi_g), k = 1..A-1, fp64-built;
 *                   scan_stride = 2 * (A_pad - 1) rounded up to 4        (A <= 16 path)
 *       steer       complex64 [A][G] exp(+i m phi_g)                      (A > 16 path; may be NULL otherwise)
 *       grid_deg    float [G]
 *       esprit_scale  lambda / (2 pi d)   (angle_estimation.py:218)
 *       grid_symmetric  1 when grid_deg[G-1-g] == -grid_deg[g] exactly (halves the scan work)
 *       ls_partials   optional double [F*nseg_per_frame][8]: per-segment fp64 sums
 *                     (sum c^2, sum s^2, sum cs, sum yc, sum ys, sum y^2, n, 0) of the velocity normal
 *                     equations with c,s = grid_cs[aidx]; needs grid_cs (double [G][2]); grid methods, A <= 16
 *       det_ntie      optional int32 [F*nseg_per_frame]: cells flagged TIE / GUARD per segment
 *       det_tielist   optional int32 [F*nseg_per_frame][RS_TIE_LIST_CAP]: leader indices of the first RS_TIE_LIST_CAP
 *                     flagged cells of each segment, in arbitrary order (lets the fp64 recheck skip the list scan)
 *       mma_table     optional, 32-bit words [mma_tiles][2][A_pad/8][32][2] (passed as const float*): fp16 hi / lo B
 *                     fragments of the (cos, sin) tables for the tensor-core scan, packed for mma.sync.m16n8k16
 *                     (radar_slam_b200/tables.py: scan_mma_table); 8 < 2 A_pad <= 32, symmetric grid;
 *                     mma_tiles = ceil(ceil(G/2)/8).  NULL selects the CUDA-core scan.
 *       cell_ws       optional workspace of 17 * F*R*D bytes (16-byte aligned) for A > 16: every distinct cell of a frame is
 *                     evaluated once (mark / evaluate / scatter) instead of once per detection -- with many antennas
 *                     a cell is flagged on many of them and all share one snapshot.  NULL: one evaluation per leader.
 *                     A == 16 with the tcgen05 scan: RS_ANGLES_WS_BYTES of scratch (16-byte aligned); the two antenna-octet
 *                     segments of a tile are then scanned as a pair and a cell flagged on both octets is evaluated once
 *                     (RS_ANGLES_DEDUP=0 turns it off).  NULL: every segment on its own.
 *       tc_table      optional, bytes [tc_halves][cos, sin][KC][32 x 16 fp16]: the same tables as UMMA B operands (K-major,
 *                     no swizzle) for the tcgen05 / TMEM scan (radar_slam_b200/tables.py: scan_tc_table), tc_halves =
 *                     ceil(ceil(G/2)/32); the default scan for 5..16 antennas on a symmetric grid (RS_ANGLES_TC=0: mma.sync).
 *                     A > 16 (MUSIC / beamforming, D a multiple of 128, cell_ws given): the B operands of the steering GEMM
 *                     [tc_halves][ceil(A/8)][hi, lo][192 x 16 fp16] (tables.py: steer_tc_table), tc_halves = ceil(G/96):
 *                     the grid scan runs as a dense contraction on the tensor cores (RS_MUSIC_TC=0: CUDA-core scan).
 *       det_power_out optional float [F*nseg_per_frame*seg_cap]: |X|^2 of every entry (the det_power of rs_detect), for
 *                     lists that came from rs_range_doppler_detect with det_power = NULL.  The scan holds the snapshot
 *                     of every flagged cell in registers anyway, so the powers cost no memory traffic here.
 *       det_nnear     optional int32 [F*nseg_per_frame], the counters rs_detect wrote: a segment without RS_FLAG_NEARMAX
 *                     entries cannot hold an RS_FLAG_DROPPED one, so the scan skips reading the entries' flags there. */
#define RS_TIE_LIST_CAP 32
#define RS_ANGLES_WS_BYTES (32 << 20)
int rs_angles(const void* rds, const float* scan_table, int scan_stride, const void* steer, const float* grid_deg, int G,
              int method, float tie_eps, double esprit_scale,
              const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead, uint8_t* det_flags,
              int32_t* det_aidx, float* det_adeg, float* det_phase,
              int seg_cap, int nseg_per_frame, int F, int R, int D, int A,
              const double* grid_cs, double* ls_partials, int grid_symmetric, int32_t* det_ntie, int32_t* det_tielist,
              const float* mma_table, int mma_tiles, void* cell_ws, const void* tc_table, int tc_halves,
              float* det_power_out, const int32_t* det_nnear, void* stream);

/* (d')  the velocity solve of rs_velocity_ls from the per-segment sums rs_angles already produced
 *       (no second pass over the detection lists); same output row layout.  det_overflow (int32 [F], may be NULL): frames
 *       whose detection segments overflowed in rs_detect report success = 0 (their row holds the solution of the
 *       truncated list). */
int rs_velocity_from_partials(const double* ls_partials, int nseg_per_frame, int F, double k_phase, double bound,
                              const int32_t* det_overflow, double* vel, void* stream);

/* (d'')  the per-segment sums of rs_velocity_from_partials from the detection lists (for what rs_angles does not fuse:
 *       ESPRIT, A > 16): ls_partials double [F*nseg_per_frame][8], one CTA per segment, deterministic.
 *       grid_cs NULL or det_aidx < 0 => cos/sin of radians(det_adeg) in fp64. */
int rs_velocity_partials(const int32_t* det_aidx, const float* det_adeg, const float* det_phase, const uint8_t* det_flags,
                         const int32_t* det_count, const double* grid_cs, double* ls_partials, int seg_cap,
                         int nseg_per_frame, int F, void* stream);

/* (d)   replaces VelocitySolver.solve_velocity / two_step_optimization (velocity_solver.py:178-355):
 *       fp64 normal equations of  y = k (v_x cos az + v_y sin az),  k = 4 pi dt / lambda, solved under
 *       the reference's box |v_x|,|v_y| <= bound; success = (n >= 3).  v_z and omega are unobservable
 *       in the reference model and are returned as 0.  irls_iters > 0 adds Huber reweighting.
 *       grid_cs  double [G][2] (cos, sin) of radians(grid_deg), fp64-built; NULL => use det_adeg. */
int rs_velocity_ls(const int32_t* det_aidx, const float* det_adeg, const float* det_phase, const uint8_t* det_flags,
                   const int32_t* det_count, const double* grid_cs, double k_phase, double bound,
                   int irls_iters, double huber_delta,
                   double* vel, int seg_cap, int nseg_per_frame, int F, void* stream);

/* fp64 re-evaluation of the decisions the fp32 kernels flag as undecidable, straight from the raw cube
 * (direct fp64 DFT of the cells involved; table128 = conj(ref)*window complex128 [S]).  A segment's items are
 * settled in list order by one owner (CTA or warp): deterministic.  stats int32 [4] is zeroed and filled.
 *
 * rs_recheck_detections_f64: RS_FLAG_NEARMAX entries (detections and the near-miss candidates rs_detect emits
 *   as RS_FLAG_DROPPED): exact 3x3 local-maximum and threshold test (dechirp.py:250-254,
 *   thr_power64 = 10^(threshold_db/10), test p + 1e-12 > thr_power64); sets/clears RS_FLAG_DROPPED, sets
 *   RS_FLAG_DETFIXED.  Run it BEFORE rs_angles (last four pointers NULL), or AFTER it with det_aidx, det_phase,
 *   grid_cs and ls_partials so a detection that changes state is moved into / out of the velocity sums.
 *   stats = {rechecked, dropped, promoted, unresolved}.
 * rs_recheck_angles_f64: RS_FLAG_TIE / RS_FLAG_GUARD cells.  Stage A: fp64 grid scan of the fp32 snapshot; the
 *   snapshot's rounding error is bounded by fft_eps * rms(|X|) per element (rms from det_psum), which bounds how far
 *   P_g - P_h can move; if the winner beats every other grid point by more than its bound the decision is final.
 *   Stage B (otherwise, and near the MUSIC 1e-12 guard): the snapshot is recomputed in fp64 from the cube.
 *   First-index argmax of the method's pseudo-spectrum (angle_estimation.py:143-152, 173); rewrites
 *   det_aidx/det_adeg, sets RS_FLAG_FIXED and corrects ls_partials.
 *   Stage B reads each (frame, antenna) plane of the cube once per 8 undecided cells of the frame.
 *   det_tielist: the list rs_angles wrote (NULL, or more than RS_TIE_LIST_CAP flagged cells: the leaders are scanned).
 *   tw_s128 / tw_c128: complex128 exp(-2 pi i k / S) [S] and exp(-2 pi i k / C) [C], built in fp64 by the host.
 *   Workspace (caller-allocated): work_idx int32 [F*nseg*16], work_cnt int32 [F*nseg], work_snap complex128
 *   [F*nseg*16*A], frame_cnt int32 [F], frame_list int32 [F][RS_RECHECK_FRAME_CAP][4] (16-byte aligned).
 *   One pass settles at most 16 undecided cells per segment and RS_RECHECK_FRAME_CAP per frame in stage B; the rest
 *   is counted in stats[3] and left unflagged-as-fixed for another pass.
 *   stats = {rechecked, index changed, needed the fp64 snapshot, unresolved}. */
#define RS_RECHECK_FRAME_CAP 256
int rs_recheck_detections_f64(const void* cube, const void* table128, const void* tw_s128, const void* tw_c128,
                              int C_total, int chirp0, int dc_removal,
                              double thr_power64, const uint32_t* det_key, uint8_t* det_flags,
                              const int32_t* det_count, const int32_t* det_nnear, int seg_cap, int nseg_per_frame,
                              int F, int A, int C, int S,
                              const int32_t* det_aidx, const float* det_phase, const double* grid_cs, double* ls_partials,
                              int32_t* stats, void* stream);
int rs_recheck_angles_f64(const void* cube, const void* table128, const void* tw_s128, const void* tw_c128,
                          int C_total, int chirp0, int dc_removal,
                          const void* rds, const void* steer128, const float* grid_deg, const double* grid_cs,
                          int G, int method, double fft_eps, const float* det_psum, const int32_t* det_ntie,
                          const int32_t* det_tielist, const uint32_t* det_key, const uint32_t* det_lead, const int32_t* det_nlead,
                          uint8_t* det_flags, int32_t* det_aidx, float* det_adeg, const float* det_phase,
                          double* ls_partials, int seg_cap, int nseg_per_frame, int F, int A, int C, int S,
                          int32_t* work_idx, int32_t* work_cnt, void* work_snap, int32_t* frame_cnt,
                          int32_t* frame_list, int32_t* stats, void* stream);

/* (c')  General-covariance MUSIC (angle_estimation.py:109-154 for an ARBITRARY Hermitian R: multi-snapshot or
 *       smoothed covariances, num_sources >= 1): Hermitian Jacobi eigendecomposition in registers, one warp per
 *       matrix (parallel cyclic Jacobi, columns across lanes, rotations by shuffle), descending eigenvalue order,
 *       noise subspace V[:, num_sources:], pseudo-spectrum 1/|a^H E_n E_n^H a| with the 1e-12 guard, first-index
 *       argmax.  cov64 complex64 [n][A][A] (row-major), 2 <= A <= 32; steer64 complex64 [A][G] or NULL (eigen only);
 *       outputs (each optional): eigvals float [n][A] descending, eigvecs64 complex64 [n][A][A] (columns),
 *       spectrum float [n][G], aidx int32 [n]. */
int rs_music_covariance(const void* cov64, int n, int A, int num_sources, const void* steer64, int G, int sweeps,
                        float* eigvals, void* eigvecs64, float* spectrum, int32_t* aidx, void* stream);

/* helpers for the legacy (list-of-dict) adapters ---------------------------------------------- */

/* rds [F][S][A][C] -> reference layout [F][A][S][C] (complex64): a permutation of whole Doppler rows */
int rs_rds_to_reference_layout(const void* rds, void* out, int F, int A, int C, int S, void* stream);

/* reference layout [F][A][S][C] complex64 -> rds [F][S][A][C] */
int rs_rds_from_reference_layout(const void* rds_ref, void* out, int F, int A, int C, int S, void* stream);

/* unit-energy snapshots (angle_estimation.py:83-88) of n cells: keys as in det_key, frame index per key.
 * out complex128 [n][A]. */
int rs_signatures_f64(const void* rds, const uint32_t* keys, const int32_t* frames, int n,
                      void* out, int F, int R, int D, int A, void* stream);

/* fp64 pseudo-spectra for n unit-energy snapshots: method MUSIC -> 1/(M - |a^H s|^2) with the 1e-12
 * guard, BEAMFORMING -> |a^H s|^2.   steer128 complex128 [A][G];  out double [n][G];  aidx int32 [n]. */
int rs_spectra_f64(const void* sig128, const void* steer128, int method, int n, int A, int G,
                   double* out, int32_t* aidx, void* stream);

/* 10 log10(|X|^2 + 1e-12) of the RDS in the reference layout: out double [F][A][S][C]
 * (power_spectrum_db of extract_range_doppler_peaks, dechirp.py:235-238, 277). */
int rs_power_db_f64(const void* rds, double* out, int F, int A, int C, int S, void* stream);

/* SignalPreprocessor.process_chirp for `rows` chirps in fp64: (x conj(ref)) w - mean (dechirp.py:143-166).
 * in128/out128 complex128 [rows][S], ref128 complex128 [S], window double [S]. */
int rs_process_chirps_f64(const void* in128, const void* ref128, const double* window, int rows, int S,
                          int dc_removal, void* out128, void* stream);

/* ---- fp64 frame path of the legacy class API (one frame per call; csrc/rs_legacy_f64.cu) ----------------------------
 * The reference computes RDS, peak mask and spectra in float64 from whatever arrays the caller holds; these entry points
 * do the same on the device so that the legacy methods are exact also for an RDS that was np.load-ed from a stage file.
 *
 * SignalPreprocessor.generate_range_doppler_spectrum (dechirp.py:168-213) in fp64: (x conj(ref)) w - mean per chirp,
 * FFT over fast time, FFT over slow time, fftshift on both axes.
 *   cube128 complex128 [A][C_total][S], ref128 complex128 [S] (the reference chirp, NOT conjugated), window double [S],
 *   twiddle_*128 complex128 [S] / [C_used] = exp(-2 pi i k / n);  rds128 complex128 [A][S][C_used] (reference layout). */
int rs_range_doppler_f64(const void* cube128, const void* ref128, const double* window, const void* twiddle_s128,
                         const void* twiddle_c128, void* rds128, int A, int C_total, int chirp0, int C_used, int S,
                         int dc_removal, void* stream);

/* extract_range_doppler_peaks (dechirp.py:215-278) in fp64 on a complex128 RDS [A][R][D]: power_db = 10 log10(|X|^2 +
 * 1e-12) (written to power_db [A][R][D]), 3x3 'reflect' local maximum on the dB values (ties count), strict threshold,
 * range gate.  keys (det_key packing) come out in the reference's order antenna -> range -> doppler; *total is the
 * number of detections (may exceed cap: then only the first cap keys were written).  row_count int32 [A*R] and
 * row_offset int64 [A*R] are workspace. */
int rs_detect_f64(const void* rds128, const uint8_t* gate, double threshold_db, double* power_db, int* row_count,
                  long long* row_offset, uint32_t* keys, long long cap, long long* total, int A, int R, int D, void* stream);

/* unit-energy snapshots (angle_estimation.py:83-88) of n cells of a complex128 RDS [A][R][D]: out complex128 [n][A]. */
int rs_signatures_c128(const void* rds128, const int* range_bin, const int* doppler_bin, int n, void* out128, int A, int R,
                       int D, void* stream);

/* AngleEstimator.estimate_angle_esprit for n snapshots in fp64 (angle_estimation.py:178-225): out double [n] deg. */
int rs_esprit_f64(const void* sig128, int n, int A, double esprit_scale, double* out, void* stream);

/* VelocitySolver.two_step_optimization for arbitrary positions/angles (velocity_solver.py:178-307):
 * bounded-variable least squares of  k [d_i, r_i x d_i].[v, w] = y_i  over the box lo..hi (double [6]).
 * nvar = 3 pins w = 0 (step 1), nvar = 6 is the full problem (step 2).  pos double [n][3], ang double [n][2]
 * (azimuth, elevation), y double [n];  out7 = v[3], w[3], cost;  pred double [n] = predicted phases. */
int rs_velocity_ls6(const double* pos, const double* ang, const double* y, int n, double k_phase,
                    const double* lo, const double* hi, int nvar, double* out7, double* pred, void* stream);

/* RobustAngleEstimator.compute_angle_confidence (robust_angle_estimation.py:88-138) for n unit-energy
 * snapshots at their estimated angles: 0.4 |a^H s|/|s| + 0.3 exp(-mean|phase error|) + 0.3 min(1, log10(SNR)/3),
 * clipped to [0, 1].  positions double [A] (metres). */
int rs_robust_confidence_f64(const void* sig128, const double* angle_deg, const double* positions,
                             double lambda_c, int n, int A, double* out, void* stream);

/* (f3)  AdvancedVelocityOptimizer.compute_regularized_cost_function (advanced_velocity_optimization.py:153-223) for nq
 *       candidate motions params double [nq][6]: wrapped residual sum of squares + the five regularisers (speed above
 *       0.8 max_velocity, rotation rate above 0.8 max_angular_velocity, change against previous_motion (double [6], may be
 *       NULL), the large-speed / large-rate product, vertical velocity), fp64.  cost double [nq]. */
int rs_regularized_cost(const double* params, const double* pos, const double* ang, const double* y, int n, int nq,
                        double k_phase, double max_velocity, double max_angular_velocity, double weight,
                        const double* previous_motion, double* cost, void* stream);

/* (f3)  Gauss-Newton polish of nq starting points v_xy double [nq][2] inside their basins of
 *       sum_i wrap(y_i - k (v_x cos az_i + v_y sin az_i))^2 + reg |v - centre|^2, clipped to the box; v_xy is updated in
 *       place, cost double [nq] receives the value of that function at the result. */
int rs_wrapped_gn_polish(const double* cos_az, const double* sin_az, const double* y, int n, double k_phase, double reg,
                         double centre_x, double centre_y, double lo_x, double hi_x, double lo_y, double hi_y, double* v_xy,
                         double* cost, int nq, int iters, void* stream);

/* (f3)  ImprovedVelocitySolver (src/algorithms/velocity_solver_improved.py), inter-frame ego-velocity.
 *   rs_associate_targets: associate_targets_across_frames (:74-129) for `pairs` frame pairs.  cur_xy double
 *       [pairs][nc_max][2], prev_xy double [pairs][np_max][2] = range (cos az, sin az); n_cur / n_prev int32 [pairs].
 *       Current targets are visited in order; each takes the nearest still unused previous target closer than
 *       `threshold` (Euclidean distance in fp64 as scipy cdist, first index on ties).  match_idx int32
 *       [pairs][nc_max] (-1 = none), match_dist double [pairs][nc_max].
 *   rs_wrapped_cost: cost_function (:223-266) for nq candidate motions params double [nq][6] = v, w:
 *       sum_i wrap(y_i - k [v + w x pos_i] . dir(az_i, el_i))^2 + reg_v |v|^2 + reg_w |w|^2, wrap = atan2(sin, cos).
 *       pos double [n][3], ang double [n][2], y double [n];  cost double [nq].
 *   rs_wrapped_lattice_search: that cost on the lattice v = (vx_lo + ix h, vy_lo + iy h), ix < nx, iy < ny, for the
 *       planar model x_i = k (v_x cos az_i + v_y sin az_i) (what :336-355 builds: elevation 0, position = range * dir)
 *       with phases in cycles: ax = k cos az / 2 pi, by = k sin az / 2 pi, y_cycles = y / 2 pi (double [n], n <= 2048).
 *       fp32 evaluation with an fp64 re-base every 32 points.  Output: the best lattice point of every tile of
 *       8 rows x 1024 columns (rs_wrapped_lattice_tiles gives the tile grid): tile_cost float, tile_ix / tile_iy int32
 *       [tiles_y][tiles_x]. */
int rs_associate_targets(const double* cur_xy, const int32_t* n_cur, const double* prev_xy, const int32_t* n_prev,
                         double threshold, int32_t* match_idx, double* match_dist, int pairs, int nc_max, int np_max,
                         void* stream);
int rs_wrapped_cost(const double* params, const double* pos, const double* ang, const double* y, int n, int nq,
                    double k_phase, double reg_v, double reg_w, double* cost, void* stream);
int rs_wrapped_lattice_tiles(long long nx, long long ny, int* tiles_x, int* tiles_y);
int rs_wrapped_lattice_search(const double* ax_cycles, const double* by_cycles, const double* y_cycles, int n,
                              double vx_lo, double vy_lo, double h, long long nx, long long ny, double reg,
                              float* tile_cost, int32_t* tile_ix, int32_t* tile_iy, void* stream);

/* (f2)  FMCWRadarSimulator.synthesize_frame (scripts/simulate_raw.py:147-221) for F frames on the device:
 *       cube[f][a][c][s] = sum over scatterers of amplitude exp(i(doppler + antenna phase)) delayed_chirp conj(ref_chirp)
 *       (the same [A][S] plane for every chirp, :190-209) + sqrt(noise_power) (N(0,1) + i N(0,1))  (:216-219).
 *       scatterers    double [F][n_max][4] = range_m, azimuth_rad, rcs_db, radial_velocity; rows that the reference
 *                     skips (range <= 0, non-finite, :177) are skipped; n_scatterers int32 [F] or NULL (= n_max each)
 *       antenna_pos   double [A] metres;  chirp_rate = bandwidth / chirp_duration;  t = linspace(0, chirp_duration, S)
 *       noise         Philox4x32-10 keyed by seed, counter = (cell pair in frame, first_frame + f): frame k does not
 *                     depend on the batch it is generated in.  Same distribution as the reference, not the same samples.
 *       plane_ws      complex64 [F][A][S] workspace;  cube complex64 [F][A][C][S]. */
int rs_synthesize_frames(const double* scatterers, const int32_t* n_scatterers, int n_max, double fc,
                         double chirp_rate, double chirp_duration, double lambda_c, const double* antenna_pos,
                         double noise_power, unsigned long long seed, long long first_frame, void* plane_ws,
                         void* cube, int F, int A, int C, int S, void* stream);

#ifdef __cplusplus
}
#endif
#endif
