/*
 * scopyon_b200 -- C ABI of the B200-native image-formation hot path.
 *
 * The reference (ecell/scopyon) is pure Python and has no FFI; the seam this
 * library plugs into is `_EPIFMSimulator.output_frame / generate_frames`
 * (/root/reference/src/scopyon/_epifm.py:1017-1049,1121-1225) and
 * `sample_inputs / sample` (sampling.py:121-162, sampling2.py:70-149).
 * Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions cross the boundary.
 *   - every pointer named d_* is DEVICE memory owned by the caller.
 *   - every call takes the CUDA stream to enqueue on (a cudaStream_t passed as
 *     void*); calls are asynchronous with respect to the host.
 *   - return value: 0 = ok, <0 = invalid argument (SCB_E_*), >0 = cudaError_t.
 *     scb_last_error() returns a human-readable message for the calling thread.
 *   - no global state apart from the measurement hook (scb_profile_begin / _end);
 *     thread-safe as long as streams and buffers differ.
 *   - image axis 0 ("w", rows) is the particle's x, axis 1 ("h", columns,
 *     contiguous in memory) is y  (_epifm.py:225-231).
 */
#ifndef SCOPYON_B200_H
#define SCOPYON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_VERSION 100

#define SCB_E_INVALID   (-1)   /* bad argument value                          */
#define SCB_E_NULL      (-2)   /* required pointer is NULL                    */
#define SCB_E_WORKSPACE (-3)   /* workspace too small                         */
#define SCB_E_UNSUPPORTED (-4)

/* PSF model, fluorophore.type (_epifm.py:62-74) */
#define SCB_PSF_BORN_WOLF 0
#define SCB_PSF_GAUSSIAN  1

/* detector.type (_epifm.py:1445-1461) */
#define SCB_DET_CMOS  0
#define SCB_DET_EMCCD 1
#define SCB_DET_CCD   2

/* analog_to_digital_converter.type (_epifm.py:941-954) */
#define SCB_FPN_NONE   0
#define SCB_FPN_PIXEL  1
#define SCB_FPN_COLUMN 2

/* element type of image buffers */
#define SCB_F32 0
#define SCB_F64 1

/* Camera grid + PSF-table geometry (_epifm.py:1175-1176, 57-60, 226-231). */
typedef struct scb_geometry {
    int32_t n_w, n_h;        /* detector.image_size                                   */
    int32_t n_radial;        /* radial samples: len(arange(0, radial_cutoff, 1 nm))   */
    int32_t n_depth_keys;    /* integer-nm depth keys 0..n_depth_keys-1; the frozen   */
                             /* beyond-cutoff table ("key -1") is key n_depth_keys    */
    int32_t sat_modulus;     /* phase count M of the SAT block layout (see below), >= 1 */
    int32_t reserved;
    double pixel_length;     /* detector.pixel_length / magnification  [m]            */
    double resolution;       /* table sample pitch, 1e-9 m                            */
    double depth_cutoff;     /* fluorophore.depth_cutoff [m]                          */
    double focal[3];         /* detector.focal_point as (depth, x, y) [m]             */
    double box_peak;         /* largest box-table entry * resolution^2 * inv_scale over the tables in use =  */
                             /* largest fraction of a spot's photons on one pixel; sizes the 32-bit          */
                             /* accumulators of the fp32 render (0 = unknown: treated as 1)                  */
} scb_geometry;

/* Photophysics scalars (_epifm.py:1281-1315, 1343-1360, 1486-1491); all config-only
 * values are evaluated once on the host in the reference's operation order. */
typedef struct scb_photophysics {
    double amplitude0;         /* snells_law() amplitude at depth 0                    */
    double penetration_depth;  /* snells_law() depth; +inf for epi-illumination        */
    double x_sec;              /* ln10 * abs_coeff * 0.1 / N_A                         */
    double quantum_yield;
    double absorb_frac;        /* 1 - 10^(-A)                                          */
    double norm_scale;         /* sum(fluoem_norm) * psf_normalization                 */
    double budget_scale;       /* half_life / ln2 * N_emit0; <=0: photobleaching off   */
} scb_photophysics;

/* Detector + ADC scalars (_epifm.py:1430-1484, 926-963). */
typedef struct scb_detector {
    int32_t type;              /* SCB_DET_*                                            */
    int32_t fpn_type;          /* SCB_FPN_*                                            */
    int32_t bit;               /* ADC bit depth                                        */
    int32_t background_on;
    double qe;                 /* detector.QE                                          */
    double background;         /* effects.background.mean [photons/pixel]              */
    double readout_noise;      /* e-, Gaussian sigma (EMCCD, CCD)                      */
    double emgain;             /* EMCCD multiplication gain                            */
    double fullwell;           /* e-                                                   */
    double adc_offset;         /* ADC0 [counts]                                        */
    double fpn_count;          /* fixed-pattern-noise sigma [counts]                   */
} scb_detector;

/* CMOS read-noise table as a Walker alias table (replaces the per-pixel
 * rng.choice over catalog/detector/RNDist_F40.csv, _epifm.py:332-345). */
typedef struct scb_alias_entry {
    float value;        /* electrons if the coin lands below `threshold`   */
    float alias_value;  /* electrons otherwise                             */
    float threshold;    /* in [0,1]                                        */
    float pad;
} scb_alias_entry;

int scb_version(void);
const char *scb_last_error(void);

/* ---- PSF tables -------------------------------------------------------------- */

/* Radial PSF profile psf(r, z) on r = 0,1,..,n_radial-1 nm for each table depth.
 * Replaces PointSpreadingFunction.get_distribution:
 *   Born-Wolf  _epifm.py:181-213 (NA = 1.4, 100-term rho sum), Gaussian _epifm.py:133-134.
 * d_depths[n_keys] (m) -> d_radial[n_keys][n_radial] (1/m^2). */
int scb_psf_radial_build(int psf_type, double wave_length, double radial_width,
                         int n_radial, int n_keys, const double *d_depths,
                         double *d_radial, void *stream);

/* Bytes of scratch scb_psf_sat_build needs. */
size_t scb_psf_sat_workspace_bytes(int n_radial, int n_keys);

/* Fixed-point summed-area tables of the Cartesian PSF table.
 * Replaces radial_to_cartesian (_epifm.py:98-126) and turns the per-pixel slice sums
 * of overlay_signal_ (_epifm.py:277-282) into four table reads:
 *   T[a][b]  = lerp(radial, min(sqrt((a-c)^2+(b-c)^2), c)),  c = n_radial-1, a,b in [0, 2c]
 *   Q[a][b]  = llrint(T[a][b] * scale_k),  scale_k a power of two chosen per table
 *   S[a][b]  = sum_{a'<a, b'<b} Q[a'][b']                     a,b in [0, 2c+1]
 * Memory layout ("phase blocks"): the pixel edges of one footprint are ~pixel_length/1nm
 * samples apart on both axes.  With M = sat_modulus = round(pixel_length / 1 nm), entry (a, b)
 * is stored in block (a % M, b % M) at slot (a / M, b / M):
 *   d_sat[key][ (((a % M) * M + b % M) * B + a / M) * B + b / M ],
 * B = scb_psf_sat_slots() >= ceil((2c+2) / M) + 1; slots past the last sample repeat the last
 * row / column (value S[min(ir*M+pr, 2c+1)][min(ic*M+pc, 2c+1)]), so the clamped closing edge of a
 * footprint is the next slot.  All corners of a footprint are then one dense rectangle of
 * one B x B block: contiguous B*8-byte rows (two 128-byte lines for B = 32) that the render
 * kernel moves with TMA bulk copies (M = 1 is the plain row-major layout).
 * d_sat[n_keys][scb_psf_sat_table_entries()] (int64), d_inv_scale[n_keys] = 1/scale_k.
 * d_box (optional, same shape, box_type SCB_F64 or SCB_F32): the "box table" -- entry (r, c) of
 * block (pr, pc) is the box sum between slots (r-1, r) x (c-1, c) of that block (slot -1 = 0)
 * rounded to box_type, i.e. the integrated PSF of one pixel for a footprint whose pixel edges
 * have phases (pr, pc).  The render kernel reads it for footprints with evenly spaced edges
 * (whole-nanometre pixel pitch); all other footprints are summed from d_sat and rounded the
 * same way, so both give bit-identical pixel values.  SCB_F32 halves the table and the DRAM
 * traffic of the render at 6e-8 relative accuracy per pixel (meant for SCB_F32 frames). */
int64_t scb_psf_sat_table_entries(int n_radial, int sat_modulus);
int scb_psf_sat_slots(int n_radial, int sat_modulus);
int scb_psf_sat_build(const double *d_radial, int n_radial, int n_keys, int sat_modulus,
                      int64_t *d_sat, void *d_box, int box_type, double *d_inv_scale,
                      void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- particles --------------------------------------------------------------- */

/* Brownian displacement, replaces move_points (sampling.py:88-119) and
 * sampling2.__move_points (sampling2.py:33-50):  x[d] += N(0, sigma[d]) for n_steps
 * consecutive steps, sigma[d] = sqrt(2 D[d] dt).  Normals come from Philox4x32-10
 * keyed (seed; particle, step) so any shard of particles / steps reproduces the same
 * trajectory.  d_x/d_y/d_z: SoA coordinates (NULL for an absent axis), updated in place
 * unless d_out_* are given.  d_state + d_sigma_state[n_states]: per-state isotropic
 * sigma (sampling2); otherwise sigma_axis[3].  periodic != 0 wraps into [lower, upper)
 * per axis after each step (sampling2.py:43-49). */
int scb_diffuse(uint64_t seed, uint64_t first_step, int n_steps, int64_t n, int64_t first_particle,
                const double *d_x, const double *d_y, const double *d_z,
                double *d_out_x, double *d_out_y, double *d_out_z,
                const double sigma_axis[3],
                const int32_t *d_state, const double *d_sigma_state, int n_states,
                int periodic, const double lower[3], const double upper[3],
                void *stream);

/* State transitions of sampling2.__transition_states (sampling2.py:52-68, with the
 * intended side='left'): next = searchsorted(cumsum(P[state]), u).  d_pacc is the
 * row-wise cumulative transition matrix [n_states][n_states]. */
int scb_transition_states(uint64_t seed, uint64_t step, int64_t n, int64_t first_particle,
                          int32_t *d_state, const double *d_pacc, int n_states, void *stream);

/* Uniform initial placement, replaces sample_points (sampling.py:78-83):
 * coord[d] = lower[d] + u * (upper[d]-lower[d]),  u from Philox keyed (seed; particle). */
int scb_place_uniform(uint64_t seed, int64_t n, int64_t first_particle,
                      double *d_x, double *d_y, double *d_z,
                      const double lower[3], const double upper[3], void *stream);

/* Photon emission, photobleaching budget and per-spot PSF weight for one snapshot.
 * Replaces the scalar part of __overlay_molecule_plane (_epifm.py:1281-1315,1321-1333):
 *   amplitude = amplitude0 * exp(-|depth - focal[0]| / penetration_depth)
 *   N_emit    = QY * (amplitude * x_sec * unit_time) * absorb_frac
 *   budget    : first sight -> Exp(1) * budget_scale (Philox keyed (seed; molecule id));
 *               budget -= N_emit; if <= 0: budget = 0, p_state = 0
 *   weight    = norm_scale * (p_state * N_emit / 4 pi)
 * d_budget may be NULL (no bleaching: form_image).  d_mol_slot maps a particle row to its
 * slot in d_budget / d_true_data (NULL: identity); d_mol_id is the id used as RNG key
 * (NULL: identity).  d_true_data[n_slots][8] accumulates the reference's per-molecule
 * vector (_epifm.py:1324-1333) when non-NULL. */
int scb_emit_bleach(uint64_t seed, int64_t n,
                    const double *d_depth, const double *d_x, const double *d_y,
                    const double *d_p_state,
                    const int32_t *d_mol_slot, const int64_t *d_mol_id,
                    double unit_time, double focal_depth, const scb_photophysics *phys,
                    double *d_budget, double *d_weight, double *d_true_data,
                    void *stream);

/* Advance a device-resident movie by n_frames WITHOUT rendering: per frame, photon
 * emission + budget depletion at the current position (as scb_emit_bleach, unit_time =
 * exposure), then one Brownian step (as scb_diffuse, step index = frame index).  This is
 * the trajectory/budget prefix a rank replays before its own block of frames when a
 * movie is partitioned by frame blocks across GPUs (generate_frames is strictly
 * sequential in the reference, _epifm.py:1045-1049; the counter-based RNG removes that
 * dependency).  d_depth/d_x/d_y and d_budget are updated in place. */
int scb_replay_frames(uint64_t diffuse_seed, uint64_t budget_seed, uint64_t first_frame, int64_t n_frames,
                      int64_t n, int64_t first_particle,
                      double *d_depth, double *d_x, double *d_y,
                      const double sigma_dxy[3], double unit_time, double focal_depth,
                      const scb_photophysics *phys, double *d_budget, void *stream);

/* scb_replay_frames that also records what every frame is rendered from: per frame the
 * positions before that frame's Brownian step and the PSF weight of every molecule (as
 * scb_emit_bleach computes it), in [n_frames][n] arrays.  Feeds scb_render_expected_frames. */
int scb_movie_frames(uint64_t diffuse_seed, uint64_t budget_seed, uint64_t first_frame, int64_t n_frames,
                     int64_t n, int64_t first_particle,
                     double *d_depth, double *d_x, double *d_y,
                     const double sigma_dxy[3], double unit_time, double focal_depth,
                     const scb_photophysics *phys, double *d_budget,
                     double *d_out_depth, double *d_out_x, double *d_out_y, double *d_out_weight,
                     void *stream);

/* ---- rendering --------------------------------------------------------------- */

/* Upper bound of scratch bytes scb_render_expected needs for n_spots spots. */
size_t scb_render_workspace_bytes(const scb_geometry *geom, int64_t n_spots);

/* Expected (pre-noise) photon image of one frame.  Replaces get_molecule_plane +
 * PointSpreadingFunction.overlay_signal_ (_epifm.py:1262-1264, 224-282) with identical
 * pixel-edge index arithmetic (IEEE fp64, same operation order):
 *   out[i][j] (+)= sum_spots  weight * 1e-18 * box_sum(T_key; left_i..left_{i+1}, top_j..top_{j+1})
 * Spots are binned to 8 x 128-pixel screen strips by a counting sort; one warp renders one
 * strip with 64-bit fixed-point accumulators in shared memory (no atomics on the image, result
 * independent of the order of the spots).  Footprints with evenly spaced pixel edges are read
 * from d_box by TMA bulk copies (d_box may be NULL: every footprint then takes the SAT path);
 * the others gather four corners per pixel from d_sat.  d_slot_of_key[n_depth_keys+1] maps a
 * depth key to its table index in d_sat / d_box (-1: table absent -> the spot is counted in
 * *d_errors).  out_type: SCB_F32 / SCB_F64; accumulate != 0 adds to the existing image. */
int scb_render_expected(const scb_geometry *geom, int64_t n_spots,
                        const double *d_depth, const double *d_x, const double *d_y,
                        const double *d_weight,
                        const int64_t *d_sat, const void *d_box, int box_type, const double *d_inv_scale,
                        const int32_t *d_slot_of_key,
                        void *d_out, int out_type, int accumulate,
                        void *d_workspace, size_t workspace_bytes,
                        int32_t *d_errors, void *stream);

/* The same two calls on particle ROWS as the host API holds them -- (n, 5) float64 rows
 * (depth, x, y, molecule id, p_state), the output of EPIFMSimulator.__format_data
 * (base.py:61-110) -- read in place, no transposition: what the particle loop of
 * _EPIFMSimulator.output_frame (_epifm.py:1183-1202, 1262-1264) becomes for one
 * (unit_time, particles) snapshot.  The budget RNG key is the id column. */
int scb_emit_bleach_rows(uint64_t budget_seed, int64_t n, const double *d_rows,
                         const int32_t *d_mol_slot, double unit_time, double focal_depth,
                         const scb_photophysics *phys, double *d_budget, double *d_weight,
                         double *d_true_data, void *stream);
int scb_render_expected_rows(const scb_geometry *geom, int64_t n, const double *d_rows,
                             const double *d_weight, const int64_t *d_sat, const void *d_box, int box_type,
                             const double *d_inv_scale, const int32_t *d_slot_of_key, void *d_out,
                             int out_type, int accumulate, void *d_workspace, size_t workspace_bytes,
                             int32_t *d_errors, void *stream);

/* A block of n_frames images in one call: frame f is rendered from spots
 * [f * n_per_frame, (f + 1) * n_per_frame) of the arrays into image f of d_out
 * ([n_frames][n_w][n_h]).  Every kernel of the pipeline runs once for the whole block. */
size_t scb_render_frames_workspace_bytes(const scb_geometry *geom, int64_t n_per_frame, int n_frames);
int scb_render_expected_frames(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                               const double *d_depth, const double *d_x, const double *d_y,
                               const double *d_weight, const int64_t *d_sat, const void *d_box,
                               int box_type, const double *d_inv_scale, const int32_t *d_slot_of_key,
                               void *d_out, int out_type, void *d_workspace, size_t workspace_bytes,
                               int32_t *d_errors, void *stream);
/* The same on particle ROWS: frame f is the (n_per_frame, 5) float64 rows at d_rows + f * n_per_frame * 5
 * (the snapshots of consecutive frames as EPIFMSimulator.__format_data makes them, base.py:61-110, uploaded
 * back to back); d_weight[n_frames][n_per_frame] from scb_emit_bleach_rows.  generate_frames
 * (_epifm.py:1017-1049) renders blocks of frames through it when every frame is one snapshot. */
int scb_render_expected_rows_frames(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                    const double *d_rows, const double *d_weight, const int64_t *d_sat,
                                    const void *d_box, int box_type, const double *d_inv_scale,
                                    const int32_t *d_slot_of_key, void *d_out, int out_type,
                                    void *d_workspace, size_t workspace_bytes, int32_t *d_errors, void *stream);
/* The same with a visiting order: slot s of every frame's spot list shows particle d_order[s]
 * (a permutation of 0 .. n_per_frame - 1, or NULL).  Images do not depend on it (integer accumulation);
 * when it lists the particles tile by tile -- e.g. sorted by a coarse screen cell, refreshed every few
 * blocks, molecules move about a pixel per frame -- the 32 spots of a warp share most of their strips and
 * the census of the binning adds them with a tenth of the atomics. */
int scb_render_expected_frames_ordered(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                       const int32_t *d_order, const double *d_depth, const double *d_x,
                                       const double *d_y, const double *d_weight, const int64_t *d_sat,
                                       const void *d_box, int box_type, const double *d_inv_scale,
                                       const int32_t *d_slot_of_key, void *d_out, int out_type,
                                       void *d_workspace, size_t workspace_bytes, int32_t *d_errors,
                                       void *stream);

/* The same for consecutive blocks of one movie: the binning in one pass.  scb_render_expected_frames_ordered counts the
 * (spot, strip) overlaps, scans the counts into list segments and fills the lists in a second pass over the spots.
 * When the workspace still holds the LIST PLAN a previous call left behind -- every strip's room, a quarter above
 * what that block's census counted: molecules move about a pixel per frame -- the census, the hand-out of list
 * positions and the writing of the units need no scan between them (plan_mode bit 0: use the plan in the workspace;
 * bit 1: leave a plan behind for the next call).  The caller sets bit 0 only when the previous call on this workspace
 * had bit 1 set and the same geometry, n_per_frame and n_frames; a unit beyond its strip's room goes to a short
 * overflow list the render also reads, so the images never depend on the plan: they are those of
 * scb_render_expected_frames_ordered, bit for bit.  More than 65536 overflowing units are counted in d_errors (the
 * caller renders that block again with plan_mode 0 or 2). */
int scb_render_expected_frames_planned(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                       const int32_t *d_order, const double *d_depth, const double *d_x,
                                       const double *d_y, const double *d_weight, const int64_t *d_sat,
                                       const void *d_box, int box_type, const double *d_inv_scale,
                                       const int32_t *d_slot_of_key, void *d_out, int out_type,
                                       void *d_workspace, size_t workspace_bytes, int32_t *d_errors,
                                       int plan_mode, void *stream);

/* Tensor-core variant of scb_render_expected for the separable Gaussian PSF
 * (fluorophore.type == 'Gaussian', _epifm.py:133-134): a 128 x 128 screen tile is the
 * contraction D[i][j] = sum_s (w_s Ex_s(i)) Ey_s(j) over the spots binned to it, issued as
 * tcgen05.mma.kind::tf32 (3 x tf32 split operands, fp32 accumulator in TMEM).  Ex/Ey are
 * differences of d_prefix[0 .. 2c+1], the prefix sums of the 1-D Gaussian on the 1-nm grid
 * (times 1 nm), staged into shared memory by a TMA bulk copy.  Agrees with the reference
 * table to <= 1e-5 of the image maximum (the reference interpolates the radial profile
 * linearly, which is not exactly separable); the SAT path stays exact.  Requires
 * pixel_length >= ~33 nm (footprint <= 64 pixels), else returns SCB_E_UNSUPPORTED. */
size_t scb_gaussian_tc_workspace_bytes(const scb_geometry *geom, int64_t n_spots);
int scb_render_gaussian_tc(const scb_geometry *geom, int64_t n_spots,
                           const double *d_x, const double *d_y, const double *d_weight,
                           const double *d_prefix,
                           void *d_out, int out_type, int accumulate,
                           void *d_workspace, size_t workspace_bytes,
                           int32_t *d_errors, void *stream);

/* Measurement hook (bench.py): between scb_profile_begin and scb_profile_end every
 * scb_render_expected brackets its tile-render kernel with CUDA events on the stream it
 * was given; scb_profile_end synchronises those events and returns the summed device time
 * and the number of bracketed launches.  Process-wide, not thread-safe; never enabled
 * on the product path. */
int scb_profile_begin(int max_launches);
int scb_profile_end(double *total_ms, int64_t *launches);

/* ---- detector ---------------------------------------------------------------- */

/* ADC offset map with fixed-pattern noise, replaces
 * calculate_analog_to_digital_converter_gain (_epifm.py:926-963):
 * offset = rint(N(ADC0, fpn_count)) per pixel (SCB_FPN_PIXEL, n = n_w*n_h) or per
 * axis-1 index (SCB_FPN_COLUMN, n = n_h); Philox keyed (seed; index). */
int scb_adc_offsets(uint64_t seed, int64_t n, double adc0, double fpn_count,
                    void *d_offset, int elem_type, void *stream);

/* Fused detector pass: background + QE, shot noise (Poisson; EMCCD: Poisson -> Gamma
 * multiplication register, truncated like EMCCD.probability_distribution), readout
 * noise (Gaussian, or CMOS alias table), full-well clip, ADC gain/offset, clip to
 * [0, 2^bit - 1].  Replaces __detector_output + CMOS/EMCCD/CCD + ADC
 * (_epifm.py:1430-1484, 329-433).  One pass over the frame: reads d_photons once,
 * writes d_adc once.  Philox keyed (seed; frame, pixel).
 *   d_photons      expected photons per pixel (render output), elem_type
 *   d_offset       NULL (SCB_FPN_NONE), [n_h] (COLUMN) or [n_w*n_h] (PIXEL), elem_type
 *   d_cmos_alias   alias table [n_alias] (CMOS only)
 *   d_expectation  optional out: QE*(photons+background)          (camera[:,:,0])
 *   d_in_signal / d_in_noise  optional injected draws (bit-exact ADC tests)
 *   d_out_signal / d_out_noise optional taps of the drawn signal / noise (statistics tests)
 *   d_workspace    scratch of scb_detector_workspace_bytes(): the streaming pass finishes
 *                  every pixel whose expectation is < 12 e- (and, for EMCCD, drew no
 *                  electron) and lists the others; a second pass runs the general
 *                  samplers on that list with the same per-pixel random streams.
 */
size_t scb_detector_workspace_bytes(int32_t n_w, int32_t n_h);
int scb_detector_adc(uint64_t seed, uint64_t frame, const scb_detector *det,
                     int32_t n_w, int32_t n_h, int elem_type,
                     const void *d_photons, const void *d_offset,
                     const scb_alias_entry *d_cmos_alias, int n_alias,
                     void *d_adc, void *d_expectation,
                     const void *d_in_signal, const void *d_in_noise,
                     void *d_out_signal, void *d_out_noise,
                     void *d_workspace, size_t workspace_bytes,
                     void *stream);

/* Frames first_frame .. first_frame + n_frames - 1 of a movie in one pair of launches: fp32 images
 * [n_frames][n_w][n_h] in and out (no taps), n_frames * scb_detector_workspace_bytes() of scratch.
 * The draws of frame f are those scb_detector_adc makes for that frame. */
int scb_detector_adc_frames(uint64_t seed, uint64_t first_frame, int n_frames, const scb_detector *det,
                            int32_t n_w, int32_t n_h, int elem_type,
                            const void *d_photons, const void *d_offset,
                            const scb_alias_entry *d_cmos_alias, int n_alias,
                            void *d_adc, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- finished frames ------------------------------------------------------------ */

/* 8-bit scaling of a frame stack, Image.__as_8bit (image.py:98-123; "same as
 * scipy.misc.bytescale") with the common limits Video.save uses for a movie (image.py:261-264):
 *   out = uint8( ((data - cmin) * (high - low) / (cmax - cmin) + low).clip(low, high) + 0.5 ),
 * low where cmax == cmin; fp64 arithmetic in the reference's order.  scb_frames_minmax reduces n
 * elements to d_minmax[2] = (min, max) (16 bytes of workspace); scb_frames_to_8bit takes the
 * limits from d_limits[2] on the device when given, else from cmin / cmax. */
int scb_frames_minmax(const void *d_frames, int64_t n, int elem_type, double *d_minmax, void *d_workspace,
                      void *stream);
/* 16-bit camera counts of a frame stack for export: out = uint16(rint(clip(v, 0, 65535))), NaN -> 0
 * (what `img.as_array().round().clip(0, 65535).astype(uint16)` gives; the reference's ADC does not
 * quantise, _epifm.py:1472-1484, so this is an export format, not the API's return type). */
int scb_frames_to_u16(const void *d_frames, int64_t n, int elem_type, uint16_t *d_out, void *stream);
int scb_frames_to_8bit(const void *d_frames, int64_t n, int elem_type, const double *d_limits,
                       double cmin, double cmax, double low, double high, uint8_t *d_out, void *stream);

/* ---- spot detection (the step after image formation) ---------------------------- */

/* Laplacian-of-Gaussian scale space of one image, the cube skimage.feature.blob_log builds for
 * scopyon.analysis.blob_detection (analysis/spot_detection.py:14-47):
 *   cube[s] = -scipy.ndimage.gaussian_laplace(image, sigma_s) * sigma_s^2,
 * planes [n_sigma][n_w][n_h] fp64.  d_weights[n_sigma][2][weight_pitch] holds, per scale, the
 * right half (taps 0..radius) of the Gaussian kernel and of its second derivative as
 * scipy.ndimage.gaussian_filter1d computes them; d_radius[n_sigma] the radii, max_radius their
 * maximum (<= scb_log_max_radius(), the shared-memory tile limit), d_sigma2[n_sigma] = sigma^2.
 * Boundary mode 'reflect'.  The sums are formed in scipy's order without fused multiply-adds, so
 * the cube equals scipy's bit for bit.  Workspace: scb_log_workspace_bytes(n_w, n_h, n_sigma);
 * a caller short of memory builds the cube a few scales at a time. */
size_t scb_log_workspace_bytes(int n_w, int n_h, int n_sigma);
int scb_log_max_radius(void);
int scb_log_scale_space(int n_w, int n_h, int n_sigma, const double *d_image, const int32_t *d_radius,
                        int max_radius, const double *d_weights, int weight_pitch, const double *d_sigma2,
                        double *d_cube, void *d_workspace, size_t workspace_bytes, void *stream);

/* Scale-space peaks, skimage.feature.peak_local_max(cube, threshold_abs=threshold,
 * footprint=ones((3,3,3)), exclude_border=False): voxels > threshold that no neighbour in the
 * 3x3x3 box (edges replicated) exceeds.  Appends (i, j, scale index) to d_peaks[capacity][3] and
 * the value to d_values[capacity] in no particular order; *d_count = number found (may exceed
 * capacity: then only the first capacity were stored).  A count equal to the number of voxels
 * means a flat cube, for which peak_local_max reports nothing. */
int scb_log_peaks(int n_w, int n_h, int n_sigma, const double *d_cube, double threshold, int32_t *d_peaks,
                  double *d_values, int64_t capacity, unsigned long long *d_count, void *stream);

/* Per-blob background plane and Gaussian fit, scopyon.analysis.spot_detection's worker
 * (analysis/spot_detection.py:110-137).  d_blobs[n_blobs][blob_stride] starts with (x, y); the ROI
 * is rows int(x - roi_size) .. int(x + roi_size), columns likewise, clipped to the image
 * (roi_size <= 15).  d_spots[n_blobs][6] = (center_x, center_y, intensity, bg, height, sigma)
 * where d_status[blob] == 0; status 1 = no signal in the ROI, 2 = no background plane,
 * 3 = fit did not converge within max_iterations, 4 = fitted centre outside the ROI: the blobs
 * the reference skips.  The least-squares minimum is the one scipy.optimize.least_squares
 * converges to from the same start; the two agree to the optimiser's tolerance, not bitwise. */
int scb_spot_fit(int n_w, int n_h, const double *d_image, int64_t n_blobs, const double *d_blobs,
                 int blob_stride, double roi_size, int max_iterations, double *d_spots, int32_t *d_status,
                 void *stream);

/* ---- host side of the end-to-end path ---------------------------------------------------
 * scopyon hands frames to the caller as float64 (Nw, Nh) arrays (image.py:12-36,
 * _epifm.py:1177); the fp32 pipeline downloads them as float32 into pinned staging memory
 * (half the PCIe bytes) and widens them here, on a small pool of host threads, while the next
 * frames are in flight.  (double)float is exact: the result equals a device-side widening.
 * scb_host_widen_start queues src[n] -> dst[n]; the work starts once `cuda_event` (the event
 * recorded after the download on `device`; NULL = the source is ready) has completed.  Returns a
 * ticket > 0 (or -1); tickets finish in order.  scb_host_widen_wait blocks until the ticket is
 * done: 0, or the CUDA error the download ended with.  scb_host_widen_threads sets the number
 * of worker threads (1..64, applied when the queue is idle) and returns the previous one.
 * Pure host code: no kernel is launched. */
int64_t scb_host_widen_start(const float *h_src, double *h_dst, int64_t n, void *cuda_event, int device);
int scb_host_widen_wait(int64_t ticket);
int scb_host_widen_threads(int n_threads);
/* Pins worker w to cpus[w % n_cpus] (n_cpus = 0: no pinning), applied when the queue is idle: the ranks
 * of one box (one process per GPU) give their workers disjoint cores next to their GPU. */
int scb_host_widen_affinity(const int *cpus, int n_cpus);
/* Host memory bandwidth with the access patterns of this path, n_threads threads on private buffers,
 * best of `repeats`: mode 0 = copy with streaming stores (read b + write b), mode 1 = the float32 ->
 * float64 widening above (read b + write 2b).  *bytes_per_s = bytes read + written per second: the
 * bound the end-to-end frame rate of a box is reported against (bench.py).  Pure host code. */
int scb_host_bandwidth(int mode, int64_t bytes_per_thread, int n_threads, int repeats, double *bytes_per_s);

/* ---- device properties / test hooks ------------------------------------------------------ */

/* Multiprocessors of the current device (grids of the persistent kernels are sized from it). */
int scb_device_sm_count(void);
/* Known-answer hooks for the tests; not used by the product path.  scb_philox4x32_10: the counter-based
 * generator on the host (Random123 vectors).  scb_test_poisson_inversion: the detector's inversion
 * sampler (expectation < 12 photoelectrons) on n given (expectation, 32-bit word) pairs -> counts. */
void scb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int scb_test_poisson_inversion(int64_t n, const float *d_lambda, const uint32_t *d_word, float *d_count,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCOPYON_B200_H */
