/*
 * beatgpu.h -- C-ABI of libbeatgpu.so: the B200 (sm_100a) implementation of BEAT's per-chain
 * forward model + log-likelihood hot path, batched over SMC/PT chains.
 *
 * This is the drop-in boundary.  Plain pointers and sizes only; no torch,
 * numpy or CUDA types in any signature.  Every entry point names the reference interface it
 * replaces (paths relative to the hvasbath/beat tree, v2.0.5).  The reference side binds this
 * with ctypes (our own binding is beat_b200/lib.py).
 *
 * Conventions
 *  - all functions return 0 on success, a BEATGPU_E_* code otherwise; beatgpu_last_error()
 *    gives the message.  Nothing is clamped or silently repaired: out-of-range library indices
 *    and non-finite results are reported (the reference would raise IndexError / ValueError,
 *    beat/ffi/base.py:650-674, beat/sampler/metropolis.py:279-284).
 *  - "host" entry points take HOST pointers, copy in/out and synchronise before returning.
 *    "_dev" entry points take DEVICE pointers, enqueue on the context's stream and do not
 *    synchronise (call beatgpu_sync()).
 *  - arrays are C-contiguous (row-major), float64 unless said otherwise; B = number of chains.
 *  - a context belongs to one (process, GPU); calls on one context are serialised on its stream;
 *    it is not fork-safe (the reference forks workers, beat/parallel.py:243 -- the batched path
 *    runs with n_jobs=1, one process per GPU).
 *  - the library never frees or retains caller memory; all device scratch is owned by the ctx and
 *    grown on demand (no cudaMalloc on the steady-state path once B is stable).
 */
#ifndef BEATGPU_H
#define BEATGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEATGPU_VERSION 100          /* 0.1.0 */
#define BEATGPU_MAX_SLIPVARS 3       /* uparr, uperp, utens  (beat/config.py:83-94) */

/* status codes */
#define BEATGPU_OK            0
#define BEATGPU_E_CUDA        1      /* CUDA runtime error (message has the detail)          */
#define BEATGPU_E_ARG         2      /* bad argument / call order                             */
#define BEATGPU_E_INDEX       3      /* a GF-library index left the library (IndexError)      */
#define BEATGPU_E_NOTREADY    4      /* an operand required by the call has not been uploaded */
#define BEATGPU_E_NONFINITE   5      /* reserved: non-finite llk (ValueError in the sampler)  */
#define BEATGPU_E_IO          6      /* file I/O of the trace writer failed (OSError)         */

/* storage dtype of a GF library on the device */
#define BEATGPU_F32 0
#define BEATGPU_F64 1

/* interpolation of rupture start time / rise time into the library
 * (beat/config.py:571-575; beat/ffi/base.py:506-517,553-564,649-698) */
#define BEATGPU_NEAREST      0
#define BEATGPU_MULTILINEAR  1

typedef struct beatgpu_ctx beatgpu_ctx;

/* ---------------------------------------------------------------- context ------------------ */
int         beatgpu_version(void);
int         beatgpu_ctx_create(int device, beatgpu_ctx** out);
void        beatgpu_ctx_destroy(beatgpu_ctx* ctx);
const char* beatgpu_last_error(const beatgpu_ctx* ctx);    /* ctx may be NULL: last create() error */
int         beatgpu_sync(beatgpu_ctx* ctx);                 /* cudaStreamSynchronize on the ctx stream */
/* external != 0: enqueue on the caller's CUDA stream `cuda_stream` (e.g. torch's current stream; the handle 0 is
 * the default stream) so the library's kernels are ordered with the caller's work on that stream;
 * external == 0: go back to a private non-blocking stream (cuda_stream ignored).                       */
int         beatgpu_set_stream(beatgpu_ctx* ctx, void* cuda_stream, int external);
/* Page-lock a caller-owned host buffer (cudaHostRegister) so the host-pointer entries copy it at full PCIe rate;
 * a sampler registers its parameter / output arrays once and reuses them every step.  The memory stays the
 * caller's; unregister before freeing it.                                                                   */
int         beatgpu_host_register(beatgpu_ctx* ctx, void* ptr, int64_t bytes);
int         beatgpu_host_unregister(beatgpu_ctx* ctx, void* ptr);
/* number of SMs, device name; for sizing/reporting */
int         beatgpu_device_info(beatgpu_ctx* ctx, int* n_sm, char* name, int name_len);

/* ---------------------------------------------------------------- static operands ----------
 * These replace the pytensor shared variables the reference builds once in
 * Composite.get_formula (beat/models/seismic.py:1210-1349, geodetic.py:1030-1084,
 * laplacian.py:40-62) and updates between SMC stages (seismic.py:1509-1534).               */

/* Fault discretisation: per subfault the patch grid and (square) patch size [km].
 * Replaces FaultGeometry.ordering / Sweeper.__init__ (beat/pytensorf.py:426-430,
 * beat/models/seismic.py:1097-1111).  Patch order inside a subfault is dip-major:
 * flat = dip * n_patch_strike + strike (beat/pytensorf.py:475-482).                          */
int beatgpu_set_fault(beatgpu_ctx* ctx, int n_subfaults, const int32_t* n_patch_dip,
                      const int32_t* n_patch_strike, const double* patch_size_km);

/* How a chain's flat parameter vector q[n_params] maps onto the model variables; replaces pymc's
 * DictToArrayBijection over value_vars (beat/backend.py:147,163-165; variable list
 * beat/config.py:1506-1542).  An offset of -1 means "not sampled": the value comes from
 * `fixed` (the reference's fixed_rvs, seismic.py:1243).  Lengths are implied by the fault
 * (npatches, n_subfaults) and by n_hypers / n_time_shifts.                                   */
typedef struct beatgpu_layout {
    int32_t n_params;
    int32_t n_slipvars;                        /* 1..3, order = slip_varnames                 */
    int32_t off_slip[BEATGPU_MAX_SLIPVARS];    /* each [npatches]                             */
    int32_t off_durations;                     /* [npatches]                                  */
    int32_t off_velocities;                    /* [npatches]                                  */
    int32_t off_nucleation_strike;             /* [n_subfaults]                               */
    int32_t off_nucleation_dip;                /* [n_subfaults]                               */
    int32_t off_time;                          /* [n_subfaults]                               */
    int32_t off_hypers;                        /* [n_hypers]  (h_* in canonical order)        */
    int32_t n_hypers;
    int32_t off_time_shifts;                   /* [n_time_shifts] hierarchical station corr.  */
    int32_t n_time_shifts;
} beatgpu_layout;
/* fixed: canonical vector [nslip*np | np | np | nsf | nsf | nsf | n_hypers | n_time_shifts] used for
 * every variable whose offset is -1 (may be NULL if nothing is fixed).                        */
int beatgpu_set_layout(beatgpu_ctx* ctx, const beatgpu_layout* layout, const double* fixed);

/* Declare one seismic wavemap (a WaveformMapping: beat/heart.py:2884-3148) with n_targets
 * datasets of n_samples each.  station_idx[n_targets] maps a target to its entry of the
 * time_shifts hierarchical (wmap.station_correction_idxs, seismic.py:1283-1291) or is NULL.
 * hyper_idx[n_targets] indexes the hypers block (hp_specific or not: distributions.py:121-126).
 * nsamples[n_targets] is dataset.samples (M, distributions.py:120).  Returns the wavemap id in
 * *wmap_id; datasets are appended to the output row in declaration order (seismic.py:1348).   */
int beatgpu_add_wavemap(beatgpu_ctx* ctx, int n_targets, int n_samples, int interpolation,
                        const int32_t* station_idx, const int32_t* hyper_idx,
                        const int32_t* nsamples, int* wmap_id);

/* Upload one slip component's GF library: traces laid out (ntargets, npatches, ndurations,
 * nstarttimes, nsamples) C-order exactly as the reference's `.traces.npy`
 * (beat/ffi/base.py:161-189,378) of dtype src_dtype, stored on the device as store_dtype
 * (BEATGPU_F64 = strict parity mode, BEATGPU_F32 = half the bytes).  Axis origin/step are
 * SeismicGFLibraryConfig.duration_min/duration_sampling/starttime_min/starttime_sampling
 * (beat/config.py:1900-1919).  `traces` may be a memory-mapped file; it is streamed in chunks.
 * Replaces SeismicGFLibrary.init_optimization (beat/ffi/base.py:387-404).                    */
int beatgpu_upload_gflib(beatgpu_ctx* ctx, int wmap_id, int slipvar, const void* traces,
                         int src_dtype, int store_dtype, const int64_t dims[5],
                         double duration_min, double duration_step,
                         double starttime_min, double starttime_step);
/* Allocate the library on the device without filling it and hand back the raw device pointer
 * (store_dtype elements, row stride *row_stride elements) so a caller can generate / copy
 * a library device-side (bench: synthetic libraries are built directly in HBM).              */
int beatgpu_alloc_gflib(beatgpu_ctx* ctx, int wmap_id, int slipvar, int store_dtype,
                        const int64_t dims[5], double duration_min, double duration_step,
                        double starttime_min, double starttime_step,
                        void** device_ptr, int64_t* row_stride);

/* Observed data, one row per target: wmap.shared_data_array (beat/heart.py:3126-3133).      */
int beatgpu_upload_data(beatgpu_ctx* ctx, int wmap_id, const double* data /*[nt, ns]*/);

/* Weights = Covariance.chol_inverse (upper-triangular U with U^T U = C^-1, beat/heart.py:211-237)
 * and slog_pdet = Covariance.log_pdet (:239-245) for every target of the wavemap; called at
 * setup (SeismicComposite.init_weights, seismic.py:363-378) and again between SMC stages
 * (update_weights, seismic.py:1527-1534).  The library inspects U: entries whose magnitude is
 * below band_rtol * max|U| are treated as structural zeros, and U is stored as diagonal,
 * upper-banded (Toeplitz `exponential` noise gives bandwidth 1) or dense.  band_rtol < 0
 * selects the default (1e-13); band_rtol = 0 forces exact structure detection.               */
int beatgpu_update_weights(beatgpu_ctx* ctx, int wmap_id, const double* U /*[nt, ns, ns]*/,
                           const double* slog_pdet /*[nt]*/, double band_rtol);

/* Same for weight matrices that were computed ON the device (the per-stage covariance update from the residuals of
 * the MAP point, beat/covariance.py:397-427,716-771 -> heart.Covariance.chol_inverse): U_dev [nt, ns, ns] and
 * slog_pdet_dev [nt] are device pointers; structure detection and repacking run on the device, nothing but three
 * flags crosses PCIe.  Synchronises before returning (the caller may free its buffers).                    */
int beatgpu_update_weights_dev(beatgpu_ctx* ctx, int wmap_id, const double* U_dev, const double* slog_pdet_dev,
                               double band_rtol);

/* Geodetic static composite (beat/models/geodetic.py:1030-1084): one library per slip var,
 * G[var] (npatches, nobs) as GeodeticGFLibrary (beat/ffi/base.py:192-305); data and odw [nobs];
 * n_datasets slices [lo, hi) of the concatenated observation vector (Bij.srmap); U per dataset
 * concatenated (sum n_i^2 doubles), slog_pdet, nsamples, hyper_idx per dataset.             */
int beatgpu_set_geodetic(beatgpu_ctx* ctx, int n_obs, int n_datasets, const int32_t* slice_lo,
                         const int32_t* slice_hi, const double* const* G /*[n_slipvars] each [np, nobs]*/,
                         const double* data, const double* odw, const double* U_concat,
                         const double* slog_pdet, const int32_t* nsamples, const int32_t* hyper_idx);
int beatgpu_update_geodetic_weights(beatgpu_ctx* ctx, const double* U_concat, const double* slog_pdet);

/* Laplacian smoothing prior (beat/models/laplacian.py:40-139): operator L (npatches, npatches),
 * sdet = log_determinant(L.T * L) (:57-60) and the index of h_laplacian in the hypers block.  */
int beatgpu_set_laplacian(beatgpu_ctx* ctx, const double* L, double sdet, int hyper_idx);

/* Number of outputs per chain: seismic datasets of all wavemaps, then geodetic datasets, then one
 * laplacian_like (if set) -- the per-dataset logpts the reference traces store
 * (seis_like / geo_like / laplacian_like, beat/sampler/metropolis.py:160-162).              */
int beatgpu_n_outputs(beatgpu_ctx* ctx, int* n_out);

/* ---------------------------------------------------------------- hot path ----------------- */

/* Batched Sweeper: replaces Sweeper.perform -> fast_sweep_ext.fast_sweep
 * (beat/pytensorf.py:443-500, beat/fast_sweeping/fast_sweep_ext.c:120-206) for B chains of one
 * subfault.  slowness [B, np_sf] (= 1/velocities), nuc_dip_idx / nuc_strike_idx [B] patch indices;
 * out start times [B, np_sf] (no `time` offset added, exactly like the Op).
 * n_iter [B] (optional, may be NULL) receives the outer iteration count.                     */
int beatgpu_fast_sweep_batch(beatgpu_ctx* ctx, int subfault, int B, const double* slowness,
                             const int32_t* nuc_dip_idx, const int32_t* nuc_strike_idx,
                             double* starttimes, int32_t* n_iter);

/* Batched SeismicGFLibrary.stack_all (beat/ffi/base.py:607-709) summed over the wavemap's slip
 * components as the composite does (seismic.py:1317-1330): durations [B, np], starttimes
 * [B, nt, np], slips [n_slipvars, B, np]  ->  synthetics [B, nt, ns].  n_slipvars_used <=
 * uploaded libraries; interpolation as declared for the wavemap.                              */
int beatgpu_stack_batch(beatgpu_ctx* ctx, int wmap_id, int B, int n_slipvars_used,
                        const double* durations, const double* starttimes, const double* slips,
                        double* synthetics);

/* Batched multivariate_normal_chol (beat/models/distributions.py:72-140) for one wavemap:
 * residuals [B, nt, ns], hypers [B, n_hypers]  ->  logpts [B, nt].                             */
int beatgpu_misfit_batch(beatgpu_ctx* ctx, int wmap_id, int B, const double* residuals,
                         const double* hypers, int n_hypers, double* logpts);

/* Same with DEVICE pointers, enqueued on the ctx stream without synchronisation (e.g. the log-likelihood of a
 * geometry-mode problem whose synthetics are produced elsewhere on the device).                          */
int beatgpu_misfit_batch_dev(beatgpu_ctx* ctx, int wmap_id, int B, const double* residuals_dev,
                             const double* hypers_dev, int n_hypers, double* logpts_dev);

/* The fused evaluation: replaces one call of the compiled logp_forw_func(q)
 * (beat/sampler/base.py:598-615; graph of seismic.py:1253-1349 [+ geodetic.py:1065-1084,
 * laplacian.py:98-139]) for each of B chains: q [B, n_params] -> logpts [B, n_out] and
 * like [B] = row sums (beat/models/problems.py:228-247).                                       */
int beatgpu_ffi_loglike_batch(beatgpu_ctx* ctx, int B, const double* q, double* logpts, double* like);
int beatgpu_ffi_loglike_batch_dev(beatgpu_ctx* ctx, int B, const double* q_dev, double* logpts_dev,
                                  double* like_dev);

/* Forward model only: replaces SeismicDistributerComposite.get_synthetics(point, outmode="array")
 * (beat/models/seismic.py:1351-1507) for B chains of one wavemap: q [B, n_params] -> synthetics
 * [B, nt, ns] (rupture sweep + stacking of all slip components, no residual / misfit).  Used between SMC
 * stages to form the residuals the covariance update works on (seismic.py:1509-1534).              */
int beatgpu_ffi_synthetics_batch(beatgpu_ctx* ctx, int wmap_id, int B, const double* q, double* synthetics);

/* After a loglike batch: per-chain rupture start times [B, npatches] (Deterministic-style
 * inspection / parity of the sweep inside the fused path); host pointer.                     */
int beatgpu_get_starttimes(beatgpu_ctx* ctx, int B, double* starttimes);

/* Count of (chain, target, patch) taps that fell outside a library since the last call of this
 * function (the counter is reset).  Affected logpts are NaN.                                  */
int beatgpu_index_violations(beatgpu_ctx* ctx, int64_t* count);

/* Counters for reporting: kernels launched by this ctx since creation.                       */
int beatgpu_launch_count(beatgpu_ctx* ctx, int64_t* n_launches);

/* Time the most recent fused evaluation's dominant kernel (gf stack + misfit) on the ctx stream
 * with CUDA events: milliseconds of the last loglike batch's stack kernel(s).                */
int beatgpu_last_stack_ms(beatgpu_ctx* ctx, float* ms);

/* Sum of the CUDA-event durations [ms] of the dominant kernels (gf stack + misfit pass) of the fused evaluations
 * enqueued since the last reset, and how many evaluations that covers (at most the last 1024).  The event pairs are
 * recorded inside every beatgpu_ffi_loglike_batch[_dev] call on the ctx stream, so a caller can time a whole loop
 * and learn the kernels' share of it afterwards (bench.py: roofline.kernel_ms is measured INSIDE the timed region).
 * Synchronises on the last recorded event.  reset != 0 clears the accumulator after reading.                  */
int beatgpu_stack_ms_accum(beatgpu_ctx* ctx, int reset, double* sum_ms, int64_t* n_evals);

/* How the stacking pass of a wavemap is blocked for the L2 cache: patches per chunk, number of chunks, the library
 * bytes one chunk spans (chunk * n_slipvars * ndurations * nstarttimes * row bytes) and the device's L2 size.  The
 * chunk is derived from cudaDeviceProp.l2CacheSize (working set <= 60 % of L2, nominal bytes; BEATGPU_L2_FRAC / BEATGPU_CHUNK
 * override) so that libraries with larger per-patch blocks keep the all-chains-stream-through-L2 behaviour.   */
int beatgpu_stack_blocking(beatgpu_ctx* ctx, int wmap_id, int n_slipvars, int* chunk_patches, int* n_chunks,
                           int64_t* chunk_bytes, int64_t* l2_bytes);

/* Hash of the CUDA / C++ sources this library was built from (build.py passes it to nvcc): lets a committed ncu
 * traffic record be matched against the build that is actually running.                                      */
const char* beatgpu_source_hash(void);

/* ---------------------------------------------------------------- geometry mode -------------
 * The point-source ("geometry") seismic composite of BASELINE config 2: per chain a double-couple source
 * (east_shift, north_shift, depth [km], strike, dip, rake [deg], magnitude, time [s], STF duration [s]) is turned into
 * synthetic seismograms by a weighted, delayed sum of GF-store traces, filtered, tapered, chopped and compared with
 * the data.  Replaces SeisSynthesizer.perform (beat/pytensorf.py:241-302) -> heart.seis_synthetics
 * (beat/heart.py:3564-3762) -> post_process_trace (:3466-3525) and the likelihood of
 * SeismicGeometryComposite.get_formula (beat/models/seismic.py:737-837).  The synthesis itself is pyrocko's
 * (un-vendored dependency, >= 2023.10.11): its published algorithm is restated, see csrc/geom.cuh and
 * oracle/geom_oracle.py.  Times are relative to the reference event's origin time.  Supported: DC sources (one or
 * several, stacked), HalfSinusoid / Boxcar / Triangular STF, GF component scheme 'elastic10', store type A (source depth x distance), pre_stack_cut
 * = True (the reference's default), time domain, station corrections.  A context is either finite-fault or
 * geometry mode.                                                                                            */

/* Flat parameter vector -> source variables (pymc bijection over value_vars, beat/backend.py:147,163-165; variable
 * list beat/config.py:83-94 for DCSource + 'duration' of the STF, beat/utility.py:773-797).  Offset -1 = not
 * sampled: value taken from fixed[] in the canonical order [east_shift[n_sources], north_shift[..], depth, strike,
 * dip, rake, magnitude, time, duration, hypers..., time_shifts...].                                         */
typedef struct beatgpu_geom_layout {
    int32_t n_params;
    int32_t off_east_shift, off_north_shift, off_depth;      /* [km] (utility.adjust_point_units converts to m) */
    int32_t off_strike, off_dip, off_rake;                   /* [deg]                                            */
    int32_t off_magnitude, off_time, off_duration;
    int32_t off_hypers, n_hypers;
    int32_t off_time_shifts, n_time_shifts;                  /* hierarchical station corrections [s] (time_shifts_<mapid>,
                                                                beat/models/seismic.py:198-294); n = 0: none            */
    int32_t n_sources;                                       /* every source variable is a block of n_sources values
                                                                (pymc vectors of shape (n_sources,), config.py:1506-1542);
                                                                their synthetics are stacked (beat/heart.py:3719-3724); 0 = 1 */
} beatgpu_geom_layout;
/* event_lat/lon: origin the source's north/east shifts refer to (the reference event, beat/config.py:2045-2066);
 * stf_anchor: HalfSinusoidSTF.anchor, -1 in the reference (beat/config.py:2060).                             */
int beatgpu_geom_set_source(beatgpu_ctx* ctx, const beatgpu_geom_layout* layout, const double* fixed,
                            double event_lat, double event_lon, double stf_anchor);

/* Source time function of the geometry-mode sources.  BEAT offers pyrocko's Boxcar, Triangular and HalfSinusoid STFs
 * (stf_catalog, beat/sources.py:723-729; SeismicGeometryConfig.stf_type, beat/config.py:1359-1365, default
 * HalfSinusoid), creates them with anchor = -1 (beat/config.py:2058-2060) and routes the sampled `duration` -- and
 * `peak_ratio` of the triangle (bounds beat/defaults.py:238-240) -- to them (utility.update_source, :773-797).
 * off_peak_ratio: column of q holding peak_ratio (a block of n_sources values) or -1 = fixed_peak_ratio for every chain;
 * ignored unless stf_type is BEATGPU_STF_TRIANGULAR.  Call after beatgpu_geom_set_source (which resets the STF to
 * HalfSinusoid with the anchor given there).                                                                  */
#define BEATGPU_STF_HALFSINUSOID 0
#define BEATGPU_STF_BOXCAR       1
#define BEATGPU_STF_TRIANGULAR   2
int beatgpu_geom_set_stf(beatgpu_ctx* ctx, int stf_type, double stf_anchor, int off_peak_ratio, double fixed_peak_ratio);

/* Upload a GF store (what engine.get_store(target.store_id) opens in the reference, beat/heart.py:3657): dims =
 * (n_source_depths, n_distances, 10, row_length); record (iz, ix, g) holds nsamples[iz,ix,g] <= row_length float32
 * samples starting at sample index itmin[iz,ix,g] (time = index * deltat after the source origin); outside its span
 * a record repeats its first / last sample.  z0/dz, x0/dx: source-depth and distance axes [m].               */
int beatgpu_geom_upload_store(beatgpu_ctx* ctx, const int64_t dims[4], double z0, double dz, double x0, double dx,
                              double deltat, const float* traces, const int32_t* itmin, const int32_t* nsamples,
                              int* store_id);

/* Declare a geometry-mode wavemap of n_targets datasets with n_samples each (the constructor arguments of
 * SeisSynthesizer, beat/pytensorf.py:160-206): target positions [deg] and sensor orientation (azimuth, dip [deg]:
 * N = (0, 0), E = (90, 0), Z = (0, -90)), the fixed phase arrival times [s] (wmap._arrival_times), the ArrivalTaper
 * (a, b, c, d) [s] relative to the arrival (beat/heart.py:266-336), chop bounds as indices into (a, b, c, d)
 * (the likelihood uses 1, 2 = ["b", "c"], seismic.py:755), and the filter as a cascade of IIR sections exactly as
 * scipy.signal.butter returns them (Filter.apply, heart.py:377-392: stepwise = high-pass with demean, then
 * low-pass): sec_b / sec_a are [n_sections, 9] (zero padded), demean_first != 0 removes the trace mean before the
 * first section.  station_idx[n_targets] (or NULL) maps a target to its entry of the time_shifts hierarchical
 * (wmap.station_correction_idxs; SeisSynthesizer.perform adds it to the arrival time, beat/pytensorf.py:248-252):
 * the engine window, the chop and the taper then move with the chain's correction while the data stay put.
 * Targets sharing position and window are synthesised together.  Data and weights are uploaded with
 * beatgpu_upload_data / beatgpu_update_weights under the returned id.                                         */
int beatgpu_geom_add_wavemap(beatgpu_ctx* ctx, int store_id, int n_targets, int n_samples, int interpolation,
                             const double* lats, const double* lons, const double* azimuths, const double* dips,
                             const double* arrival_times, const double taper_abcd[4], int chop_lo, int chop_hi,
                             int n_sections, const int32_t* sec_order, const double* sec_b, const double* sec_a,
                             int demean_first, const int32_t* station_idx, const int32_t* hyper_idx,
                             const int32_t* nsamples, int* wmap_id);

/* One evaluation of the compiled logp_forw_func(q) of a geometry-mode seismic problem for each of B chains:
 * q [B, n_params] -> logpts [B, n_out], like [B].  Chains whose source leaves the GF store are reported
 * (BEATGPU_E_INDEX; their logpts are NaN) -- the reference raises (beat/heart.py:3659-3662).                */
int beatgpu_geom_loglike_batch(beatgpu_ctx* ctx, int B, const double* q, double* logpts, double* like);
int beatgpu_geom_loglike_batch_dev(beatgpu_ctx* ctx, int B, const double* q_dev, double* logpts_dev, double* like_dev);

/* Forward model only: heart.seis_synthetics(..., outmode="array") for B chains: synthetics [B, nt, ns].        */
int beatgpu_geom_synthetics_batch(beatgpu_ctx* ctx, int wmap_id, int B, const double* q, double* synthetics);

/* GF-store bulk copies that timed out in beatgpu_geom_loglike_batch_dev calls since the last query (the host-pointer
 * entries report theirs as BEATGPU_E_CUDA).  The chains concerned already carry NaN logpts (rejected by a sampler);
 * a non-zero count means the device lost copies.  Reads and resets the counter; synchronises.                  */
int beatgpu_geom_timeouts(beatgpu_ctx* ctx, int64_t* count);

/* ---------------------------------------------------------------- trace files (host only) ----
 * Append n_steps records of each of n_chains chains to <dir_path>/chain-<chain_offset + c>.bin -- what
 * NumpyChain.record_buffer does per chain with ndarray.tofile on a file opened in append mode
 * (beat/backend.py:822-845; file layout :765-820).  records: [n_steps, n_chains, rec_bytes], step-major (what a lock-step
 * sampler produces: the outputs of all chains per step); every chain's records are written with one writev straight
 * from that buffer, chains split over n_threads native threads.  The files must exist with their header
 * (BatchedNumpyChains.setup).  No CUDA context involved; errors: BEATGPU_E_IO, message via beatgpu_last_error(NULL)
 * on the calling thread.                                                                                       */
int beatgpu_trace_append(const char* dir_path, int chain_offset, int n_chains, int n_steps, int64_t rec_bytes,
                         const void* records, int n_threads);

/* Diagnostics (not on the product path): measured ceiling of the access pattern the GF stacking uses.  Gathers
 * pseudo-random rows of row_bytes (multiple of 16, <= 16384) from a zero-filled working set of ws_bytes with
 * mode 0 = one warp-wide 16-byte-per-lane load per row (gf_stack_chunk_kernel's pattern), mode 1 = one bulk
 * asynchronous copy (TMA) per row and per mbarrier into shared memory which the warp then reads, mode 2 = those bulk
 * copies alone, mode 3 / 4 = the same with BATCHED copies (16 rows per mbarrier, 16 lanes issuing, 3 batches in flight
 * per warp; 3 reads the rows back, 4 only ingests), mode 5 = rows already resident in shared memory (what a row-staged
 * kernel would read), mode 6 / 7 / 8 = rows gathered from the shared memory of a cluster of 2 / 4 / 8 CTAs (DSMEM).
 * A working set well below the 126 MB L2 measures the L2->SM path, one far above it the HBM gather rate (modes 5..8
 * ignore it).  bench.py reports the stack kernel against the mode-0 number.                                 */
int beatgpu_probe_gather(beatgpu_ctx* ctx, int mode, int64_t ws_bytes, int row_bytes, int rows_per_warp,
                         int n_launch, float* ms_per_launch, double* bytes_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* BEATGPU_H */
