// geom.cuh -- geometry-mode seismic forward model (BASELINE config 2: double-couple point source), batched over chains.
//
// Replaces, for B chains at once, what the reference does per chain inside SeisSynthesizer.perform
// (beat/pytensorf.py:241-302) -> heart.seis_synthetics (beat/heart.py:3564-3762):
//   source update (utility.update_source, beat/utility.py:773-797; km -> m, :651-675)
//   engine.process(sources, targets)           beat/heart.py:3657  -- pyrocko's weighted, delayed sum of GF-store traces
//   post_process_trace: filter, taper, chop    beat/heart.py:3466-3525 (Filter.apply :377-392)
// followed by residual = data - synth and multivariate_normal_chol (beat/models/seismic.py:819-827).
// pyrocko (>= 2023.10.11) is an un-vendored dependency of the reference: the synthesis below restates its published
// algorithm (gf.seismosizer / gf.meta 'elastic10' sum parameters / gf.store.sum / orthodrome / moment_tensor /
// HalfSinusoidSTF.discretize_t / trace.CosTaper); oracle/geom_oracle.py is the CPU statement it is tested against.
//
// Three kernels:
//   geom_plan_kernel         thread per (chain, receiver): moment tensor, source->receiver distance / azimuths on the
//                            ellipsoid, GF-store nodes + interpolation weights, component weights; thread r == 0 also
//                            discretises the chain's source time function onto the store's time grid.
//   gf_delay_sum_kernel      CTA per (receiver, chain) -- THE BYTE MOVER.  The <= 4 nodes x 10 components GF traces the
//                            plan selected are windowed and staged in shared memory by the TMA engine
//                            (cp.async.bulk; two pipeline stages of 3 rows, one mbarrier + one CTA barrier per batch);
//                            256 threads accumulate the north / east / down seismograms in f32 registers (the store's
//                            dtype, as pyrocko does) with branch-free loops, then per target channel: sensor
//                            projection, convolution with the STF amplitudes on aligned LDS.128 windows, trace mean,
//                            raw trace to HBM as float4 in a chain-interleaved layout.  The 3 channels of a station
//                            share one pass over the rows.  Receiver-major grid: concurrent CTAs read the same few
//                            store nodes out of L2.
//   trace_filter_misfit_kernel  thread per (target, chain): streams its raw trace (coalesced float4 across chains)
//                            through the IIR cascade in f64 (scipy.signal.lfilter's direct form II transposed), taper,
//                            chop, residual against the data and -- for diagonal / narrow-band weights -- the
//                            covariance-weighted misfit and logpt, without the synthetic ever being written.
// Several sources per chain are stacked by running plan + delay-and-sum once per source (the second pass adds to the
// raw traces); station corrections move the window, the chop and the taper with the chain's time shift.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "tma.cuh"

namespace beatgpu {

constexpr int kGeomNComp = 10;                 // GF component scheme 'elastic10'
constexpr int kGeomMaxNodes = 4;               // multilinear vicinity in (source depth, distance)
constexpr int kGeomMaxRows = kGeomMaxNodes * kGeomNComp;
constexpr int kGeomMaxStf = 64;                // STF points on the time grid (duration <= 63 * deltat)
constexpr int kGeomThreads = 256;
constexpr int kGeomMaxHalfRows = 4;            // rows per pipeline half: template parameter HALF (3 or 4)
constexpr int kGeomMaxSec = 3;                 // IIR sections in cascade
constexpr int kGeomMaxOrder = 8;

struct ChainVar { const double* p; long stride; };     // value of chain c = p[c * stride]

struct GeomStoreDev {
    const float* traces;                       // [nz, nx, 10, ld]
    const int* itmin;                          // [nz, nx, 10] first sample index of a record (relative to source origin time)
    const int* nsamp;                          // [nz, nx, 10]
    int nz, nx;
    long ld;                                   // row stride in floats (multiple of 4)
    double z0, dz, x0, dx, deltat;
};

struct __align__(16) RcvPlan {
    int node[kGeomMaxNodes];                   // iz * nx + ix, or -1
    float nw[kGeomMaxNodes];                   // interpolation weight
    float wn[6], we[6], wd[4];                 // 'elastic10' weights of the north / east / down seismograms (moment included)
    int itmin, nraw;                           // window the engine computes (moves with a station correction)
    int pad0, pad1;
};

struct __align__(16) ChainPlan {
    int id0;                                   // delay of the first STF point in samples
    int n_stf;
    int pad0, pad1;
    float amp[kGeomMaxStf];
};

struct GeomPlanArgs {
    int B, nr;
    ChainVar east, north, depth, strike, dip, rake, magnitude, time, duration;     // km, km, km, deg, deg, deg, Mw, s, s
    double ev_lat, ev_lon, stf_anchor;
    int stf_type;                                                                   // BEATGPU_STF_*
    ChainVar peak_ratio;                                                            // TriangularSTF.peak_ratio
    const double* rcv_lat; const double* rcv_lon;                                   // [nr]
    GeomStoreDev store;
    int interpolation;
    // window of every receiver: fixed (rcv_itmin / rcv_nraw from the host) or, with station corrections, recomputed per
    // chain from arrival + time_shift (SeisSynthesizer.perform, pytensorf.py:248-252; update_target_times, heart.py:474-477)
    const int* rcv_itmin; const int* rcv_nraw;                                      // [nr]
    const double* rcv_arrival; const int* rcv_station;                              // [nr]; rcv_station nullptr = no corrections
    ChainVar tshift;                                                                // time_shifts [n_time_shifts]
    double ta, tb, td;                                                              // taper a, b, d relative to the arrival
    int nraw_cap;                                                                   // largest window the scratch holds
    RcvPlan* rplan;                            // [B, nr]
    ChainPlan* cplan;                          // [B]
    unsigned char* chain_bad;                  // [B]
    unsigned long long* violations;
};

__device__ __forceinline__ double clamp1(double x) { return fmin(1.0, fmax(-1.0, x)); }

// pyrocko orthodrome.ne_to_latlon (spherical, arcsin form)
__device__ inline void ne_to_latlon_dev(double lat0, double lon0, double north_m, double east_m, double& lat, double& lon)
{
    const double d2r = CUDART_PI / 180.0, r2d = 180.0 / CUDART_PI;
    const double a = sqrt(north_m * north_m + east_m * east_m) / (6371.0 * 1000.0);
    double gamma = atan2(east_m, north_m);
    const double b = CUDART_PI / 2.0 - lat0 * d2r;
    const double alphasign = gamma < 0.0 ? -1.0 : 1.0;
    gamma = fabs(gamma);
    const double c = acos(clamp1(cos(a) * cos(b) + sin(a) * sin(b) * cos(gamma)));
    const double sc = sin(c);
    double alpha = asin(clamp1(sc != 0.0 ? sin(a) * sin(gamma) / sc : 0.0));
    if (cos(a) - cos(b) * cos(c) < 0.0) alpha = alpha > 0.0 ? CUDART_PI - alpha : -CUDART_PI - alpha;
    lat = r2d * (CUDART_PI / 2.0 - c);
    lon = lon0 + r2d * alpha * alphasign;
    lon = lon + 180.0;
    lon = lon - 360.0 * floor(lon / 360.0) - 180.0;                       // wrap to [-180, 180)
}

// pyrocko orthodrome.azimuth
__device__ inline double azimuth_dev(double alat, double alon, double blat, double blon)
{
    const double d2r = CUDART_PI / 180.0, r2d = 180.0 / CUDART_PI;
    const double cd = fmin(1.0, sin(alat * d2r) * sin(blat * d2r) + cos(alat * d2r) * cos(blat * d2r) * cos(d2r * (blon - alon)));
    return r2d * atan2(cos(alat * d2r) * cos(blat * d2r) * sin(d2r * (blon - alon)), sin(d2r * blat) - sin(d2r * alat) * cd);
}

// pyrocko orthodrome.distance_accurate50m
__device__ inline double distance_accurate50m_dev(double alat, double alon, double blat, double blon)
{
    const double d2r = CUDART_PI / 180.0;
    const double f = (alat + blat) * d2r / 2.0, g = (alat - blat) * d2r / 2.0, h = (alon - blon) * d2r / 2.0;
    const double sg = sin(g), cg = cos(g), sh = sin(h), ch = cos(h), sf = sin(f), cf = cos(f);
    const double s = sg * sg * ch * ch + cf * cf * sh * sh;
    const double c = cg * cg * ch * ch + sf * sf * sh * sh;
    const double w = atan(sqrt(s / c));
    if (w == 0.0) return 0.0;
    const double r = sqrt(s * c) / w;
    const double d = 2.0 * w * (6378.14 * 1000.0);
    const double h1 = (3.0 * r - 1.0) / (2.0 * c), h2 = (3.0 * r + 1.0) / (2.0 * s);
    const double ob = 1.0 / 298.257223563;
    return d * (1.0 + ob * h1 * sf * sf * cg * cg - ob * h2 * cf * cf * sg * sg);
}

__global__ void __launch_bounds__(128) geom_plan_kernel(GeomPlanArgs a)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * a.nr) return;
    const int c = (int)(idx / a.nr), r = (int)(idx % a.nr);
    const double d2r = CUDART_PI / 180.0, r2d = 180.0 / CUDART_PI;
    const double east = a.east.p[(long)c * a.east.stride] * 1000.0;        // adjust_point_units: km -> m
    const double north = a.north.p[(long)c * a.north.stride] * 1000.0;
    const double depth = a.depth.p[(long)c * a.depth.stride] * 1000.0;
    const double strike = a.strike.p[(long)c * a.strike.stride], dip = a.dip.p[(long)c * a.dip.stride],
                 rake = a.rake.p[(long)c * a.rake.stride];
    const double moment = pow(10.0, 1.5 * (a.magnitude.p[(long)c * a.magnitude.stride] + 10.7)) * 1.0e-7;
    bool bad = !(isfinite(east) && isfinite(north) && isfinite(depth) && isfinite(strike) && isfinite(dip) && isfinite(rake) &&
                 isfinite(moment));

    // ---- chain part (one thread per chain): source time function on the store's time grid
    if (r == 0) {
        ChainPlan cp;
        const double tref = a.time.p[(long)c * a.time.stride], dur = a.duration.p[(long)c * a.duration.stride];
        const double dt = a.store.deltat;
        // [pyrocko] {HalfSinusoid,Boxcar,Triangular}STF.discretize_t: support [tmin_stf, tmax_stf], grid points
        // round(t/dt)*dt, amplitudes = integrals of the shape over the sampling intervals centred on the grid, normalised
        double tmin_stf, tmax_stf, t_peak = 0.0;
        if (a.stf_type == BEATGPU_STF_TRIANGULAR) {
            const double ra = a.peak_ratio.p[(long)c * a.peak_ratio.stride], rb = 1.0 - ra;
            const double ca = ra + (rb * rb / 3.0 - ra * ra / 3.0) / (ra + rb), cb = 1.0 - ca;        // centroid_ratio
            if (a.stf_anchor <= 0.0) { tmin_stf = tref - ca * dur * (a.stf_anchor + 1.0); tmax_stf = tmin_stf + dur; }
            else                     { tmax_stf = tref + cb * dur * (1.0 - a.stf_anchor); tmin_stf = tmax_stf - dur; }
            t_peak = tmin_stf + dur * ra;
            if (!(ra >= 0.0 && ra <= 1.0)) bad = true;
        } else {
            tmin_stf = tref - dur * (a.stf_anchor + 1.0) * 0.5; tmax_stf = tref + dur * (1.0 - a.stf_anchor) * 0.5;
        }
        const double g0 = rint(tmin_stf / dt), g1 = rint(tmax_stf / dt);
        const double tmin = g0 * dt, tmax = g1 * dt;
        const double nf = rint((tmax - tmin) / dt) + 1.0;
        int n = 1;
        // the boxcar's centroid shift may add one point
        if (!(isfinite(tref) && isfinite(dur)) || !(nf >= 1.0) || nf > (double)(kGeomMaxStf - (a.stf_type == BEATGPU_STF_BOXCAR ? 1 : 0))) bad = true;
        else n = (int)nf;
        int id0 = bad ? 0 : (int)g0;
        double w[kGeomMaxStf + 1];
        w[0] = 1.0;
        if (n > 1) {
            const double start = tmin - 0.5 * dt, stop = tmax + 0.5 * dt, step = (stop - start) / n;
            double sum = 0.0, prev = 0.0;
            for (int k = 0; k <= n; ++k) {
                const double e = (k == n) ? stop : start + k * step;                 // numpy.linspace
                const double te = fmax(tmin_stf, fmin(tmax_stf, e));                  // the shapes vanish outside the support
                double F;                                                            // antiderivative of the shape at te
                if (a.stf_type == BEATGPU_STF_BOXCAR) F = te - tmin_stf;
                else if (a.stf_type == BEATGPU_STF_TRIANGULAR) {
                    if (te <= t_peak) F = (t_peak > tmin_stf) ? (te - tmin_stf) * (te - tmin_stf) / (2.0 * (t_peak - tmin_stf)) : 0.0;
                    else F = 0.5 * (t_peak - tmin_stf) + (te - t_peak) - (te - t_peak) * (te - t_peak) / (2.0 * (tmax_stf - t_peak));
                } else F = -cos((te - tmin_stf) * (CUDART_PI / dur));
                if (k > 0) { w[k - 1] = F - prev; sum += w[k - 1]; }
                prev = F;
            }
            for (int k = 0; k < n; ++k) w[k] /= sum;
        }
        if (a.stf_type == BEATGPU_STF_BOXCAR && !bad) {
            // tshift = sum(amplitudes * times) - centroid_time; sshift(times, amplitudes, -tshift, deltat)
            const double tstep = n > 1 ? (tmax - tmin) / (n - 1) : 0.0;
            double cen = 0.0;
            for (int k = 0; k < n; ++k) cen += w[k] * ((k == n - 1 && n > 1) ? tmax : tmin + k * tstep);
            const double sft = -(cen - (tref - 0.5 * dur * a.stf_anchor));
            const double t0 = floor(sft / dt) * dt, t1 = ceil(sft / dt) * dt;
            if (t0 != t1) {
                const double wl = (t1 - sft) / dt, wr = (sft - t0) / dt;
                w[n] = 0.0;
                for (int k = n; k >= 1; --k) w[k] = wl * w[k] + wr * w[k - 1];
                w[0] = wl * w[0];
                n += 1;
                id0 = (int)rint((tmin + t0) / dt);
            }
        }
        cp.id0 = bad ? 0 : id0;
        cp.n_stf = n;
        cp.pad0 = cp.pad1 = 0;
        for (int k = 0; k < kGeomMaxStf; ++k) cp.amp[k] = (k < n) ? (float)w[k] : 0.f;
        a.cplan[c] = cp;
    }

    // ---- double-couple moment tensor, north-east-down (Aki & Richards 4.91 == pyrocko MomentTensor(strike, dip, rake))
    const double phi = strike * d2r, del = dip * d2r, lam = rake * d2r;
    const double sd = sin(del), cd = cos(del), s2d = sin(2.0 * del), c2d = cos(2.0 * del), sl = sin(lam), cl = cos(lam);
    const double sp = sin(phi), cp_ = cos(phi), s2p = sin(2.0 * phi), c2p = cos(2.0 * phi);
    const double mnn = -(sd * cl * s2p + s2d * sl * sp * sp) * moment;
    const double mee = (sd * cl * s2p - s2d * sl * cp_ * cp_) * moment;
    const double mdd = (s2d * sl) * moment;
    const double mne = (sd * cl * c2p + 0.5 * s2d * sl * s2p) * moment;
    const double mnd = -(cd * cl * cp_ + c2d * sl * sp) * moment;
    const double med = -(cd * cl * sp - c2d * sl * cp_) * moment;

    // ---- source -> receiver geometry (gf.meta.DiscretizedSource.distances_to / azibazis_to)
    const double rlat = a.rcv_lat[r], rlon = a.rcv_lon[r];
    double dist, azi, bazi;
    if (a.ev_lat == rlat && a.ev_lon == rlon) {
        dist = sqrt(north * north + east * east);
        azi = r2d * atan2(0.0 - east, 0.0 - north);
        bazi = azi + 180.0;
    } else {
        double slat, slon;
        ne_to_latlon_dev(a.ev_lat, a.ev_lon, north, east, slat, slon);
        dist = distance_accurate50m_dev(slat, slon, rlat, rlon);
        azi = azimuth_dev(slat, slon, rlat, rlon);
        bazi = azimuth_dev(rlat, rlon, slat, slon);
    }

    // ---- 'elastic10' sum weights (gf.meta.DiscretizedMTSource.make_weights)
    RcvPlan rp;
    {
        const double sa = sin(azi * d2r), ca = cos(azi * d2r), sa2 = sin(2.0 * azi * d2r), ca2 = cos(2.0 * azi * d2r);
        const double sb = sin(bazi * d2r - CUDART_PI), cb = cos(bazi * d2r - CUDART_PI);
        const double f0 = mnn * ca * ca + mee * sa * sa + mne * sa2;
        const double f1 = mnd * ca + med * sa;
        const double f2 = mdd;
        const double f3 = 0.5 * (mee - mnn) * sa2 + mne * ca2;
        const double f4 = med * ca - mnd * sa;
        const double f5 = mnn * sa * sa + mee * ca * ca - mne * sa2;
        rp.wn[0] = (float)(cb * f0); rp.wn[1] = (float)(cb * f1); rp.wn[2] = (float)(cb * f2); rp.wn[3] = (float)(cb * f5);
        rp.wn[4] = (float)(-sb * f3); rp.wn[5] = (float)(-sb * f4);
        rp.we[0] = (float)(sb * f0); rp.we[1] = (float)(sb * f1); rp.we[2] = (float)(sb * f2); rp.we[3] = (float)(sb * f5);
        rp.we[4] = (float)(cb * f3); rp.we[5] = (float)(cb * f4);
        rp.wd[0] = (float)f0; rp.wd[1] = (float)f1; rp.wd[2] = (float)f2; rp.wd[3] = (float)f5;
    }

    // ---- store nodes (gf.meta.ConfigTypeA index function / vicinity)
    const double xa = (depth - a.store.z0) / a.store.dz, xb = (dist - a.store.x0) / a.store.dx;
    const double eps = 1e-9;
    bool oob = !(xa >= -eps && xa <= a.store.nz - 1 + eps && xb >= -eps && xb <= a.store.nx - 1 + eps);
    for (int i = 0; i < kGeomMaxNodes; ++i) { rp.node[i] = -1; rp.nw[i] = 0.f; }
    if (!oob && !bad) {
        const double ya = fmin(fmax(xa, 0.0), a.store.nz - 1.0), yb = fmin(fmax(xb, 0.0), a.store.nx - 1.0);
        if (a.interpolation == 0) {
            rp.node[0] = (int)rint(ya) * a.store.nx + (int)rint(yb);
            rp.nw[0] = 1.f;
        } else {
            const double fa = floor(ya), fb = floor(yb);
            const double wa[2] = {1.0 - (ya - fa), ya - fa}, wb[2] = {1.0 - (yb - fb), yb - fb};
            const int ia[2] = {(int)fa, (int)ceil(ya)}, ib[2] = {(int)fb, (int)ceil(yb)};
            int n = 0;
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    if (wa[i] != 0.0 && wb[j] != 0.0) { rp.node[n] = ia[i] * a.store.nx + ib[j]; rp.nw[n] = (float)(wa[i] * wb[j]); ++n; }
        }
    }
    // ---- window
    rp.pad0 = rp.pad1 = 0;
    if (a.rcv_station) {
        const double at = a.rcv_arrival[r] + a.tshift.p[(long)c * a.tshift.stride + a.rcv_station[r]];
        const double wa = at + a.ta, wb = at + a.tb, wd_ = at + a.td;
        const double tol = 2.0 * (wb - wa);
        const double lo = floor((wa - tol) / a.store.deltat), hi = ceil((wd_ + tol) / a.store.deltat);
        if (!isfinite(at) || !(fabs(lo) < 2.0e9) || !(hi - lo + 1.0 <= (double)a.nraw_cap)) { bad = true; rp.itmin = 0; rp.nraw = 1; }
        else { rp.itmin = (int)lo; rp.nraw = (int)(hi - lo) + 1; }
    } else {
        rp.itmin = a.rcv_itmin[r];
        rp.nraw = a.rcv_nraw[r];
    }
    if (oob && !bad) atomicAdd(a.violations, 1ULL);
    if (oob || bad) a.chain_bad[c] = 1;
    a.rplan[idx] = rp;
}

// ------------------------------------------------------------------------------------------------------------------
// delay-and-sum
// ------------------------------------------------------------------------------------------------------------------
struct GeomSumArgs {
    int B, nr, nt;
    GeomStoreDev store;
    const RcvPlan* rplan;                      // [B, nr]
    const ChainPlan* cplan;                    // [B]
    unsigned char* chain_bad;                  // [B]; a CTA that loses a bulk copy marks its chain so the filter pass emits NaN
    const int* rcv_first;                      // [nr + 1] CSR into tgt_of
    const int* tgt_of;                         // target index of each channel of a receiver
    const float* tgt_f;                        // [nt, 3] sensor factors (north, east, down) = (ca*cd, sa*cd, sd) of azimuth/dip
    int slot_floats;                           // floats per shared-memory row slot (multiple of 4)
    int half_floats;                           // floats per pipeline half: HALF slots + a pad that absorbs the over-reads
                                               // of the branch-free loops (they must never touch a half that is in flight)
    int n4;                                    // float4 groups per raw trace
    float* rawT;                               // [nt, n4, B, 4] raw traces, chain-interleaved
    double* mean;                              // [B, nt] mean of each raw trace
    int accumulate;                            // second and further sources: add to what the previous source left (heart.py:3719-3724)
    unsigned int* err;                         // [2]: bulk-copy time-outs of this call (err[0], cleared per call) and since the
                                               // last beatgpu_geom_timeouts query (err[1]); both must stay 0
};

struct __align__(16) RowInfo {
    long src;                                  // float offset into store.traces of the first staged sample
    int bytes;                                 // staged bytes (multiple of 16)
    int rel;                                   // record sample of comb index 0:  j = c + rel
    int ja;                                    // first staged record sample
    int jhi;                                   // last staged record sample that the window can address
    float w0, w1;                              // kind 0: weights into north / east; kind 1: w0 into down
    int kind;
    int clamp;                                 // window leaves the record: indices must be clamped (repeat end values)
};

// HALF rows per pipeline half: 2*HALF slots of ~9 KB at config-2 size -> HALF = 4: 72 KB, three CTAs per SM;
// HALF = 3: 54 KB, four CTAs per SM (BEATGPU_GEOM_HALF selects; measurements in profiles/README.md).
// NST pipeline stages of HALF rows each (NST = 2: double buffering; NST = 3 keeps two batches in flight).
template <int ACC, int HALF, int NST>
__global__ void __launch_bounds__(kGeomThreads) gf_delay_sum_kernel(GeomSumArgs a)
{
    extern __shared__ __align__(128) float slots[];                               // [NST][half_floats >= HALF * slot_floats]
    __shared__ RowInfo rows[kGeomMaxRows];
    // candidate table (before compaction) lives at the start of the ring (>= 2304 B for any window): the bulk copies
    // that overwrite it are issued by thread 0 behind a fence.proxy.async after the compaction has been barriered
    RowInfo* cand = reinterpret_cast<RowInfo*>(slots);
    __shared__ unsigned char s_valid[kGeomMaxRows];
    __shared__ __align__(8) uint64_t bar[NST];
    __shared__ __align__(16) float s_amp[kGeomMaxStf + 4];                        // zero padded to a multiple of 4 taps
    __shared__ double red[kGeomThreads / 32];
    __shared__ int s_nrows;

    const int tid = threadIdx.x;
    const int r = blockIdx.x / a.B, c = blockIdx.x % a.B;
    if (a.chain_bad[c]) return;                                                   // uniform per CTA
    if (*(volatile unsigned int*)a.err) {                                         // after a time-out the grid drains: this chain gets no
        if (tid == 0) a.chain_bad[c] = 1;                                         // raw trace, so it must not be filtered from stale scratch
        return;
    }
    const ChainPlan& cp = a.cplan[c];
    const RcvPlan& rp = a.rplan[(long)c * a.nr + r];
    const int n_stf = cp.n_stf;
    const int n_raw = rp.nraw;                                                    // window from the plan (absolute first sample rp.itmin)
    const int ncomb = n_raw + n_stf - 1;
    const int m0 = rp.itmin - cp.id0 - (n_stf - 1);                         // sample (rel. to source origin) of comb[0]

    if (tid >= 64 && tid < 64 + kGeomMaxStf + 4) s_amp[tid - 64] = (tid - 64 < kGeomMaxStf) ? cp.amp[tid - 64] : 0.f;
    // row table: one thread per (node, component) candidate, so the record headers are fetched in parallel
    if (tid < kGeomMaxRows) {
        const int i = tid / kGeomNComp, g = tid % kGeomNComp;
        const int node = rp.node[i];
        bool valid = node >= 0;
        if (valid) {
            // component -> (kind, slot): north/east use g = 0,1,2,8,3,4; down uses g = 5,6,7,9
            const int kind = (g >= 5 && g != 8) ? 1 : 0;
            const int k = kind == 0 ? (g < 3 ? g : (g == 8 ? 3 : g + 1)) : (g == 9 ? 3 : g - 5);
            RowInfo ri;
            ri.kind = kind;
            ri.w0 = rp.nw[i] * (kind == 0 ? rp.wn[k] : rp.wd[k]);
            ri.w1 = kind == 0 ? rp.nw[i] * rp.we[k] : 0.f;
            valid = !(ri.w0 == 0.f && ri.w1 == 0.f);
            const long rec = (long)node * kGeomNComp + g;
            const int it_rec = a.store.itmin[rec];
            const int nrec = a.store.nsamp[rec];
            ri.rel = m0 - it_rec;
            const int jlo = min(max(ri.rel, 0), nrec - 1), jhi = min(max(ri.rel + ncomb - 1, 0), nrec - 1);
            ri.clamp = (ri.rel < 0 || ri.rel + ncomb - 1 > nrec - 1) ? 1 : 0;
            ri.jhi = jhi;
            ri.ja = jlo & ~3;
            const int jb = (int)min((long)a.store.ld, (long)((jhi + 4) & ~3));
            ri.bytes = (jb - ri.ja) * 4;
            ri.src = rec * a.store.ld + ri.ja;
            cand[tid] = ri;
        }
        s_valid[tid] = valid ? 1 : 0;
    }
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid < kGeomMaxRows) {                                                     // compact, keeping (node, component) order
        int pos = 0;
        for (int j = 0; j < tid; ++j) pos += s_valid[j];
        if (s_valid[tid]) rows[pos] = cand[tid];
        if (tid == kGeomMaxRows - 1) s_nrows = pos + s_valid[tid];
    }
    __syncthreads();
    const int nrows = s_nrows;
    const int nbatch = (nrows + HALF - 1) / HALF;

    // Accumulator u of a thread belongs to comb index tid + 256 u.  The loops below are branch-free: indices past the
    // window (>= ncomb) read whatever lies behind it in shared memory (the ring is padded so that this stays in
    // bounds) and those accumulators are never read.
    float acc_n[ACC], acc_e[ACC], acc_d[ACC];
#pragma unroll
    for (int u = 0; u < ACC; ++u) acc_n[u] = acc_e[u] = acc_d[u] = 0.f;

    auto issue = [&](int nb) {                                                    // thread 0 only
        const int h = nb % NST, r0 = nb * HALF, r1 = min(nrows, r0 + HALF);
        uint32_t total = 0;
        for (int i = r0; i < r1; ++i) total += (uint32_t)rows[i].bytes;
        mbar_arrive_expect_tx(&bar[h], total);
        for (int i = r0; i < r1; ++i)
            tma_load_1d(slots + (long)h * a.half_floats + (long)(i - r0) * a.slot_floats, a.store.traces + rows[i].src,
                        (uint32_t)rows[i].bytes, &bar[h]);
    };

    if (tid == 0) {
        fence_proxy_async();
        for (int nb = 0; nb < min(nbatch, NST - 1); ++nb) issue(nb);
    }
    bool ok = true;
    for (int nb = 0; nb < nbatch; ++nb) {
        // stage (nb + NST - 1) % NST == (nb - 1) % NST was released by the barrier that ended iteration nb - 1
        if (tid == 0 && nb + NST - 1 < nbatch) { fence_proxy_async(); issue(nb + NST - 1); }
        const int h = nb % NST;
        // every thread waits on the barrier itself (measured: one polling warp with nanosleep back-off while the others
        // sleep in bar.sync is 4 % slower -- the back-off adds latency that the three co-resident CTAs do not hide).
        // A wait that times out does not leave the loop (all threads must keep meeting at the barrier below): the CTA
        // finishes on whatever the slot holds and the failure is reported through a.err.
        if (!mbar_wait_bounded(&bar[h], (uint32_t)(nb / NST) & 1u, 1u << 20)) ok = false;
        const int r0 = nb * HALF, r1 = min(nrows, r0 + HALF);
        for (int i = r0; i < r1; ++i) {
            const RowInfo ri = rows[i];
            const float* s = slots + h * a.half_floats + (i - r0) * a.slot_floats;
            if (!ri.clamp) {                                                      // window inside the record: s[c + off]
                const float* so = s + (ri.rel - ri.ja) + tid;                     // one LDS [R + imm] per element
                if (ri.kind == 0) {
#pragma unroll
                    for (int u = 0; u < ACC; ++u) {
                        const float x = so[u * kGeomThreads];
                        acc_n[u] = fmaf(ri.w0, x, acc_n[u]); acc_e[u] = fmaf(ri.w1, x, acc_e[u]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < ACC; ++u) acc_d[u] = fmaf(ri.w0, so[u * kGeomThreads], acc_d[u]);
                }
            } else {                                                              // repeat the record's first / last value
                const float* so = s - ri.ja;
                const int hi = ri.jhi, j0 = tid + ri.rel;                         // indices past the window clamp onto its last staged sample
                if (ri.kind == 0) {
#pragma unroll
                    for (int u = 0; u < ACC; ++u) {
                        const float x = so[min(max(j0 + u * kGeomThreads, 0), hi)];
                        acc_n[u] = fmaf(ri.w0, x, acc_n[u]); acc_e[u] = fmaf(ri.w1, x, acc_e[u]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < ACC; ++u) acc_d[u] = fmaf(ri.w0, so[min(max(j0 + u * kGeomThreads, 0), hi)], acc_d[u]);
                }
            }
        }
        __syncthreads();
    }
    if (__syncthreads_or(!ok)) {
        if (tid == 0) { atomicAdd(a.err, 1u); atomicAdd(a.err + 1, 1u); a.chain_bad[c] = 1; }   // chain rejected (NaN), never stale data
        return;
    }

    // ---- per target channel of this receiver: sensor projection, STF convolution, mean, store.
    // The combined trace is laid out in slot 0 so that comb[n_stf - 1] is 16-byte aligned: a thread then produces four
    // consecutive output samples from aligned LDS.128 windows (raw[i] = sum_k amp[k] comb[i + n_stf-1 - k]; per 4 taps
    // 3 vector loads + 16 FMA) and writes them as one float4 into the chain-interleaved raw layout.  Zero guards in front
    // of and behind the trace stand in for the taps / samples that padding to multiples of four adds.
    const int padc = (4 - ((n_stf - 1) & 3)) & 3;
    float* comb = slots + 8 + padc;
    if (tid < 8 + padc) slots[tid] = 0.f;
    const int n_tap4 = (n_stf + 3) >> 2;
    const int nq = (n_raw + 3) >> 2;
    for (int q = a.rcv_first[r]; q < a.rcv_first[r + 1]; ++q) {
        const int t = a.tgt_of[q];
        const float fn = a.tgt_f[3 * t], fe = a.tgt_f[3 * t + 1], fd = a.tgt_f[3 * t + 2];
#pragma unroll
        for (int u = 0; u < ACC; ++u) {
            const int cc = tid + u * kGeomThreads;
            if (cc < ncomb) comb[cc] = fn * acc_n[u] + fe * acc_e[u] + fd * acc_d[u];
        }
        if (tid < 4) comb[ncomb + tid] = 0.f;
        __syncthreads();
        double lsum = 0.0;
        float4* dst = (float4*)a.rawT + ((long)t * a.n4 + tid) * a.B + c;                     // quad i4 = tid + 256 m
        const long dstep = (long)kGeomThreads * a.B;
        const float* wb = comb + (n_stf - 1) + 4 * tid;                                       // 16-byte aligned
        for (int i4 = tid; i4 < nq; i4 += kGeomThreads, dst += dstep, wb += 4 * kGeomThreads) {
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            for (int c4 = 0; c4 < n_tap4; ++c4) {
                const float4 am = *(const float4*)(s_amp + 4 * c4);
                const float4 lo = *(const float4*)(wb - 4 * c4 - 4);                          // comb offsets -4 .. -1
                const float4 hi = *(const float4*)(wb - 4 * c4);                              //               0 ..  3
                v0 = fmaf(am.x, hi.x, v0); v0 = fmaf(am.y, lo.w, v0); v0 = fmaf(am.z, lo.z, v0); v0 = fmaf(am.w, lo.y, v0);
                v1 = fmaf(am.x, hi.y, v1); v1 = fmaf(am.y, hi.x, v1); v1 = fmaf(am.z, lo.w, v1); v1 = fmaf(am.w, lo.z, v1);
                v2 = fmaf(am.x, hi.z, v2); v2 = fmaf(am.y, hi.y, v2); v2 = fmaf(am.z, hi.x, v2); v2 = fmaf(am.w, lo.w, v2);
                v3 = fmaf(am.x, hi.w, v3); v3 = fmaf(am.y, hi.z, v3); v3 = fmaf(am.z, hi.y, v3); v3 = fmaf(am.w, hi.x, v3);
            }
            if (a.accumulate) { const float4 p = *dst; v0 += p.x; v1 += p.y; v2 += p.z; v3 += p.w; }
            *dst = make_float4(v0, v1, v2, v3);
            const int i = 4 * i4;                                                             // the last quad may be partial
            lsum += (double)v0;
            if (i + 1 < n_raw) lsum += (double)v1;
            if (i + 2 < n_raw) lsum += (double)v2;
            if (i + 3 < n_raw) lsum += (double)v3;
        }
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = lsum;
        __syncthreads();                                                          // also: comb fully consumed
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < kGeomThreads / 32; ++w) sum += red[w];
            a.mean[(long)c * a.nt + t] = sum / (double)n_raw;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// filter + taper + chop + residual + misfit
// ------------------------------------------------------------------------------------------------------------------
struct GeomFilterArgs {
    int B, nt, ns;
    int n4;
    const float* rawT;                         // [nt, n4, B, 4]
    const double* mean;                        // [B, nt]
    const unsigned char* chain_bad;
    const int* tgt_nraw;                       // [nt]
    const int* tgt_ibeg;                       // [nt] first chopped sample within the raw window
    const double* taper;                       // [nt, ns] factors on the chopped samples, or nullptr
    // station corrections: window, chop start and taper follow arrival + time_shift of the chain
    const RcvPlan* rplan; const int* tgt_rcv; int nr;                               // rplan nullptr = fixed windows
    const double* tgt_arrival; const int* tgt_station; ChainVar tshift;
    double ta, tb, tc, td, dt; int chop_lo, chop_hi;
    // IIR cascade, coefficients normalised by a[0]; b[s][0..ORD], a[s][1..ORD] (a[s][0] unused), zero padded
    double fb[kGeomMaxSec][kGeomMaxOrder + 1];
    double fa[kGeomMaxSec][kGeomMaxOrder + 1];
    int demean;                                // remove the trace mean before the first section
    // misfit (fused modes)
    const double* data;                        // [nt, ns]
    const double* W; int bw;                   // band weights [nt, bw+1, ns] (bw = 0: diagonal [nt, ns])
    const double* slog_pdet; const int* nsamp; const int* hyper_idx;
    ChainVar hyp;                              // hypers of chain c at hyp.p[c*hyp.stride + hyper_idx[t]]
    double* logpts; long logpts_sc; int out_ofs;
    // output mode
    double* out; int out_resid;                // [B, nt, ns]: synthetics (out_resid = 0) or residuals (1)
};

// MODE 0: fused misfit with band width <= 1; MODE 1: fused misfit with band width <= 8; MODE 2: write synthetics / residuals
// The kernel is FP64-pipe work (about 24 DFMA-class instructions per sample and thread) that only runs at rate when the
// loads are off the critical path: raw samples are prefetched one float4 ahead and the data / weight / taper operands
// of a block of four samples are fetched together at the top of the block (measured: 1.9 ms -> see profiles/README.md
// with per-sample loads at config-2 size, long-scoreboard stalls 56 % of all stall cycles).
constexpr int kFilterThreads = 128;
// CAP: register cap (80 -> 6 CTAs per SM instead of 5 at the uncapped 94; BEATGPU_FILTER_CAP selects, see profiles/README.md)
// DYN: station corrections -- window, chop start and taper flanks follow the chain's time shift (kept out of the common
// instantiation: the extra state costs ~25 registers)
template <int NSEC, int ORD, int MODE, int CAP, bool DYN>
__global__ void __launch_bounds__(kFilterThreads) __maxnreg__(CAP) trace_filter_misfit_kernel(GeomFilterArgs a)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.nt * a.B) return;
    const int t = (int)(idx / a.B), c = (int)(idx % a.B);
    if (a.chain_bad[c]) {
        if (MODE != 2) a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = CUDART_NAN;
        else for (int k = 0; k < a.ns; ++k) a.out[((long)c * a.nt + t) * a.ns + k] = CUDART_NAN;
        return;
    }
    constexpr int MAXBW = MODE == 0 ? 1 : 8;
    const int ns = a.ns;
    int ibeg, n_raw_t;
    // station corrections: taper flanks evaluated on the fly (apply_costaper); h* = first sample index at / after a, b, c, d
    int h_a = 0, h_b = 0, h_c = 0x7fffffff, h_d = 0x7fffffff;
    double wa = 0.0, wb = 1.0, wc = 0.0, wd = 1.0, x0 = 0.0;
    constexpr bool dyn = DYN;
    if (dyn) {
        const RcvPlan& rp = a.rplan[(long)c * a.nr + a.tgt_rcv[t]];
        const double at = a.tgt_arrival[t] + a.tshift.p[(long)c * a.tshift.stride + a.tgt_station[t]];
        wa = at + a.ta; wb = at + a.tb; wc = at + a.tc; wd = at + a.td;
        x0 = (double)rp.itmin * a.dt;
        n_raw_t = rp.nraw;
        const double lower = a.chop_lo == 0 ? wa : (a.chop_lo == 1 ? wb : wc);
        ibeg = max(0, (int)floor((lower - x0) / a.dt));                               // Trace.chop, snap = (floor, floor)
        const double nn = (double)n_raw_t;
        h_a = (int)fmax(0.0, fmin(nn, ceil((wa - x0) / a.dt)));
        h_b = (int)fmax(0.0, fmin(nn, ceil((wb - x0) / a.dt)));
        h_c = (int)fmax(0.0, fmin(nn, ceil((wc - x0) / a.dt)));
        h_d = (int)fmax(0.0, fmin(nn, ceil((wd - x0) / a.dt)));
    } else {
        ibeg = a.tgt_ibeg[t];
        n_raw_t = a.tgt_nraw[t];
    }
    const int iend = min(n_raw_t, ibeg + ns);                       // nothing after the chop window is needed
    if (dyn && ibeg + ns > n_raw_t) {
        // the shifted chop window leaves the computed trace: the reference cannot stack such traces (ValueError,
        // beat/heart.py:3700-3716); report instead of evaluating on fewer samples
        if (MODE != 2) a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = CUDART_NAN;
        else for (int k = 0; k < ns; ++k) a.out[((long)c * a.nt + t) * ns + k] = CUDART_NAN;
        return;
    }
    const double mu = a.demean ? a.mean[(long)c * a.nt + t] : 0.0;
    const float4* src = (const float4*)a.rawT + (long)t * a.n4 * a.B + c;
    const double* dat = a.data ? a.data + (long)t * ns : nullptr;
    const double* tap = a.taper ? a.taper + (long)t * ns : nullptr;
    const int bw = a.bw;
    const double* Wt = (MODE != 2) ? a.W + (long)t * (bw + 1) * ns : nullptr;

    double z[NSEC][ORD];
#pragma unroll
    for (int s = 0; s < NSEC; ++s)
#pragma unroll
        for (int j = 0; j < ORD; ++j) z[s][j] = 0.0;
    double sh[MAXBW + 1];
#pragma unroll
    for (int j = 0; j <= MAXBW; ++j) sh[j] = 0.0;
    double quad = 0.0;

    auto push_resid = [&](int k, double rk) {       // k = index of rk; emits z_{k-bw}
        if (MODE == 0) {
            if (bw == 0) { const double zz = Wt[k] * rk; quad = fma(zz, zz, quad); }
            else {
                sh[0] = sh[1]; sh[1] = rk;
                const int kk = k - 1;
                if (kk >= 0) { const double zz = fma(Wt[ns + kk], sh[1], Wt[kk] * sh[0]); quad = fma(zz, zz, quad); }
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < MAXBW; ++j) sh[j] = sh[j + 1];
            sh[MAXBW] = rk;
            const int kk = k - bw;
            if (kk >= 0) {
                double zz = 0.0;
#pragma unroll
                for (int j = 0; j <= MAXBW; ++j) {
                    const int jj = j - (MAXBW - bw);
                    if (jj >= 0) zz = fma(Wt[(long)jj * ns + kk], sh[j], zz);
                }
                quad = fma(zz, zz, quad);
            }
        }
    };

    float4 nxt = __ldcs(src);
    for (int i0 = 0; i0 < iend; i0 += 4) {
        const float4 v4 = nxt;
        if (i0 + 4 < iend) nxt = __ldcs(src + (long)((i0 >> 2) + 1) * a.B);       // prefetch the next four samples
        const float xs[4] = {v4.x, v4.y, v4.z, v4.w};
        // operands of the window samples of this block, all requested before the arithmetic starts (indices clamped into
        // the window; a value fetched for a sample outside it is never used)
        const int k0 = i0 - ibeg;
        double cd[4], ct[4], cw0[4], cw1[4];
        if (k0 > -4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kc = min(max(k0 + e, 0), ns - 1);
                cd[e] = dat ? dat[kc] : 0.0;
                ct[e] = tap ? tap[kc] : 1.0;
                if (MODE == 0) {
                    const int kk = max(kc - bw, 0);
                    cw0[e] = Wt[kk];
                    cw1[e] = bw ? Wt[ns + kk] : 0.0;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            if (i < iend) {
                double x = (double)xs[e] - mu;
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {                       // scipy.signal.lfilter: direct form II transposed
                    const double y = fma(a.fb[s][0], x, z[s][0]);
#pragma unroll
                    for (int j = 0; j < ORD - 1; ++j) z[s][j] = fma(-a.fa[s][j + 1], y, fma(a.fb[s][j + 1], x, z[s][j + 1]));
                    z[s][ORD - 1] = fma(-a.fa[s][ORD], y, a.fb[s][ORD] * x);
                    x = y;
                }
                const int k = i - ibeg;
                if (k >= 0) {
                    if (tap) x *= ct[e];
                    if (dyn && (i < h_b || i >= h_c)) {             // on a taper flank (at most the first sample for chop (b, c))
                        double f;
                        if (i < h_a || i >= h_d) f = 0.0;
                        else if (i < h_b) f = 0.5 - 0.5 * cos((a.dt * (double)i - (wa - x0)) / (wb - wa) * CUDART_PI);
                        else f = 0.5 + 0.5 * cos((a.dt * (double)i - (wc - x0)) / (wd - wc) * CUDART_PI);
                        x *= f;
                    }
                    if (MODE == 2) a.out[((long)c * a.nt + t) * ns + k] = a.out_resid ? cd[e] - x : x;
                    else if (MODE == 1) push_resid(k, cd[e] - x);
                    else {                                               // band width <= 1 with the prefetched weights
                        const double rk = cd[e] - x;
                        if (bw == 0) { const double zz = cw0[e] * rk; quad = fma(zz, zz, quad); }
                        else {
                            sh[0] = sh[1]; sh[1] = rk;
                            if (k >= 1) { const double zz = fma(cw1[e], sh[1], cw0[e] * sh[0]); quad = fma(zz, zz, quad); }
                        }
                    }
                }
            }
        }
    }
    if constexpr (MODE != 2) {
        for (int k = ns; k < ns + bw; ++k) push_resid(k, 0.0);         // flush: rows whose band is cut by the matrix edge
        const double h = a.hyp.p[(long)c * a.hyp.stride + a.hyper_idx[t]];
        const double M = (double)(short)a.nsamp[t];                   // int16 cast of the reference (distributions.py:120)
        const double norm = M * (2.0 * h + 1.8378770664093453);       // distributions.py:129-137, as misfit_kernel
        a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = (-0.5) * (a.slog_pdet[t] + norm + (1.0 / exp(h * 2.0)) * quad);
    }
}

}  // namespace beatgpu
