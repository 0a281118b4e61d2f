// pose.h -- device-resident state of one Trimmed::_align and the closed-form
// pose step that ends every iteration (K6).
//
// Replaces: the tail of Pairs::getTransform (reference icp/icp.cpp:68-99) and the
// loop bookkeeping of Trimmed::_align (icp/icp.hpp:459-485).
//
// The reference does NOT compute Horn's optimum: it builds
//     Q = [[tr S, d^T], [d, A - tr S * I]],  A = S - S^T,  d = (A12, A20, A01)
// (A, not S + S^T, in the lower-right block) and takes the eigenvector of the
// largest real eigenvalue of this non-symmetric Q.  Since A u = -d x u, the
// vector (lambda + t, d) with lambda = sqrt(t^2 + |d|^2), t = tr S, satisfies
// Q v = lambda v, and the spectrum is {+-lambda, -t +- i|d|}: the reference's
// rotation is "about d by atan2(|d|, t)".  We reproduce exactly that (parity),
// in closed form, on the device.  All arithmetic fp64, no FMA contraction
// (the library is compiled with -fmad=false).
#pragma once
#include "grid_view.h"

#define ICPB_NSUMS 17 // sum p (3), sum x (3), sum p x^T (9, row-major), sum d^2, n
#define ICPB_TRACE_COLS_DEV 28

struct IcpState {
    // transforms, 3x4 row-major [L | t]
    double C[12];    // C_global2globalnew
    double Cl2g[12]; // source C_local2global
    double M[12];    // C_local2globalnew = C * Cl2g   (setOffsetTransform)
    double Mprev[12]; // M of the previous loop body (search skipping: how far every query moved since then)
    double d_box, mse, mse_diff;
    double min_mse, min_mse_diff, overlap;
    unsigned long long iter, max_iter, pairs, n_po;
    unsigned long long n_local, n_global_offset;
    // per-iteration scratch
    unsigned long long found; // pairs found by K3 (all ranks after the collective)
    unsigned long long far;   // queries that started above level 0 (diagnostic)
    unsigned long long skipped; // queries answered by the skip test, without a search (diagnostic)
    unsigned long long deferred; // queries whose search was deferred: a pair exists, but beyond what the trim keeps (diagnostic)
    // lazy search: a query whose nearest neighbour provably is farther than tau_guard (and provably has a pair within
    // d_box) is not searched; its distance entry is a lower bound.  The trim result is exact as long as the new tau does
    // not exceed tau_guard -- otherwise the loop body is run again with force_exact (icpb_iteration_finish)
    double tau_guard, lazy_margin; // lazy_margin <= -1: lazy search off
    int force_exact, redone;
    double sums[ICPB_NSUMS];
    // trim (K4)
    unsigned long long sel_prefix, sel_krem, sel_k;
    unsigned long long cnt_eq, quota; // entries equal to tau / how many of them are kept
    double tau;
    long long p_cut;       // ties at local positions <= p_cut are kept (-1: none, 2^32: all)
    unsigned int need_tie; // quota < cnt_eq: run the tie-break passes
    unsigned int tie_prefix;
    unsigned long long tie_local; // multi-GPU: entries equal to tau on this rank
    unsigned long long tie_krem;
    // control
    int done;     // loop condition false (or early return): later launches are no-ops
    int early;    // _align returned at pairs < MIN_PAIRS (icp/icp.hpp:475-476)
    int executed; // loop bodies executed so far (== trace rows)
    int warm;     // nn_pos[] of the previous iteration is valid
    unsigned int tickets[16];
    unsigned int q_count[2]; // search queue of this loop body: fast entries (from the front), walk entries (from the back)
    unsigned int q_cursor;   // next queue entry to hand out
    unsigned int lin_epoch, lin_ready; // multi-GPU one-launch trim: the summed histogram of body #lin_epoch+1 is published
    unsigned int comm_seq;   // collectives completed on the peer-memory path (same on every rank)
    unsigned int comm_error; // a peer did not show up within the spin budget
#ifdef ICPB_PHASE_STAMPS
    unsigned long long stamps[16]; // %globaltimer at the phase boundaries of trim_moments_kernel (last launch; study builds only)
#endif
};

ICPB_HD void icpb_aff_identity(double *a)
{
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    a[0] = a[5] = a[10] = 1.0;
}
// res = a * b for 3x4 affine [L|t] (Eigen: L = La*Lb, t = La*tb + ta)
ICPB_HD void icpb_aff_mul(const double *a, const double *b, double *res)
{
    double o[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            o[r * 4 + c] = (a[r * 4 + 0] * b[0 * 4 + c] + a[r * 4 + 1] * b[1 * 4 + c]) + a[r * 4 + 2] * b[2 * 4 + c];
        o[r * 4 + 3] = ((a[r * 4 + 0] * b[3] + a[r * 4 + 1] * b[7]) + a[r * 4 + 2] * b[11]) + a[r * 4 + 3];
    }
    for (int i = 0; i < 12; ++i) res[i] = o[i];
}
// p = L*v + t  (PointcloudAdapter::next, icp/icp.hpp:130)
ICPB_HD void icpb_aff_apply(const double *a, double vx, double vy, double vz, double &px, double &py, double &pz)
{
    px = ((a[0] * vx + a[1] * vy) + a[2] * vz) + a[3];
    py = ((a[4] * vx + a[5] * vy) + a[6] * vz) + a[7];
    pz = ((a[8] * vx + a[9] * vy) + a[10] * vz) + a[11];
}

// unit quaternion (w,x,y,z) of the reference's Q for cross-covariance s (row-major p x^T)
ICPB_HD void icpb_pose_quat(const double *s, double *q)
{
    const double d0 = s[5] - s[7];
    const double d1 = s[6] - s[2];
    const double d2 = s[1] - s[3];
    const double tr = (s[0] + s[4]) + s[8];
    const double dn = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    if (dn == 0.0) {
        // symmetric cross-covariance: Q = diag(t, -t, -t, -t).  t >= 0: the identity.  t < 0: the largest eigenvalue -t is
        // triple, its eigenvectors span (x,y,z): some rotation by pi -- which one the reference's EigenSolver returns is
        // not determined; the x axis is taken here (measure-zero case, listed among the deviations in DESIGN.md)
        q[0] = tr >= 0.0 ? 1.0 : 0.0;
        q[1] = tr >= 0.0 ? 0.0 : 1.0;
        q[2] = q[3] = 0.0;
        return;
    }
    const double lambda = sqrt(tr * tr + dn * dn);
    double w, x, y, z;
    if (tr >= 0.0) {
        w = lambda + tr;
        x = d0;
        y = d1;
        z = d2;
    } else { // (lambda + t, d) ~ (|d|, (lambda - t) d/|d|): no cancellation
        const double k = (lambda - tr) / dn;
        w = dn;
        x = k * d0;
        y = k * d1;
        z = k * d2;
    }
    const double nrm = sqrt(((w * w + x * x) + y * y) + z * z);
    q[0] = w / nrm;
    q[1] = x / nrm;
    q[2] = y / nrm;
    q[3] = z / nrm;
}

// T = Translation(mu_x - q*mu_p) * q   from the 17 sums over the kept pairs
ICPB_HD void icpb_transform_from_sums(const double *sums, double *T, double *mse)
{
    const double n_inv = 1.0 / sums[16];
    double mu_p[3], mu_x[3], sg[9];
    for (int r = 0; r < 3; ++r) {
        mu_p[r] = sums[r] * n_inv;
        mu_x[r] = sums[3 + r] * n_inv;
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) sg[r * 3 + c] = sums[6 + r * 3 + c] * n_inv - mu_p[r] * mu_x[c];
    double q[4];
    icpb_pose_quat(sg, q);
    // Quaterniond * Vector3d
    const double u0 = q[1], u1 = q[2], u2 = q[3];
    double uv0 = u1 * mu_p[2] - u2 * mu_p[1], uv1 = u2 * mu_p[0] - u0 * mu_p[2], uv2 = u0 * mu_p[1] - u1 * mu_p[0];
    uv0 += uv0;
    uv1 += uv1;
    uv2 += uv2;
    const double c0 = u1 * uv2 - u2 * uv1, c1 = u2 * uv0 - u0 * uv2, c2 = u0 * uv1 - u1 * uv0;
    const double rp0 = (mu_p[0] + q[0] * uv0) + c0, rp1 = (mu_p[1] + q[0] * uv1) + c1,
                 rp2 = (mu_p[2] + q[0] * uv2) + c2;
    // Quaterniond::toRotationMatrix
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    T[0] = 1.0 - (tyy + tzz);
    T[1] = txy - twz;
    T[2] = txz + twy;
    T[4] = txy + twz;
    T[5] = 1.0 - (txx + tzz);
    T[6] = tyz - twx;
    T[8] = txz - twy;
    T[9] = tyz + twx;
    T[10] = 1.0 - (txx + tyy);
    T[3] = mu_x[0] - rp0;
    T[7] = mu_x[1] - rp1;
    T[11] = mu_x[2] - rp2;
    *mse = sums[15] * n_inv;
}

// loop condition of Trimmed::_align (icp/icp.hpp:466)
ICPB_HD int icpb_loop_continues(const IcpState &s)
{
    return s.iter < s.max_iter && s.mse > s.min_mse && s.mse_diff > s.min_mse_diff;
}

ICPB_HD void icpb_state_init(IcpState &s, const double *Cl2g, unsigned long long max_iter, double min_mse,
                             double min_mse_diff, double overlap, unsigned long long n_po, double lazy_margin = -1.0)
{
    s.lazy_margin = lazy_margin;
    icpb_aff_identity(s.C);
    for (int i = 0; i < 12; ++i) s.Cl2g[i] = Cl2g[i];
    for (int i = 0; i < 12; ++i) s.M[i] = s.Mprev[i] = Cl2g[i]; // C_local2globalnew = C_local2global (icp/icp.hpp:118)
    s.d_box = icpb_inf();
    s.mse = s.mse_diff = icpb_inf();
    s.min_mse = min_mse;
    s.min_mse_diff = min_mse_diff;
    s.overlap = overlap;
    s.iter = 0;
    s.max_iter = max_iter;
    s.pairs = 0;
    s.n_po = n_po;
    s.found = 0;
    s.far = 0;
    s.skipped = 0;
    s.deferred = 0;
    s.tau_guard = icpb_inf(); // nothing is deferred in the first loop body
    s.force_exact = 0;
    s.redone = 0;
    s.q_count[0] = s.q_count[1] = s.q_cursor = 0;
    s.early = 0;
    s.executed = 0;
    s.warm = 0;
    s.tau = 0.0;
    s.p_cut = 0x100000000ll;
    s.need_tie = 0;
    for (int i = 0; i < 16; ++i) s.tickets[i] = 0;
    s.comm_error = 0; // comm_seq keeps counting across aligns
    s.done = icpb_loop_continues(s) ? 0 : 1;
}

// end of one loop body, after the trim (tau, kept count in s.sel_k) and the
// moment sums are known.  Writes one trace row.
ICPB_HD void icpb_iteration_finish(IcpState &s, double *trace_row)
{
    const double old_mse = s.mse;
    const unsigned long long kept = s.sel_k;
    // Lazy search: the distance entries of deferred queries are lower bounds, all above tau_guard.  The n_po smallest
    // entries -- hence tau, the kept set and the sums -- are exact iff the new tau stays at or below tau_guard; if it
    // does not, this loop body did not happen: it is run again with every query searched (or skipped) properly.
    if (s.tau_guard < icpb_inf() && kept > 0 && !(s.tau <= s.tau_guard)) {
        s.force_exact = 1;
        s.tau_guard = icpb_inf();
        s.redone += 1;
        s.found = 0;
        s.far = 0;
        s.skipped = 0;
        s.deferred = 0;
        return;
    }
    // d_box = pairs.trim(n_po) * 2.0 ; trim returns NaN when no pair is left (icp/icp.cpp:33-40)
    const double tau = kept > 0 ? s.tau : __builtin_nan("");
    s.d_box = tau * 2.0;
    s.pairs = kept;
    if (trace_row) {
        trace_row[0] = (double)s.iter;
        trace_row[1] = (double)s.found;
        trace_row[2] = (double)kept;
        trace_row[3] = tau;
        trace_row[4] = s.d_box;
        trace_row[7] = (double)s.far;
        trace_row[24] = (double)s.skipped;
        trace_row[25] = (double)s.deferred;
        trace_row[26] = (double)s.redone;
        trace_row[27] = 0.0;
    }
    if (kept < 3) { // Pairs::MIN_PAIRS (icp/icp.hpp:475-476): return result as is
        s.early = 1;
        s.done = 1;
    } else {
        double T[12], mse;
        icpb_transform_from_sums(s.sums, T, &mse);
        for (int i = 0; i < 12; ++i) s.Mprev[i] = s.M[i];
        icpb_aff_mul(T, s.C, s.C);      // C = T * C                       (icp/icp.hpp:479)
        icpb_aff_mul(s.C, s.Cl2g, s.M); // measurement.setOffsetTransform   (icp/icp.hpp:480)
        s.mse = mse;
        s.mse_diff = old_mse - mse;
        s.iter += 1;
        s.warm = 1;
        s.done = icpb_loop_continues(s) ? 0 : 1;
    }
    if (trace_row) {
        trace_row[5] = s.mse;
        trace_row[6] = s.mse_diff;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) trace_row[8 + c * 4 + r] = s.C[r * 4 + c]; // column-major 4x4
        trace_row[8 + 3] = trace_row[8 + 7] = trace_row[8 + 11] = 0.0;
        trace_row[8 + 15] = 1.0;
    }
    s.executed += 1;
    s.found = 0;
    s.far = 0;
    s.skipped = 0;
    s.deferred = 0;
    s.force_exact = 0;
    // what the next loop body may defer: everything provably farther than this loop body's tau plus a margin
    s.tau_guard = (s.lazy_margin > -1.0 && kept > 0) ? tau * (1.0 + s.lazy_margin) : icpb_inf();
    s.q_count[0] = s.q_count[1] = s.q_cursor = 0;
}
