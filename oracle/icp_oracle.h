/*
 * icp_oracle.h -- CPU restatement of envire's trimmed ICP (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle for the B200 path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  Nothing under slam-envire_b200/ links to it.
 *
 * Parity status: the reference ships no golden vectors and no assertions on
 * this path (test/unit/icp.cpp:147-162).  This restatement is pinned against
 * oracle/_ref -- the reference's OWN icp/icp.cpp + icp/icp.hpp compiled
 * unmodified against restated third-party shims (see oracle/Makefile) -- so
 * the control flow / arithmetic of Pairs, Trimmed and GoldenBracket is pinned
 * to the reference; the third-party pieces (libkdtree++, Eigen::EigenSolver,
 * neither vendored nor version-pinned by the reference) are "parity unpinned".
 *
 * All arithmetic is IEEE fp64 without FMA contraction (compile with
 * -ffp-contract=off).  Matrices crossing this API are 4x4 COLUMN-major
 * (Eigen::Affine3d::data() layout).
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NO_PAIR 0xFFFFFFFFu

typedef struct orc_model orc_model; /* FindPairsKDTree: incremental kd-tree  */
typedef struct orc_run orc_run;     /* outputs of the last _align            */

typedef struct {
    size_t max_iter;
    double min_mse, min_mse_diff, overlap;
} orc_params;

typedef struct {
    double C[16]; /* C_global2globalnew, column-major 4x4 */
    size_t iter, pairs;
    double mse, mse_diff, d_box, overlap;
} orc_result;

/* one row per executed loop body of Trimmed::_align (icp/icp.hpp:466-496) */
typedef struct {
    double iter, found, kept, tau, d_box, mse, mse_diff;
    double C[16];
} orc_trace_row;

/* ---- model (icp/icp.hpp:231-240,254-257) ---- */
orc_model *orc_model_new(void);
void orc_model_free(orc_model *);
void orc_model_clear(orc_model *);
/* addModel: density-stepped iteration (icp/icp.hpp:126-138), each vertex
 * mapped by T_local2global, inserted without rebalancing.  Accumulates. */
void orc_model_add(orc_model *, const double *xyz_aos, size_t n, const double T[16], double density);
/* FindPairsKDEAN model (icp/icp.hpp:184-193,513-515): vertices with normals (mapped by T.linear()) and the
 * SCAN_EDGE attribute bit; a model is either all-EAN or all-plain */
void orc_model_add_ean(orc_model *, const double *xyz_aos, const double *normals_aos, const int32_t *attrs, size_t n,
                       const double T[16], double density);
size_t orc_model_size(const orc_model *);
/* copy of the stored (global-frame) model points in insertion order */
void orc_model_points(const orc_model *, double *xyz_out);
int orc_model_depth(const orc_model *);

/* ---- single pieces, for kernel-level parity ---- */
/* density stepping: writes the visited indices, returns their count */
size_t orc_density_indices(size_t n, double density, size_t *out, size_t cap);
size_t orc_adapter_size(size_t n, double density);
/* p = M * v with M column-major 4x4 affine, Eigen evaluation order */
void orc_transform_points(const double M[16], const double *xyz, size_t n, double *out);
/* NN of each query (already in the global frame) within the inclusive ball
 * d_box; idx = insertion index or ORC_NO_PAIR.  threads<=0 -> all cores. */
void orc_find_pairs(const orc_model *, const double *q_xyz, size_t n, double d_box,
                    uint32_t *idx, double *dist, int threads);
void orc_find_pairs_brute(const orc_model *, const double *q_xyz, size_t n, double d_box,
                          uint32_t *idx, double *dist);
/* Pairs::trim on (dist) of found pairs: order[] receives the kept pair ids
 * sorted by (distance, id); returns kept count; *tau = largest kept or NaN */
size_t orc_trim(const double *dist, size_t n_pairs, size_t n_po, size_t *order, double *tau);
/* Pairs::getTransform (icp/icp.cpp:43-102) over kept pairs in order[] */
int orc_get_transform(const double *p_xyz, const double *x_xyz, const double *dist,
                      const size_t *order, size_t n_kept, double T[16], double *mse);
/* unit quaternion (w,x,y,z) = eigenvector of the max real eigenvalue of the
 * reference's (non-symmetric) Q matrix, in closed form; orc_build_Q exposes
 * that Q (icp/icp.cpp:76-81) so tests can cross-check with LAPACK dgeev. */
void orc_pose_quat_closed(const double sigma_px[9] /*row-major*/, double q[4]);
void orc_build_Q(const double sigma_px[9], double Q[16] /*row-major*/);

/* ---- Trimmed::_align (icp/icp.hpp:453-507) ---- */
orc_run *orc_run_new(void);
void orc_run_free(orc_run *);
int orc_align(const orc_model *, const double *xyz_aos, size_t n, const double T_local2global[16],
              double density, const orc_params *, orc_result *out, orc_run *run, int threads);
/* Trimmed<PointcloudEdgeAndNormalAdapter, FindPairsKDEAN>::_align: the pair filter of icp/icp.hpp:206-221 is
 * applied AFTER the nearest neighbour is found (a rejected source vertex gets no pair) */
int orc_align_ean(const orc_model *, const double *xyz_aos, const double *normals_aos, const int32_t *attrs, size_t n,
                  const double T_local2global[16], double density, const orc_params *, orc_result *out, orc_run *run,
                  int threads);
/* golden-section variant (icp/icp.hpp:385-401, 284-326); returns #evaluations */
int orc_align_golden(const orc_model *, const double *xyz_aos, size_t n, const double T_local2global[16],
                     double density, size_t max_iter, double min_mse, double min_mse_diff,
                     double alpha, double beta, double eps, orc_result *out, orc_run *run, int threads);
/* PointcloudAdapter::applyTransform (icp/icp.hpp:148-162): new FrameNode transform */
void orc_apply_transform(const double fn_T[16], const double C_local2global[16],
                         const double C_g2gnew[16], double out_T[16]);

size_t orc_run_pairs_distance(const orc_run *, double *out, size_t cap);
/* last executed iteration: per iterated source point, model idx / distance */
size_t orc_run_last_pairs(const orc_run *, uint32_t *idx, double *dist, size_t cap);
size_t orc_run_trace(const orc_run *, orc_trace_row *rows, size_t cap);
/* seconds spent in findPairs / trim / getTransform during the last align */
void orc_run_timing(const orc_run *, double *t_pairs, double *t_trim, double *t_solve);

#ifdef __cplusplus
}
#endif
#endif
