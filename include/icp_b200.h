/*
 * icp_b200.h -- C ABI of the B200-native trimmed-ICP hot path (libicp_b200.so).
 *
 * Drop-in boundary for envire's icp/ library (reference: rock-slam/slam-envire).
 * Each entry point names the reference interface it replaces (file:line are
 * relative to the reference tree).  Plain pointers and sizes only; all
 * pointers are HOST pointers; matrices are 4x4 COLUMN-major doubles
 * (Eigen::Affine3d::data()); points are fp64 AoS double[n][3], i.e.
 * envire::Pointcloud::vertices.data() (src/maps/Pointcloud.hpp:35).
 *
 * Every function returns an int status (ICPB_OK == 0) and never throws; the
 * C++ host mirror (slam-envire_b200/icp/icp_gpu.hpp, envire::icp::TrimmedGPU)
 * converts statuses back into the reference's exceptions.  A context owns one
 * CUDA device + stream and is not thread-safe; distinct contexts are
 * independent.  There is no CPU fallback: without a CUDA device
 * icpb_create() fails with ICPB_ERR_CUDA.
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICPB_OK 0
#define ICPB_ERR_CUDA 1         /* CUDA runtime error / no device            */
#define ICPB_ERR_ARG 2          /* bad argument                              */
#define ICPB_ERR_NO_MODEL 3     /* reserved (aligning against an empty model is legal, as in the reference: no pairs) */
#define ICPB_ERR_NO_SOURCE 4    /* align_resident() before source upload     */
#define ICPB_ERR_NCCL 5         /* NCCL missing / collective failed          */
#define ICPB_ERR_CAPACITY 6     /* output buffer too small                   */
#define ICPB_ERR_NOT_ENOUGH_PAIRS 7 /* Pairs::getTransform would throw (icp/icp.cpp:45-46) */

#define ICPB_NO_PAIR 0xFFFFFFFFu
#define ICPB_DEFERRED 0xFFFFFFFEu /* icpb_debug_pairs: a pair exists but was not resolved (lazy search, see "lazy") */
#define ICPB_NCCL_ID_BYTES 128
#define ICPB_IPC_HANDLE_BYTES 64
#define ICPB_TRACE_COLS 28

typedef struct icpb_ctx icpb_ctx;

/* Trimmed::_align arguments (icp/icp.hpp:453) */
typedef struct {
    size_t max_iter;
    double min_mse, min_mse_diff, overlap;
} icpb_params;

/* Trimmed::Result (icp/icp.hpp:334-365) */
typedef struct {
    double C_global2globalnew[16];
    size_t iter, pairs;
    double mse, mse_diff, d_box, overlap;
    int status; /* ICPB_OK, or ICPB_ERR_NOT_ENOUGH_PAIRS when _align returned early (icp/icp.hpp:475-476) */
} icpb_result;

/* ---- lifetime -------------------------------------------------------- */
/* Replaces: construction of envire::icp::TrimmedKD (icp/icp.hpp:522). */
int icpb_create(icpb_ctx **out, int device_id);
void icpb_destroy(icpb_ctx *);
const char *icpb_strerror(int status);
/* last CUDA/NCCL error text of this context (for exception messages) */
const char *icpb_last_error(const icpb_ctx *);
const char *icpb_version(void);

/* ---- tunables (GPU-only knobs; ICPConfiguration stays untouched) ------ */
/* keys: "cell_size" (m, 0 = auto), "target_pts_per_cell" (default 12), "max_cells",
 *       "trim_blocks_per_sm" / "moments_blocks_per_sm" (grid sizes of K4 / K5, default 2 / 2),
 *       "poll_every" (iterations between host polls of the done flag),
 *       "warm_start" (0/1: seed the search with the previous iteration's pair),
 *       "skip" (0/1, default 1: a query whose previous nearest neighbour provably is still the exact one -- it lies
 *       strictly inside the lower bound on all other points, reduced by the query's motion -- is not searched again),
 *       "lazy" (0/1, default 1: lazy search -- only the n_po nearest pairs survive Pairs::trim; a query whose nearest
 *       neighbour provably is farther than the previous tau * (1 + "lazy_margin") while its previous winner still lies
 *       within d_box is not searched: its pair exists and will not be kept, whichever point it is with.  If the new
 *       tau exceeds that guard the loop body is run again exactly, so results never depend on this),
 *       "skip_reach_cells" / "skip_floor_cells" (a search that needs a radius of at most `reach` cell edges covers
 *       `floor` cell edges more, so that its bound is useful; results never depend on these),
 *       "block_levels" (-1/0/1, default 1: searches whose ball spans at most 2 (0) / 3 (1) cells per axis scan that
 *       cell block flat instead of walking the pyramid),
 *       "sort_source" (0/1: order the source spatially at upload), "hilbert" (0/1, default 1: along a Hilbert curve
 *       instead of a Morton curve), "morton_bits" (cells per axis of that curve, 2^bits, default 10), "graph" (0/1, default 1: the whole _align loop is ONE CUDA-graph launch whose WHILE
 *       node is driven by the device-side loop condition), "profile" (0/1, default 0: 1 enqueues the loop kernel by
 *       kernel on the stream with CUDA events around every phase -- icpb_last_timing / icpb_iteration_times --
 *       and the host peeking at a pinned flag at most "run_ahead" iterations behind),
 *       "growth_margin" (fraction of the model extent left free around it at a full build, so that later
 *       addToModel calls inside that box are merged in without a rebuild; default 0), "merge_append" (0/1) */
int icpb_set_option(icpb_ctx *, const char *key, double value);
int icpb_get_option(const icpb_ctx *, const char *key, double *value);

/* ---- model ------------------------------------------------------------ */
/* Replaces: Trimmed::addToModel(PointcloudAdapter(model,density)) ->
 * FindPairsKDTree::addModel (icp/icp.hpp:424-428, 231-240): visits vertices
 * with the density stepping of PointcloudAdapter::next (icp/icp.hpp:126-138),
 * maps each by T_local2global, ACCUMULATES across calls. */
int icpb_model_add(icpb_ctx *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16],
                   double density);
/* Replaces: TrimmedKDEAN::addToModel(PointcloudEdgeAndNormalAdapter(model,density)) (icp/icp.hpp:176-198,513-521):
 * vertices with per-vertex normals (VERTEX_NORMAL, mapped by T.linear()) and attributes (VERTEX_ATTRIBUTES, int per
 * vertex; bit (1 << SCAN_EDGE) marks scan edges).  A model is either all edge-and-normal or all plain. */
int icpb_model_add_ean(icpb_ctx *, const double *xyz_aos, const double *normals_aos, const int32_t *attrs,
                       size_t n_vertices, const double T_local2global[16], double density);
/* Replaces: Trimmed::clearModel (icp/icp.hpp:432). */
int icpb_model_clear(icpb_ctx *);
int icpb_model_size(const icpb_ctx *, size_t *n_points);
/* force the (otherwise lazy) grid build; reports the grid that was built */
typedef struct {
    double origin[3], cell_size;
    uint32_t dims[3], levels;
    uint64_t n_cells, n_points, n_occupied;
    double build_ms;
    uint64_t merges; /* addToModel calls served by a streaming merge into the existing grid instead of a rebuild */
} icpb_grid_info;
int icpb_model_build(icpb_ctx *, icpb_grid_info *info /* may be NULL */);

/* ---- align ------------------------------------------------------------ */
/* Upload the measurement: PointcloudAdapter(meas,density) construction +
 * reset() (icp/icp.hpp:102-119,139-142).  The original vertices stay
 * resident; every iteration re-derives C_local2globalnew * v from them
 * (setOffsetTransform, icp/icp.hpp:121-124). */
int icpb_source_upload(icpb_ctx *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16],
                       double density);
/* The same for a PointcloudEdgeAndNormalAdapter measurement: the following aligns apply EdgeAndNormalPairFilter
 * (icp/icp.hpp:206-221) to each nearest neighbour (needs an edge-and-normal model). */
int icpb_source_upload_ean(icpb_ctx *, const double *xyz_aos, const double *normals_aos, const int32_t *attrs,
                           size_t n_vertices, const double T_local2global[16], double density);
/* One Trimmed::_align (icp/icp.hpp:453-507) on the resident source. */
int icpb_align_resident(icpb_ctx *, const icpb_params *, icpb_result *out);
/* upload + _align in one call: what Trimmed::align(meas, max_iter, min_mse,
 * min_mse_diff, overlap) does before applyTransform (icp/icp.hpp:413-418).
 * applyTransform / FrameNode::setTransform and the golden-section variant
 * (icp/icp.hpp:385-401) stay in the host mirror. */
int icpb_align(icpb_ctx *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16],
               double density, const icpb_params *, icpb_result *out);

/* Replaces: Trimmed::getPairsDistance (icp/icp.hpp:439, filled at :500-504):
 * kept distances of the last executed iteration, ascending; available until the next align / source upload (model
 * changes in between do not disturb them).  With
 * out == NULL only *n_out is written (the count of all ranks' kept pairs).  After a sharded (multi-GPU) align every
 * rank returns the kept distances of ITS shard (*n_out = their number): concatenate and merge for the full list. */
int icpb_pairs_distance(icpb_ctx *, double *out_sorted, size_t capacity, size_t *n_out);
/* Parity hook: correspondences of the last executed iteration, one entry per
 * visited source vertex: model index in insertion order (ICPB_NO_PAIR if
 * none within d_box; ICPB_DEFERRED: a pair exists, beyond what the trim keeps, not resolved -- the distance then is a
 * lower bound), Euclidean distance, kept flag after the trim. */
int icpb_debug_pairs(icpb_ctx *, uint32_t *model_idx, double *dist, uint8_t *kept, size_t capacity, size_t *n_out);
/* Per-iteration trace (the reference's commented-out trace, icp/icp.hpp:487-495):
 * rows of ICPB_TRACE_COLS doubles:
 * [0]=iter [1]=found [2]=kept [3]=tau [4]=d_box [5]=mse [6]=mse_diff [7]=far_queries [8..23]=C (col-major)
 * [24]=queries answered by the skip test (no search) [25]=queries deferred by the lazy search
 * [26]=loop bodies run again so far because a lazy trim did not verify [27]=reserved */
int icpb_trace(icpb_ctx *, double *rows, size_t capacity_rows, size_t *n_rows);
/* device time (ms, CUDA events) of the last align: total, and summed per phase */
typedef struct {
    double total_ms, upload_ms, search_ms, trim_ms, solve_ms;
    uint64_t iterations, kernels_launched;
} icpb_timing;
int icpb_last_timing(const icpb_ctx *, icpb_timing *out);
/* per-iteration device times (ms) of the last align: rows of 3 doubles [search, trim, moments+pose] */
int icpb_iteration_times(const icpb_ctx *, double *rows, size_t capacity_rows, size_t *n_rows);

/* ---- kernel-level entry points (parity tests of single stages) -------- */
/* K3: exact NN of n query points (already in the global frame) against the
 * model within the inclusive ball d_box (FindPairsKDTree::findPairs ->
 * kdtree.find_nearest(node, d_box), icp/icp.hpp:242-252). */
int icpb_find_pairs(icpb_ctx *, const double *q_xyz_aos, size_t n, double d_box, uint32_t *model_idx, double *dist);
/* K4: Pairs::trim (icp/icp.cpp:23-41) on n distances (+inf = no pair):
 * *tau = n_po-th smallest (NaN if none), *kept = min(n_po, found),
 * *n_ties_kept = how many entries equal to tau are kept */
int icpb_trim(icpb_ctx *, const double *dist, size_t n, size_t n_po, double *tau, size_t *kept, size_t *n_ties_kept);

/* K4+K5+K6 on caller-supplied pairs: Pairs::add xN, Pairs::trim(n_po), Pairs::getTransform (icp/icp.cpp:12-102).
 * x = model points, p = source points (fp64 AoS), dist = their distances.  Outputs the transform that moves p onto x
 * (col-major 4x4), the mean square distance over the kept pairs, tau = largest kept distance and the kept count.
 * Returns ICPB_ERR_NOT_ENOUGH_PAIRS when fewer than 3 pairs are left (where the reference throws). */
int icpb_pairs_transform(icpb_ctx *, const double *x_aos, const double *p_aos, const double *dist, size_t n, size_t n_po,
                         double T_out[16], double *mse, double *tau, size_t *kept);

/* ---- multi-GPU (one process per GPU; source sharded, model replicated) - */
/* rank 0 creates the id and ships the ICPB_NCCL_ID_BYTES to its peers by any
 * channel (torch.distributed broadcast in bench.py), then every rank calls
 * icpb_comm_init.  After that icpb_source_upload takes this rank's shard and
 * icpb_align_resident runs the two per-iteration collectives (trim histogram,
 * 17 moment sums) over NCCL. */
int icpb_nccl_unique_id(uint8_t id_out[ICPB_NCCL_ID_BYTES]);
int icpb_comm_init(icpb_ctx *, int n_ranks, int rank, const uint8_t id[ICPB_NCCL_ID_BYTES]);
/* Preferred on one NVLink/NVSwitch box: the per-iteration collectives are FUSED into the producing kernels and run
 * over peer memory instead of NCCL.  Every rank exports the IPC handle of its exchange buffer, the handles are
 * all-gathered by any channel, then every rank connects (n_ranks <= 8).  Results are identical to the NCCL path. */
int icpb_p2p_export(icpb_ctx *, uint8_t handle_out[ICPB_IPC_HANDLE_BYTES]);
int icpb_p2p_connect(icpb_ctx *, int n_ranks, int rank, const uint8_t *handles /* n_ranks * ICPB_IPC_HANDLE_BYTES */);
/* shard upload: vertices already density-stepped by the caller, plus the
 * GLOBAL PointcloudAdapter::size() (icp/icp.hpp:143-146) that n_po is derived from */
int icpb_source_upload_shard(icpb_ctx *, const double *xyz_aos_local, size_t n_local, size_t adapter_size_global,
                             const double T_local2global[16]);


/* ---- several GPUs in ONE process ---------------------------------------- */
/* The same sharding (source in contiguous ranges of the adapter's visiting order, model replicated) and the same fused
 * exchanges, for callers that are a single process -- envire::icp::TrimmedGPU(device, n_devices): one context per
 * device, peers reached through cudaDeviceEnablePeerAccess (n_devices <= 8, distinct devices of one NVLink box).
 * The calls mirror the single-context ones; results equal the single-GPU ones (same pose to fp64 re-association of the
 * 17 sums, same iteration and pair counts).  Plain (non edge-and-normal) clouds; needs the default graph loop. */
typedef struct icpb_multi icpb_multi;
int icpb_create_multi(icpb_multi **out, const int *device_ids, int n_devices); /* *out is set even on failure (for the error text): destroy it */
void icpb_destroy_multi(icpb_multi *);
const char *icpb_multi_last_error(const icpb_multi *);
int icpb_multi_size(const icpb_multi *);
icpb_ctx *icpb_multi_context(icpb_multi *, int i); /* the context of device i: traces, timings, grid info */
int icpb_multi_set_option(icpb_multi *, const char *key, double value);
int icpb_multi_model_add(icpb_multi *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16], double density);
int icpb_multi_model_clear(icpb_multi *);
int icpb_multi_source_upload(icpb_multi *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16], double density);
int icpb_multi_align_resident(icpb_multi *, const icpb_params *, icpb_result *out);
int icpb_multi_align(icpb_multi *, const double *xyz_aos, size_t n_vertices, const double T_local2global[16], double density,
                     const icpb_params *, icpb_result *out);
/* kept distances of all devices, ascending (merged on the host) */
int icpb_multi_pairs_distance(icpb_multi *, double *out_sorted, size_t capacity, size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
