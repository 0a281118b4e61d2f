/*
 * icp_oracle.c -- plain-C restatement of envire's trimmed ICP hot path.
 * TEST INFRASTRUCTURE (see icp_oracle.h): never linked into the product.
 *
 * Follows, function by function (paths relative to the reference tree):
 *   Pairs::add / trim / getTransform ........ icp/icp.cpp:12-41, 43-102
 *   PointcloudAdapter (density stepping,
 *     setOffsetTransform, applyTransform) .... icp/icp.hpp:102-162
 *   FindPairsKDTree::addModel / findPairs .... icp/icp.hpp:231-252
 *   GoldenBracket::findMin ................... icp/icp.hpp:284-326
 *   Trimmed::Result, align x2, _align ........ icp/icp.hpp:334-365, 385-418, 453-507
 * Third-party behaviour restated from published semantics (absent from the
 * reference tree, versions unpinned there): libkdtree++ KDTree<3,VertexNode>
 * (insert without optimise(); find_nearest(val,max) = exact Euclidean NN with
 * distance <= max, sqrt'ed distances) and Eigen3 (Affine3d products,
 * Quaterniond rotation, EigenSolver<Matrix4d> -> closed form, see
 * orc_pose_quat_closed).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include "icp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ */
/* small affine algebra in Eigen's evaluation order                    */
/* ------------------------------------------------------------------ */
typedef struct {
    double L[3][3];
    double t[3];
} aff;

static void aff_identity(aff *a)
{
    memset(a, 0, sizeof(*a));
    a->L[0][0] = a->L[1][1] = a->L[2][2] = 1.0;
}
static void aff_from_cm(const double M[16], aff *a)
{
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) a->L[r][c] = M[c * 4 + r];
        a->t[r] = M[12 + r];
    }
}
static void aff_to_cm(const aff *a, double M[16])
{
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) M[c * 4 + r] = a->L[r][c];
        M[12 + r] = a->t[r];
        M[r * 4 + 3] = 0.0;
    }
    M[15] = 1.0;
}
/* p = linear*v + translation */
static inline void aff_apply(const aff *a, const double v[3], double p[3])
{
    for (int r = 0; r < 3; r++)
        p[r] = ((a->L[r][0] * v[0] + a->L[r][1] * v[1]) + a->L[r][2] * v[2]) + a->t[r];
}
/* res.linear = a.linear*b.linear; res.translation = a.linear*b.translation + a.translation */
static void aff_mul(const aff *a, const aff *b, aff *res)
{
    aff o;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
            o.L[r][c] = (a->L[r][0] * b->L[0][c] + a->L[r][1] * b->L[1][c]) + a->L[r][2] * b->L[2][c];
        o.t[r] = ((a->L[r][0] * b->t[0] + a->L[r][1] * b->t[1]) + a->L[r][2] * b->t[2]) + a->t[r];
    }
    *res = o;
}
/* Transform::inverse(Eigen::Isometry): linear^T, -linear^T * translation */
static void aff_inverse_isometry(const aff *a, aff *res)
{
    aff o;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o.L[r][c] = a->L[c][r];
    for (int r = 0; r < 3; r++)
        o.t[r] = -((o.L[r][0] * a->t[0] + o.L[r][1] * a->t[1]) + o.L[r][2] * a->t[2]);
    *res = o;
}

static double det3(const double m[3][3])
{
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
           m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
/* Transform::rotation(): the orthogonal polar factor U*V^T of linear() (Eigen
 * gets it from a JacobiSVD); computed here by Newton's iteration
 * X <- (X + X^-T)/2, which converges to the same factor. */
static void polar_rotation(const double m[3][3], double R[3][3])
{
    double X[3][3];
    memcpy(X, m, sizeof(X));
    for (int it = 0; it < 100; it++) {
        double d = det3(X);
        double inv_t[3][3]; /* X^-T = cofactor / det */
        inv_t[0][0] = (X[1][1] * X[2][2] - X[1][2] * X[2][1]) / d;
        inv_t[0][1] = (X[1][2] * X[2][0] - X[1][0] * X[2][2]) / d;
        inv_t[0][2] = (X[1][0] * X[2][1] - X[1][1] * X[2][0]) / d;
        inv_t[1][0] = (X[0][2] * X[2][1] - X[0][1] * X[2][2]) / d;
        inv_t[1][1] = (X[0][0] * X[2][2] - X[0][2] * X[2][0]) / d;
        inv_t[1][2] = (X[0][1] * X[2][0] - X[0][0] * X[2][1]) / d;
        inv_t[2][0] = (X[0][1] * X[1][2] - X[0][2] * X[1][1]) / d;
        inv_t[2][1] = (X[0][2] * X[1][0] - X[0][0] * X[1][2]) / d;
        inv_t[2][2] = (X[0][0] * X[1][1] - X[0][1] * X[1][0]) / d;
        double delta = 0.0;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                double nx = 0.5 * (X[r][c] + inv_t[r][c]);
                delta = fmax(delta, fabs(nx - X[r][c]));
                X[r][c] = nx;
            }
        if (delta < 1e-16) break;
    }
    memcpy(R, X, sizeof(X));
}

/* ------------------------------------------------------------------ */
/* density stepping (icp/icp.hpp:126-146)                              */
/* ------------------------------------------------------------------ */
size_t orc_density_indices(size_t n, double density, size_t *out, size_t cap)
{
    double index = 0.0; /* reset() */
    size_t cnt = 0;
    for (;;) {
        const size_t idx = (size_t)index; /* hasNext(): idx < vertices->size() */
        if (!(idx < n)) break;
        if (out && cnt < cap) out[cnt] = idx;
        cnt++;
        index += 1.0 / density; /* next() */
    }
    return cnt;
}
size_t orc_adapter_size(size_t n, double density)
{
    return (size_t)((double)n * density); /* vertices->size() * density */
}
void orc_transform_points(const double M[16], const double *xyz, size_t n, double *out)
{
    aff a;
    aff_from_cm(M, &a);
    for (size_t i = 0; i < n; i++) aff_apply(&a, xyz + 3 * i, out + 3 * i);
}

/* ------------------------------------------------------------------ */
/* kd-tree: libkdtree++ usage in FindPairsKDTree (icp/icp.hpp:231-252)  */
/* ------------------------------------------------------------------ */
struct orc_model {
    double *pt; /* 3*n, global frame, insertion order */
    double *nrm;          /* 3*n normals in the global frame (EAN models only) */
    unsigned char *edge;  /* n edge flags (EAN models only) */
    int ean;
    int32_t *left, *right;
    size_t n, cap;
    int depth;
};

orc_model *orc_model_new(void)
{
    return (orc_model *)calloc(1, sizeof(orc_model));
}
void orc_model_clear(orc_model *m)
{
    m->n = 0;
    m->depth = 0;
    m->ean = 0;
}
void orc_model_free(orc_model *m)
{
    if (!m) return;
    free(m->pt);
    free(m->nrm);
    free(m->edge);
    free(m->left);
    free(m->right);
    free(m);
}
size_t orc_model_size(const orc_model *m) { return m->n; }
int orc_model_depth(const orc_model *m) { return m->depth; }
void orc_model_points(const orc_model *m, double *out) { memcpy(out, m->pt, m->n * 3 * sizeof(double)); }

static void model_reserve(orc_model *m, size_t want)
{
    if (want <= m->cap) return;
    size_t nc = m->cap ? m->cap : 1024;
    while (nc < want) nc *= 2;
    m->pt = (double *)realloc(m->pt, nc * 3 * sizeof(double));
    m->nrm = (double *)realloc(m->nrm, nc * 3 * sizeof(double));
    m->edge = (unsigned char *)realloc(m->edge, nc);
    m->left = (int32_t *)realloc(m->left, nc * sizeof(int32_t));
    m->right = (int32_t *)realloc(m->right, nc * sizeof(int32_t));
    m->cap = nc;
}

/* KDTree::insert: descend comparing coordinate (level % 3); strictly-less
 * goes left, otherwise right; no rebalancing (optimise() is never called). */
static void model_insert(orc_model *m, const double p[3])
{
    model_reserve(m, m->n + 1);
    const size_t id = m->n++;
    memcpy(m->pt + 3 * id, p, 3 * sizeof(double));
    m->left[id] = m->right[id] = -1;
    if (id == 0) {
        m->depth = 1;
        return;
    }
    int32_t cur = 0;
    int level = 0;
    for (;;) {
        const int dim = level % 3;
        int32_t *next = (p[dim] < m->pt[3 * cur + dim]) ? &m->left[cur] : &m->right[cur];
        level++;
        if (*next < 0) {
            *next = (int32_t)id;
            break;
        }
        cur = *next;
    }
    if (level + 1 > m->depth) m->depth = level + 1;
}

void orc_model_add(orc_model *m, const double *xyz, size_t n, const double T[16], double density)
{
    aff a;
    aff_from_cm(T, &a);
    double index = 0.0;
    for (;;) {
        const size_t idx = (size_t)index;
        if (!(idx < n)) break;
        double p[3];
        aff_apply(&a, xyz + 3 * idx, p);
        model_insert(m, p);
        index += 1.0 / density;
    }
}

#define ORC_SCAN_EDGE_BIT (1 << 0x01) /* attrs[idx] & (1 << envire::Pointcloud::SCAN_EDGE), SCAN_EDGE = 0x01 */

void orc_model_add_ean(orc_model *m, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                       const double T[16], double density)
{
    aff a;
    aff_from_cm(T, &a);
    double index = 0.0;
    m->ean = 1;
    for (;;) {
        const size_t idx = (size_t)index;
        if (!(idx < n)) break;
        double p[3];
        aff_apply(&a, xyz + 3 * idx, p);
        model_insert(m, p);
        const size_t id = m->n - 1;
        const double *v = normals + 3 * idx;
        for (int r = 0; r < 3; r++) /* C_local2globalnew.linear() * normal */
            m->nrm[3 * id + r] = (a.L[r][0] * v[0] + a.L[r][1] * v[1]) + a.L[r][2] * v[2];
        m->edge[id] = (attrs[idx] & ORC_SCAN_EDGE_BIT) ? 1 : 0;
        index += 1.0 / density;
    }
}

/* libkdtree++ node distance: sqrt(sum of squared coordinate differences) */
static inline double node_distance(const double a[3], const double b[3])
{
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

/* find_nearest(val, max): exact NN with distance <= max (inclusive), or none.
 * Equidistant candidates: the lowest insertion index wins (the library's
 * order is unspecified; parity tests exempt ties). */
static void model_nearest(const orc_model *m, const double q[3], double max, uint32_t *o_idx, double *o_d,
                          int32_t *stack)
{
    double best = max;
    uint32_t best_idx = ORC_NO_PAIR;
    if (m->n == 0) {
        *o_idx = ORC_NO_PAIR;
        *o_d = max;
        return;
    }
    int sp = 0;
    /* stack entries: node, level; far branches are pushed with their plane distance re-checked on pop */
    stack[sp++] = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        int level = stack[--sp];
        int32_t node = stack[--sp];
        int deferred = 0;
        if (level < 0) { /* deferred far branch: level encoded as -(level+1) */
            level = -(level + 1);
            deferred = 1;
        }
        if (deferred) {
            /* node here is the PARENT whose far side is pending */
            const int dim = level % 3;
            const double diff = q[dim] - m->pt[3 * node + dim];
            if (!(fabs(diff) <= best)) continue;
            node = (diff < 0.0) ? m->right[node] : m->left[node];
            if (node < 0) continue;
            level++;
        }
        while (node >= 0) {
            const double *np = m->pt + 3 * node;
            const double d = node_distance(q, np);
            if (d < best || (d == best && (best_idx == ORC_NO_PAIR || (uint32_t)node < best_idx))) {
                if (d <= max) {
                    best = d;
                    best_idx = (uint32_t)node;
                }
            }
            const int dim = level % 3;
            const double diff = q[dim] - np[dim];
            /* remember the far side, descend the near side */
            stack[sp++] = node;
            stack[sp++] = -(level + 1);
            node = (diff < 0.0) ? m->left[node] : m->right[node];
            level++;
        }
    }
    *o_idx = best_idx;
    *o_d = best;
}

void orc_find_pairs(const orc_model *m, const double *q, size_t n, double d_box, uint32_t *idx, double *dist,
                    int threads)
{
    const size_t stack_len = 2 * ((size_t)m->depth + 4);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
    {
        int32_t *stack = (int32_t *)malloc(stack_len * sizeof(int32_t));
#pragma omp for schedule(dynamic, 1024)
        for (long long i = 0; i < (long long)n; i++) model_nearest(m, q + 3 * i, d_box, &idx[i], &dist[i], stack);
        free(stack);
    }
#else
    (void)threads;
    int32_t *stack = (int32_t *)malloc(stack_len * sizeof(int32_t));
    for (size_t i = 0; i < n; i++) model_nearest(m, q + 3 * i, d_box, &idx[i], &dist[i], stack);
    free(stack);
#endif
}

void orc_find_pairs_brute(const orc_model *m, const double *q, size_t n, double d_box, uint32_t *idx, double *dist)
{
    for (size_t i = 0; i < n; i++) {
        double best = d_box;
        uint32_t bi = ORC_NO_PAIR;
        for (size_t j = 0; j < m->n; j++) {
            const double d = node_distance(q + 3 * i, m->pt + 3 * j);
            if (d <= d_box && (d < best || (d == best && bi == ORC_NO_PAIR))) {
                best = d;
                bi = (uint32_t)j;
            }
        }
        idx[i] = bi;
        dist[i] = best;
    }
}

/* ------------------------------------------------------------------ */
/* Pairs::trim (icp/icp.cpp:23-41)                                      */
/* ------------------------------------------------------------------ */
typedef struct {
    size_t index;
    double distance;
} pair_t;

static int pair_cmp(const void *a, const void *b)
{
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->distance < y->distance) return -1;
    if (x->distance > y->distance) return 1;
    /* std::sort leaves the order of equal distances unspecified; fix it by
     * pair index so the oracle is deterministic */
    return (x->index > y->index) - (x->index < y->index);
}

size_t orc_trim(const double *dist, size_t n_pairs, size_t n_po, size_t *order, double *tau)
{
    pair_t *pairs = (pair_t *)malloc((n_pairs ? n_pairs : 1) * sizeof(pair_t));
    for (size_t i = 0; i < n_pairs; i++) {
        pairs[i].index = i;
        pairs[i].distance = dist[i];
    }
    qsort(pairs, n_pairs, sizeof(pair_t), pair_cmp);
    size_t kept = n_pairs;
    if (n_po < kept) kept = n_po; /* pairs.resize(n_po) */
    for (size_t i = 0; i < kept; i++) order[i] = pairs[i].index;
    *tau = kept > 0 ? pairs[kept - 1].distance : NAN;
    free(pairs);
    return kept;
}

/* ------------------------------------------------------------------ */
/* Pairs::getTransform (icp/icp.cpp:43-102)                             */
/* ------------------------------------------------------------------ */
void orc_build_Q(const double s[9], double Q[16])
{
    /* A = sigma - sigma^T; delta = (A(1,2), A(2,0), A(0,1)) */
    double A[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) A[r][c] = s[r * 3 + c] - s[c * 3 + r];
    const double delta[3] = {A[1][2], A[2][0], A[0][1]};
    const double tr = (s[0] + s[4]) + s[8];
    /* q_px << trace, delta^T, delta, A - I*trace  -- note: A, not sigma+sigma^T */
    Q[0] = tr;
    for (int c = 0; c < 3; c++) Q[1 + c] = delta[c];
    for (int r = 0; r < 3; r++) {
        Q[(r + 1) * 4] = delta[r];
        for (int c = 0; c < 3; c++) Q[(r + 1) * 4 + 1 + c] = A[r][c] - (r == c ? tr : 0.0);
    }
}

/* The reference feeds Q = [[t, d^T],[d, A - t I]] (A antisymmetric, A u = -d x u)
 * to EigenSolver and takes the eigenvector of the largest real eigenvalue.
 * For u parallel to d:  Q (w, s d) = lambda (w, s d)  with lambda^2 = t^2 + |d|^2,
 * w = (lambda + t) s; the spectrum is {+-lambda, -t +- i|d|}, so the largest
 * real eigenvalue is +lambda and its eigenvector is (lambda + t, d), i.e. a
 * rotation about d by atan2(|d|, t).  (lambda+t, d) ~ (|d|, (lambda-t) d/|d|)
 * avoids the cancellation when t < 0.  d = 0: identity. */
void orc_pose_quat_closed(const double s[9], double q[4])
{
    const double d0 = s[1 * 3 + 2] - s[2 * 3 + 1];
    const double d1 = s[2 * 3 + 0] - s[0 * 3 + 2];
    const double d2 = s[0 * 3 + 1] - s[1 * 3 + 0];
    const double tr = (s[0] + s[4]) + s[8];
    const double dn = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    if (dn == 0.0) {
        q[0] = 1.0;
        q[1] = q[2] = q[3] = 0.0;
        return;
    }
    const double lambda = sqrt(tr * tr + dn * dn);
    double w, x, y, z;
    if (tr >= 0.0) {
        w = lambda + tr;
        x = d0;
        y = d1;
        z = d2;
    } else {
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

/* Quaterniond::toRotationMatrix */
static void quat_to_matrix(const double q[4], double R[3][3])
{
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0][0] = 1.0 - (tyy + tzz);
    R[0][1] = txy - twz;
    R[0][2] = txz + twy;
    R[1][0] = txy + twz;
    R[1][1] = 1.0 - (txx + tzz);
    R[1][2] = tyz - twx;
    R[2][0] = txz - twy;
    R[2][1] = tyz + twx;
    R[2][2] = 1.0 - (txx + tyy);
}
/* Quaterniond * Vector3d: v + w*(2 u x v) + u x (2 u x v) */
static void quat_rotate(const double q[4], const double v[3], double o[3])
{
    const double u[3] = {q[1], q[2], q[3]};
    double uv[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    for (int i = 0; i < 3; i++) uv[i] += uv[i];
    const double c[3] = {u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0]};
    for (int i = 0; i < 3; i++) o[i] = (v[i] + q[0] * uv[i]) + c[i];
}

static int get_transform(const double *p, const double *x, const double *dist, const size_t *order, size_t n,
                         aff *T, double *mse)
{
    if (n < 3) return -1; /* throws std::runtime_error("not enough pairs to get transform") */
    double mu_p[3] = {0, 0, 0}, mu_x[3] = {0, 0, 0}, sg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double mu_d = 0.0;
    for (size_t i = 0; i < n; i++) {
        const size_t idx = order[i];
        const double *pv = p + 3 * idx, *xv = x + 3 * idx;
        const double d = dist[idx];
        mu_d += d * d;
        for (int r = 0; r < 3; r++) {
            mu_p[r] += pv[r];
            mu_x[r] += xv[r];
        }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) sg[r * 3 + c] += pv[r] * xv[c]; /* pv * xv^T */
    }
    const double n_inv = 1.0 / (double)n;
    for (int r = 0; r < 3; r++) {
        mu_p[r] *= n_inv;
        mu_x[r] *= n_inv;
    }
    mu_d *= n_inv;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) sg[r * 3 + c] = sg[r * 3 + c] * n_inv - mu_p[r] * mu_x[c];

    double q[4];
    orc_pose_quat_closed(sg, q);
    double rp[3];
    quat_rotate(q, mu_p, rp);
    quat_to_matrix(q, T->L);
    for (int r = 0; r < 3; r++) T->t[r] = mu_x[r] - rp[r]; /* q_T = mu_x - q_R * mu_p */
    *mse = mu_d;
    return 0;
}

int orc_get_transform(const double *p, const double *x, const double *dist, const size_t *order, size_t n,
                      double T[16], double *mse)
{
    aff a;
    int rc = get_transform(p, x, dist, order, n, &a, mse);
    if (rc == 0) aff_to_cm(&a, T);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Trimmed::_align                                                      */
/* ------------------------------------------------------------------ */
struct orc_run {
    double *pairs_distance;
    size_t n_pairs_distance;
    uint32_t *last_idx;
    double *last_dist;
    size_t n_last;
    orc_trace_row *trace;
    size_t n_trace, cap_trace;
    double t_pairs, t_trim, t_solve;
};

orc_run *orc_run_new(void) { return (orc_run *)calloc(1, sizeof(orc_run)); }
void orc_run_free(orc_run *r)
{
    if (!r) return;
    free(r->pairs_distance);
    free(r->last_idx);
    free(r->last_dist);
    free(r->trace);
    free(r);
}
static void run_reset(orc_run *r)
{
    free(r->pairs_distance);
    free(r->last_idx);
    free(r->last_dist);
    free(r->trace);
    memset(r, 0, sizeof(*r));
}
size_t orc_run_pairs_distance(const orc_run *r, double *out, size_t cap)
{
    if (out) memcpy(out, r->pairs_distance, (r->n_pairs_distance < cap ? r->n_pairs_distance : cap) * sizeof(double));
    return r->n_pairs_distance;
}
size_t orc_run_last_pairs(const orc_run *r, uint32_t *idx, double *dist, size_t cap)
{
    const size_t n = r->n_last < cap ? r->n_last : cap;
    if (idx) memcpy(idx, r->last_idx, n * sizeof(uint32_t));
    if (dist) memcpy(dist, r->last_dist, n * sizeof(double));
    return r->n_last;
}
size_t orc_run_trace(const orc_run *r, orc_trace_row *rows, size_t cap)
{
    if (rows) memcpy(rows, r->trace, (r->n_trace < cap ? r->n_trace : cap) * sizeof(orc_trace_row));
    return r->n_trace;
}
void orc_run_timing(const orc_run *r, double *a, double *b, double *c)
{
    *a = r->t_pairs;
    *b = r->t_trim;
    *c = r->t_solve;
}
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static void trace_push(orc_run *run, const orc_trace_row *row)
{
    if (run->n_trace == run->cap_trace) {
        run->cap_trace = run->cap_trace ? 2 * run->cap_trace : 64;
        run->trace = (orc_trace_row *)realloc(run->trace, run->cap_trace * sizeof(orc_trace_row));
    }
    run->trace[run->n_trace++] = *row;
}

static int align_core(const orc_model *m, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                      const double T_l2g[16], double density, const orc_params *prm, orc_result *out, orc_run *run,
                      int threads);

int orc_align(const orc_model *m, const double *xyz, size_t n, const double T_l2g[16], double density,
              const orc_params *prm, orc_result *out, orc_run *run, int threads)
{
    return align_core(m, xyz, NULL, NULL, n, T_l2g, density, prm, out, run, threads);
}

int orc_align_ean(const orc_model *m, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                  const double T_l2g[16], double density, const orc_params *prm, orc_result *out, orc_run *run,
                  int threads)
{
    if (!m->ean && m->n) return -1;
    return align_core(m, xyz, normals, attrs, n, T_l2g, density, prm, out, run, threads);
}

static int align_core(const orc_model *m, const double *xyz, const double *normals, const int32_t *attrs, size_t n,
                      const double T_l2g[16], double density, const orc_params *prm, orc_result *out, orc_run *run,
                      int threads)
{
    orc_run *own = NULL;
    if (!run) run = own = orc_run_new();
    run_reset(run);

    /* the adapter's visiting order is fixed by density; materialise it once */
    const size_t n_it = orc_density_indices(n, density, NULL, 0);
    size_t *sel = (size_t *)malloc((n_it ? n_it : 1) * sizeof(size_t));
    orc_density_indices(n, density, sel, n_it);
    double *v = (double *)malloc((n_it ? n_it : 1) * 3 * sizeof(double));
    for (size_t k = 0; k < n_it; k++) memcpy(v + 3 * k, xyz + 3 * sel[k], 3 * sizeof(double));
    double *vn = NULL;
    unsigned char *vedge = NULL;
    if (normals) { /* PointcloudEdgeAndNormalAdapter: normals and the SCAN_EDGE bit ride along */
        vn = (double *)malloc((n_it ? n_it : 1) * 3 * sizeof(double));
        vedge = (unsigned char *)malloc(n_it ? n_it : 1);
        for (size_t k = 0; k < n_it; k++) {
            memcpy(vn + 3 * k, normals + 3 * sel[k], 3 * sizeof(double));
            vedge[k] = (attrs[sel[k]] & ORC_SCAN_EDGE_BIT) ? 1 : 0;
        }
    }
    free(sel);
    const size_t meas_size = orc_adapter_size(n, density);

    double *node = (double *)malloc((n_it ? n_it : 1) * 3 * sizeof(double)); /* C_local2globalnew * v */
    uint32_t *nn_idx = (uint32_t *)malloc((n_it ? n_it : 1) * sizeof(uint32_t));
    double *nn_d = (double *)malloc((n_it ? n_it : 1) * sizeof(double));
    /* Pairs: x (model pts), p (source pts), distance -- only for found pairs */
    double *px = (double *)malloc((n_it ? n_it : 1) * 3 * sizeof(double));
    double *pp = (double *)malloc((n_it ? n_it : 1) * 3 * sizeof(double));
    double *pd = (double *)malloc((n_it ? n_it : 1) * sizeof(double));
    size_t *order = (size_t *)malloc((n_it ? n_it : 1) * sizeof(size_t));

    aff C, C_l2g, C_l2g_new;
    aff_identity(&C);
    aff_from_cm(T_l2g, &C_l2g);
    C_l2g_new = C_l2g;

    orc_result res;
    memset(&res, 0, sizeof(res));
    res.iter = 0;
    res.overlap = prm->overlap;
    res.d_box = INFINITY;
    res.mse_diff = res.mse = INFINITY;
    double old_mse = res.mse;
    size_t kept = 0;
    int early = 0;
    (void)old_mse;

    while (res.iter < prm->max_iter && res.mse > prm->min_mse && res.mse_diff > prm->min_mse_diff) {
        old_mse = res.mse;
        /* pairs.clear(); findPairs.findPairs(measurement, pairs, d_box) */
        double t0 = now_s();
        for (size_t k = 0; k < n_it; k++) aff_apply(&C_l2g_new, v + 3 * k, node + 3 * k);
        orc_find_pairs(m, node, n_it, res.d_box, nn_idx, nn_d, threads);
        if (vn) { /* EdgeAndNormalPairFilter (icp/icp.hpp:206-221), applied to the nearest neighbour only */
            for (size_t k = 0; k < n_it; k++) {
                if (nn_idx[k] == ORC_NO_PAIR) continue;
                const double *sn = vn + 3 * k, *mn = m->nrm + 3 * (size_t)nn_idx[k];
                double an[3];
                for (int r = 0; r < 3; r++)
                    an[r] = (C_l2g_new.L[r][0] * sn[0] + C_l2g_new.L[r][1] * sn[1]) + C_l2g_new.L[r][2] * sn[2];
                const double dot = (an[0] * mn[0] + an[1] * mn[1]) + an[2] * mn[2];
                const int edge = vedge[k] || m->edge[nn_idx[k]];
                const double normal_angle = acos(dot < 1.0 ? dot : 1.0); /* std::min(1.0, dot) */
                if (!(!edge && (normal_angle < M_PI / 8.0))) {
                    nn_idx[k] = ORC_NO_PAIR;
                    nn_d[k] = res.d_box;
                }
            }
        }
        size_t found = 0;
        for (size_t k = 0; k < n_it; k++) {
            if (nn_idx[k] == ORC_NO_PAIR) continue;
            memcpy(px + 3 * found, m->pt + 3 * (size_t)nn_idx[k], 3 * sizeof(double));
            memcpy(pp + 3 * found, node + 3 * k, 3 * sizeof(double));
            pd[found] = nn_d[k];
            found++;
        }
        double t1 = now_s();
        run->t_pairs += t1 - t0;
        const double in_d_box = res.d_box;
        (void)in_d_box;
        /* const size_t n_po = measurement.size() * result.overlap; */
        const size_t n_po = (size_t)((double)meas_size * res.overlap);
        double tau;
        kept = orc_trim(pd, found, n_po, order, &tau);
        res.d_box = tau * 2.0;
        res.pairs = kept;
        double t2 = now_s();
        run->t_trim += t2 - t1;

        orc_trace_row row;
        memset(&row, 0, sizeof(row));
        row.iter = (double)res.iter;
        row.found = (double)found;
        row.kept = (double)kept;
        row.tau = tau;
        row.d_box = res.d_box;

        /* remember this iteration's correspondences (debug / parity) */
        run->n_last = n_it;
        run->last_idx = (uint32_t *)realloc(run->last_idx, (n_it ? n_it : 1) * sizeof(uint32_t));
        run->last_dist = (double *)realloc(run->last_dist, (n_it ? n_it : 1) * sizeof(double));
        memcpy(run->last_idx, nn_idx, n_it * sizeof(uint32_t));
        memcpy(run->last_dist, nn_d, n_it * sizeof(double));

        if (res.pairs < 3) { /* Pairs::MIN_PAIRS: return before pairs_distance is filled */
            row.mse = res.mse;
            row.mse_diff = res.mse_diff;
            aff_to_cm(&C, row.C);
            trace_push(run, &row);
            early = 1;
            break;
        }
        aff T;
        double mse;
        get_transform(pp, px, pd, order, kept, &T, &mse);
        aff_mul(&T, &C, &C);              /* C = T * C */
        aff_mul(&C, &C_l2g, &C_l2g_new);  /* setOffsetTransform */
        res.mse = mse;
        res.mse_diff = old_mse - res.mse;
        res.iter++;
        run->t_solve += now_s() - t2;

        row.mse = res.mse;
        row.mse_diff = res.mse_diff;
        aff_to_cm(&C, row.C);
        trace_push(run, &row);
    }
    if (!early && run->n_trace > 0) {
        /* pairs.pairs[i].distance, i < pairs.size(): sorted kept distances of the last iteration */
        run->pairs_distance = (double *)malloc((kept ? kept : 1) * sizeof(double));
        for (size_t i = 0; i < kept; i++) run->pairs_distance[i] = pd[order[i]];
        run->n_pairs_distance = kept;
    }
    aff_to_cm(&C, res.C);
    *out = res;

    free(v);
    free(vn);
    free(vedge);
    free(node);
    free(nn_idx);
    free(nn_d);
    free(px);
    free(pp);
    free(pd);
    free(order);
    if (own) orc_run_free(own);
    return 0;
}

/* ------------------------------------------------------------------ */
/* golden-section search over the overlap (icp/icp.hpp:284-326, 385-401) */
/* ------------------------------------------------------------------ */
typedef struct {
    orc_result r;
    orc_run *run;
} gs_eval;

static double opt_func(const orc_result *r) { return r->mse / pow(r->overlap, 1.0 + 2.0); }

typedef struct {
    const orc_model *m;
    const double *xyz;
    size_t n;
    const double *T;
    double density;
    orc_params prm;
    int threads;
    int evals;
} gs_ctx;

static void gs_f(gs_ctx *c, double overlap, gs_eval *e)
{
    c->prm.overlap = overlap;
    orc_align(c->m, c->xyz, c->n, c->T, c->density, &c->prm, &e->r, e->run, c->threads);
    c->evals++;
}
static void gs_copy(gs_eval *dst, gs_eval *src)
{ /* value semantics of Result: swap the run buffers, copy the scalars */
    orc_run *t = dst->run;
    dst->run = src->run;
    src->run = t;
    dst->r = src->r;
}

int orc_align_golden(const orc_model *m, const double *xyz, size_t n, const double T[16], double density,
                     size_t max_iter, double min_mse, double min_mse_diff, double alpha, double beta, double eps,
                     orc_result *out, orc_run *run, int threads)
{
    gs_ctx c = {m, xyz, n, T, density, {max_iter, min_mse, min_mse_diff, 0.0}, threads, 0};
    const double R = 0.6180339;
    const double Cc = (1.0 - R);
    const double ax = alpha, bx = (alpha + beta) / 2.0, cx = beta;
    gs_eval f1, f2, ft;
    memset(&f1, 0, sizeof(f1));
    memset(&f2, 0, sizeof(f2));
    memset(&ft, 0, sizeof(ft));
    f1.run = orc_run_new();
    f2.run = orc_run_new();
    ft.run = orc_run_new();
    double x0, x1, x2, x3;
    x0 = ax;
    x3 = cx;
    if (fabs(cx - bx) > fabs(bx - ax)) {
        x1 = bx;
        x2 = bx + Cc * (cx - bx);
    } else {
        x2 = bx;
        x1 = bx - Cc * (bx - ax);
    }
    gs_f(&c, x1, &f1);
    gs_f(&c, x2, &f2);
    while (fabs(x3 - x0) > eps * (fabs(x1) + fabs(x2))) {
        if (opt_func(&f2.r) < opt_func(&f1.r)) {
            /* SHIFT3(x0,x1,x2,R*x1+C*x3); SHIFT2(f1,f2,f(x2)) */
            const double nx = R * x1 + Cc * x3;
            x0 = x1;
            x1 = x2;
            x2 = nx;
            gs_f(&c, x2, &ft);
            gs_copy(&f1, &f2);
            gs_copy(&f2, &ft);
        } else {
            /* SHIFT3(x3,x2,x1,R*x2+C*x0); SHIFT2(f2,f1,f(x1)) */
            const double nx = R * x2 + Cc * x0;
            x3 = x2;
            x2 = x1;
            x1 = nx;
            gs_f(&c, x1, &ft);
            gs_copy(&f2, &f1);
            gs_copy(&f1, &ft);
        }
    }
    gs_eval *best = (opt_func(&f1.r) < opt_func(&f2.r)) ? &f1 : &f2;
    *out = best->r;
    if (run) { /* hand the winner's buffers to the caller's run */
        run_reset(run);
        *run = *best->run;
        memset(best->run, 0, sizeof(orc_run));
    }
    orc_run_free(f1.run);
    orc_run_free(f2.run);
    orc_run_free(ft.run);
    return c.evals;
}

/* PointcloudAdapter::applyTransform (icp/icp.hpp:148-162) */
void orc_apply_transform(const double fn_T[16], const double C_l2g_cm[16], const double C_cm[16], double out_T[16])
{
    aff fn, l2g, C, inv, r;
    aff_from_cm(fn_T, &fn);
    aff_from_cm(C_l2g_cm, &l2g);
    aff_from_cm(C_cm, &C);
    aff_inverse_isometry(&l2g, &inv);
    aff_mul(&fn, &inv, &r); /* fn->getTransform() * C_local2global.inverse(Isometry) */
    aff_mul(&r, &C, &r);    /* * C_global2globalnew */
    aff_mul(&r, &l2g, &r);  /* * C_local2global */
    double rot[3][3];
    polar_rotation(r.L, rot); /* C_localnew2global.linear() = C_localnew2global.rotation() */
    memcpy(r.L, rot, sizeof(rot));
    aff_to_cm(&r, out_T);
}
