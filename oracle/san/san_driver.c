/* Sanitizer build of the oracle (test infrastructure): exercises model add (plain), align (fixed and golden-section
 * overlap), density stepping, ties and tiny / empty inputs under -fsanitize=address,undefined.
 * Built and run by tests/test_oracle.py::test_oracle_under_sanitizers; prints "SAN_OK" when nothing tripped. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../icp_oracle.h"

static double rnd(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / 16777216.0; }

int main(void)
{
    unsigned seed = 7;
    const size_t n = 4000, ns = 3000;
    double *m = malloc(n * 3 * sizeof(double)), *s = malloc(ns * 3 * sizeof(double));
    for (size_t i = 0; i < n; ++i) {
        const double x = rnd(&seed) * 10 - 5, y = rnd(&seed) * 10 - 5;
        m[3 * i] = x, m[3 * i + 1] = y, m[3 * i + 2] = 0.3 * sin(0.5 * x) * cos(0.4 * y);
    }
    for (size_t i = 0; i < ns; ++i) { /* a slightly rotated / shifted copy */
        const double x = m[3 * i], y = m[3 * i + 1], z = m[3 * i + 2], c = cos(0.03), sn = sin(0.03);
        s[3 * i] = c * x - sn * y + 0.05, s[3 * i + 1] = sn * x + c * y - 0.04, s[3 * i + 2] = z + 0.02;
    }
    double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    orc_model *M = orc_model_new();
    orc_model_add(M, m, n / 2, I, 1.0);
    orc_model_add(M, m + 3 * (n / 2), n - n / 2, I, 0.7); /* accumulates, density-stepped */
    orc_params p = {20, 1e-10, 1e-12, 0.8};
    orc_result r;
    orc_run *run = orc_run_new();
    int rc = orc_align(M, s, ns, I, 0.37, &p, &r, run, 2);
    double pd[4096];
    const size_t npd = orc_run_pairs_distance(run, pd, 4096);
    printf("align rc %d iter %zu pairs %zu mse %.3g pd %zu\n", rc, (size_t)r.iter, (size_t)r.pairs, r.mse, npd);
    rc = orc_align_golden(M, s, ns, I, 1.0, 6, 1e-10, 1e-12, 0.5, 1.0, 0.1, &r, run, 2);
    printf("golden evals %d overlap %.4f\n", rc, r.overlap);
    /* ties, tiny and empty inputs */
    double lat[27 * 3];
    for (int i = 0; i < 27; ++i) lat[3 * i] = i % 3, lat[3 * i + 1] = (i / 3) % 3, lat[3 * i + 2] = i / 9;
    orc_model_clear(M);
    orc_model_add(M, lat, 27, I, 1.0);
    double q[3] = {0.5, 0.5, 0.5};
    uint32_t idx;
    double d;
    orc_find_pairs(M, q, 1, INFINITY, &idx, &d, 1);
    orc_find_pairs(M, q, 0, 1.0, &idx, &d, 1);
    orc_params p2 = {5, 1e-4, 1e-5, 1.0};
    rc = orc_align(M, q, 1, I, 1.0, &p2, &r, run, 1);
    rc = orc_align(M, q, 0, I, 1.0, &p2, &r, run, 1);
    orc_model_clear(M);
    rc = orc_align(M, s, 10, I, 1.0, &p2, &r, run, 1); /* empty model */
    orc_run_free(run);
    orc_model_free(M);
    free(m);
    free(s);
    printf("SAN_OK\n");
    return 0;
}
