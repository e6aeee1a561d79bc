/* A C consumer of the boundary: the entry points that are pure host code (no GPU needed) called from plain C99.
 *
 *   gcc -std=c99 -I include examples/host_entry_points.c -o /tmp/hep -L cindm_b200/lib -lcindm_b200 -Wl,-rpath,$PWD/cindm_b200/lib -lm
 *
 * tests/test_capi_symbols.py builds and runs it.  It checks the known answers SURVEY.md section 8(c) quotes for the cosine
 * schedule (reference model/diffusion_1d.py:470-480, :853-897), the cover counts of the C4 composition (28 pairs x 3
 * windows of 24 rows over 44 rows, :977-990), and that a bad configuration comes back as an error code + message. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cindm_b200.h"

static int fail(const char* what) {
    fprintf(stderr, "FAILED: %s\n", what);
    return 1;
}

int main(void) {
    if (cindm_version() < 100) return fail("version");

    float* tab = (float*)malloc(13 * 1000 * sizeof(float));
    if (cindm_schedule_tables(1000, tab) != 0) return fail("cindm_schedule_tables");
    const float* betas = tab;                         /* table 0 */
    const float* acp = tab + 1000;                    /* table 1: alphas_cumprod */
    const float* post_logvar = tab + 9 * 1000;        /* table 9: posterior_log_variance_clipped */
    if (fabs(betas[0] - 4.1284e-5) > 1e-8 || fabs(betas[999] - 0.999) > 1e-6) return fail("betas");
    if (fabs(acp[999] - 2.4288e-9) > 1e-12) return fail("alphas_cumprod[999]");
    if (fabs(post_logvar[0] + 46.0517) > 1e-3) return fail("posterior_log_variance_clipped[0]");

    int32_t win[3], pi[28], pj[28], cover[44];
    if (cindm_build_index_maps(8, 2, 10, 24, win, pi, pj, cover) != 0) return fail("cindm_build_index_maps");
    int total = 0;
    for (int t = 0; t < 44; ++t) total += cover[t];
    if (win[0] != 0 || win[1] != 10 || win[2] != 20 || total != 3 * 24) return fail("windows / cover");
    if (pi[0] != 0 || pj[0] != 1 || pi[27] != 6 || pj[27] != 7) return fail("pair list");

    cindm_config bad = {45, 8, 64, 1000};             /* odd horizon: undefined in the reference */
    cindm_engine* e = NULL;
    if (cindm_create(&bad, &e) >= 0 || strlen(cindm_last_error()) == 0) return fail("error reporting");

    printf("ok: version %d, betas[0] = %.4e, cover sum = %d, refusal: %s\n", cindm_version(), betas[0], total, cindm_last_error());
    free(tab);
    return 0;
}
