/* Minimal C consumer of the C-ABI in include/tramba_b200.h (plain C99, no CUDA headers, no Python):
 *
 *     gcc -std=c99 -Wall -Iinclude examples/abi_demo.c -Ltramba_b200 -ltramba_b200 -Wl,-rpath,$PWD/tramba_b200 -o abi_demo
 *
 * Host-only entry points (scan-order tables, plans, DCT basis) run anywhere; the compute entry points need an
 * sm_100 device and otherwise return TRAMBA_E_NODEVICE -- shown at the end.  tests/test_boundary.py builds and runs
 * this file, which is also the proof that the header is valid C and the library links without torch. */
#include <stdio.h>
#include <stdlib.h>

#include "tramba_b200.h"

static int check_permutation(const int32_t* t, int L) {
    char* seen = (char*)calloc((size_t)L, 1);
    int ok = 1, i;
    for (i = 0; i < L; ++i) {
        if (t[i] < 0 || t[i] >= L || seen[t[i]]) ok = 0;
        else seen[t[i]] = 1;
    }
    free(seen);
    return ok;
}

int main(void) {
    const int H = 24, L = H * H;
    int32_t* table;
    int32_t* plan;
    float basis[16];
    int K, rc, k;

    printf("abi %d\n", tramba_abi_version());
    if (tramba_abi_version() != TRAMBA_ABI_VERSION) return 1;

    /* window order (Models/SS2D/Window.py, ws = 8 at H = 24): four permutations of the pixels */
    K = tramba_order_num_directions(TRAMBA_ORDER_WINDOW);
    table = (int32_t*)malloc(sizeof(int32_t) * (size_t)K * L);
    rc = tramba_order_table(TRAMBA_ORDER_WINDOW, H, H, 8, table);
    if (rc != TRAMBA_OK) return 2;
    for (k = 0; k < K; ++k)
        if (!check_permutation(table + (size_t)k * L, L)) return 3;
    printf("window K=%d first tokens %d %d %d %d\n", K, table[0], table[1], table[8], table[L]);

    /* helix order: the plan (forward table + CSR inverse) the device kernels consume */
    K = tramba_order_num_directions(TRAMBA_ORDER_LINE);
    plan = (int32_t*)malloc(sizeof(int32_t) * tramba_order_plan_elems(K, L));
    rc = tramba_order_plan(TRAMBA_ORDER_LINE, H, H, 0, plan);
    if (rc != TRAMBA_OK) return 4;
    printf("helix K=%d plan elems %lu\n", K, (unsigned long)tramba_order_plan_elems(K, L));

    /* errors are codes, never exceptions or aborts */
    rc = tramba_order_table(TRAMBA_ORDER_WINDOW, H, H, 5, table); /* 24 % 5 != 0 */
    printf("bad window size -> %d (%s)\n", rc, tramba_error_string(rc));
    if (rc != TRAMBA_E_ORDER) return 5;

    tramba_dct_basis(4, basis);
    printf("dct basis[0][0] = %.6f\n", basis[0]);

    /* a compute entry point: refused unless an sm_100 device is current.  The buffers here are HOST memory, so
     * this call is only made on a box that is known to have no GPU (TRAMBA_DEMO_NO_DEVICE=1); with a device the
     * pointers must be device addresses. */
    if (getenv("TRAMBA_DEMO_NO_DEVICE")) {
        rc = tramba_cross_scan(table, table, plan, TRAMBA_ORDER_LINE, K, 1, 1, H, H, TRAMBA_F32, NULL);
        printf("cross_scan without a device -> %d (%s)\n", rc, tramba_error_string(rc));
        if (rc != TRAMBA_E_NODEVICE) return 6;
    }
    free(table);
    free(plan);
    return 0;
}
