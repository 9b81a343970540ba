/* Plain C11 client of the C ABI: proves include/spinoza_b200.h is valid C and that the library needs nothing but
 * pointers and sizes.  Exit 0 + "C_ABI_OK" on a GPU box, exit 3 when no CUDA device is visible (no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/spinoza_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != SPZ_OK) { printf("%s -> %d (%s): %s\n", #call, rc_, spz_status_string(rc_), spz_last_error()); return 1; } \
    } while (0)

int main(void) {
    spz_state *st = NULL;
    if (spz_abi_version() != SPZ_ABI_VERSION) return 2;
    int rc = spz_create(3, 0, &st);
    if (spz_device_count() == 0) {
        printf("no device: rc=%d %s\n", rc, spz_last_error());
        return rc == SPZ_ERR_NO_DEVICE && st == NULL ? 3 : 4;
    }
    CHECK(rc);
    /* h_gate_3_qubits, gates.rs:1532-1544 */
    spz_gate h = {SPZ_GATE_H, 0, 0, 0, {0, 0, 0}};
    for (int t = 0; t < 3; ++t) CHECK(spz_apply(st, &h, t));
    double re[8], im[8];
    CHECK(spz_download(st, re, im, 0, 8));
    for (int i = 0; i < 8; ++i)
        if (fabs(re[i] - 0.35355339059327384) > 1e-10 || fabs(im[i]) > 1e-10) { printf("bad amplitude %d\n", i); return 1; }
    /* a fused execute: value encoding of 4 on 3 qubits -> |4> (circuit.rs:1076-1113) */
    CHECK(spz_reset_zero(st));
    spz_op ops[16];
    int n = 0;
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < 3; ++t) { spz_op o = {SPZ_GATE_H, t, 0, 0, {0, 0, 0}, SPZ_CTRL_NONE, 0, 0, 0}; ops[n++] = o; }
    for (int t = 0; t < 3; ++t) { spz_op o = {SPZ_GATE_P, t, 0, 0, {2.0 * PI / ldexp(1.0, t + 1) * 4.0, 0, 0}, SPZ_CTRL_NONE, 0, 0, 0}; ops[n++] = o; }
    const int targets[3] = {2, 1, 0}; /* iqft circuit.rs:438-445 */
    for (int j = 2; j >= 0; --j) {
        spz_op o = {SPZ_GATE_H, targets[j], 0, 0, {0, 0, 0}, SPZ_CTRL_NONE, 0, 0, 0};
        ops[n++] = o;
        for (int k = j - 1; k >= 0; --k) {
            spz_op c = {SPZ_GATE_P, targets[k], 0, 0, {-PI / ldexp(1.0, j - k), 0, 0}, SPZ_CTRL_SINGLE, 0, 1ull << targets[j], 0};
            ops[n++] = c;
        }
    }
    CHECK(spz_execute(st, ops, n, SPZ_EXEC_FUSE, NULL, NULL));
    CHECK(spz_download(st, re, im, 0, 8));
    for (int i = 0; i < 8; ++i)
        if (fabs(re[i] - (i == 4 ? 1.0 : 0.0)) > 1e-4 || fabs(im[i]) > 1e-4) { printf("value encoding failed at %d: %g %g\n", i, re[i], im[i]); return 1; }
    int bit = -1;
    CHECK(spz_measure_qubit(st, 2, 1, -1, &bit));
    if (bit != 1) { printf("measured %d, expected 1\n", bit); return 1; }
    spz_gate m = {SPZ_GATE_M, 0, 0, 0, {0, 0, 0}};
    if (spz_apply(st, &m, 0) != SPZ_ERR_UNSUPPORTED) { printf("apply(M) must be unsupported\n"); return 1; }
    CHECK(spz_destroy(st));
    printf("C_ABI_OK launches=%lld\n", (long long)spz_launch_count());
    return 0;
}
