/* The boundary is a C ABI: this file must compile as plain C99 against include/pfe_b200.h and link against
 * libpfe_b200.so.  Without a GPU it checks the no-device behaviour; with one it runs a tiny flatten + blur
 * through the host-pointer tier exactly as a C caller (or a cgo / Rust FFI shim) would. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pfe_b200.h"

int main(void) {
    pfe_ctx *ctx = NULL;
    int rc = pfe_ctx_create(0, &ctx);
    if (rc == PFE_ERR_NO_DEVICE) {
        printf("no device: pfe_ctx_create refused (rc %d), abi version %d\n", rc, pfe_abi_version());
        return 0;
    }
    if (rc != PFE_OK) { printf("pfe_ctx_create failed: %d\n", rc); return 1; }
    enum { W = 96, H = 80 };
    uint8_t *a = malloc(W * H * 4), *b = malloc(W * H * 4), *flat = malloc(W * H * 4), *blur = malloc(W * H * 4);
    for (int i = 0; i < W * H; i++) {
        a[i * 4] = (uint8_t)(i * 7); a[i * 4 + 1] = (uint8_t)(i * 3); a[i * 4 + 2] = 40; a[i * 4 + 3] = 255;
        b[i * 4] = 200; b[i * 4 + 1] = (uint8_t)(i * 5); b[i * 4 + 2] = (uint8_t)i; b[i * 4 + 3] = (uint8_t)(i % 256);
    }
    pfe_layer_desc layers[2];
    memset(layers, 0, sizeof(layers));
    layers[0].rgba = a; layers[0].opacity = 1.0f; layers[0].visible = 1;
    layers[1].rgba = b; layers[1].opacity = 0.5f; layers[1].visible = 1; layers[1].blend = 8; /* Overlay */
    rc = pfe_flatten(ctx, layers, 2, W, H, NULL, flat);
    if (rc == PFE_OK) rc = pfe_gaussian_blur(ctx, flat, W, H, 2.0f, NULL, blur, PFE_GAUSS_EXACT);
    if (rc != PFE_OK) { printf("call failed: %d (%s)\n", rc, pfe_last_error(ctx)); return 1; }
    unsigned long sum = 0;
    for (int i = 0; i < W * H * 4; i++) sum += blur[i];
    printf("ok: launches %llu, checksum %lu\n", (unsigned long long)pfe_ctx_launch_count(ctx), sum);
    pfe_ctx_destroy(ctx);
    free(a); free(b); free(flat); free(blur);
    return 0;
}
